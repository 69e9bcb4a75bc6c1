cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print("e2e Gbp/s %.2f pageable %.2f ceiling %.1f; verified %s" % (e["value"], e["value_pageable_host_buffers"], e["h2d_concurrent_ceiling_gbs"], d["verified_vs_oracle"]))'
for b in 32 64 128 256 512; do echo "== batch MB $b"; DBB_HOST_BATCH_MB=$b timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --check-rows 8 2> gpurun_out/hb.err | python -c "$show"; done
