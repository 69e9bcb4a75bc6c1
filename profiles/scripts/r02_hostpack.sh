cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)" 
timeout 600 python -m pytest tests/test_hostpack.py tests/test_abi.py tests/test_signatures_gpu.py -m gpu -q -x 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print("e2e Gbp/s %.1f pageable %.1f ceiling %.1f GB/s; upload %s; verified %s" % (e["value"], e["value_pageable_host_buffers"], e["h2d_concurrent_ceiling_gbs"], {k: v for k, v in e["upload"].items() if k != "note"}, d["verified_vs_oracle"]))'
echo "== pack off"; DBB_HOST_PACK=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --check-rows 8 2> gpurun_out/hp.err | python -c "$show"
for t in 4 8 12 16 24; do echo "== threads $t"; DBB_HOST_PACK_THREADS=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --check-rows 8 2> gpurun_out/hp.err | python -c "$show"; done
echo "== default"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --check-rows 8 2> gpurun_out/hp.err | python -c "$show"
tail -3 gpurun_out/hp.err
