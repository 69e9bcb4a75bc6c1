cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hostpack.py tests/test_abi.py tests/test_signatures_gpu.py tests/test_preprocess_gpu.py tests/test_greedy_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print("e2e Gbp/s %.2f pageable %.2f class-list %.2f ceiling %.1f; verified %s" % (e["value"], e["value_pageable_host_buffers"], e["value_class_api_list_of_buffers"], e["h2d_concurrent_ceiling_gbs"], d["verified_vs_oracle"]))'
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --check-rows 8 2> gpurun_out/pt.err | python -c "$show"
tail -3 gpurun_out/pt.err
