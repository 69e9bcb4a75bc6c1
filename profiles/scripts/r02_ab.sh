cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pairs_gpu.py tests/test_hostpack.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["stages_ms"], d["pruning"]["ms_sweep"], d["pruning"]["ms_second_sweep_and_merge"], d["verified_vs_oracle"], d["roofline_pairs"]["frac"])'
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --check-rows 32 2> gpurun_out/ab.err | python -c "$show"; done
timeout 300 python bench.py --workload C3 --scale 0.4 --steps 3 --warmup 2 --no-cpu --no-e2e --check-rows 32 2> gpurun_out/ab.err | python -c "$show"
