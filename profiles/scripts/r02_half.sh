cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["stages_ms"], d["pruning"]["ms_sweep"], d["pruning"]["ms_second_sweep_and_merge"], d["verified_vs_oracle"], d["roofline_pairs"])'
for w in 8 4 8 4; do echo WARPS $w; DBB_K2_WARPS=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --check-rows 32 2> gpurun_out/half.err | python -c "$show"; done
echo C3 scale 0.4
for w in 8 4; do echo WARPS $w; DBB_K2_WARPS=$w timeout 300 python bench.py --workload C3 --scale 0.4 --steps 3 --warmup 2 --no-cpu --no-e2e --check-rows 32 2> gpurun_out/half.err | python -c "$show"; done
echo full sweep
for w in 8 4; do echo WARPS $w; DBB_K2_WARPS=$w timeout 300 python bench.py --no-prune --steps 3 --warmup 2 --no-cpu --no-e2e --check-rows 32 2> gpurun_out/half.err | python -c "$show"; done
