cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["stages_ms"], d["pruning"]["ms_sweep"], d["pruning"]["ms_second_sweep_and_merge"], d["verified_vs_oracle"], d["roofline_pairs"]["frac"])'
for it in 2960 6000 10000 16000; do echo ITEMS $it; DBB_PRUNE_ITEMS=$it timeout 300 python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu --no-e2e --check-rows 16 2> gpurun_out/items.err | python -c "$show"; done
