cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_signatures_gpu.py tests/test_pairs_gpu.py tests/test_abi.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["roofline"]["frac"], d["stages_ms"], d["pruning"]["ms_sweep"], d["pruning"]["ms_second_sweep_and_merge"], d["verified_vs_oracle"])'
for b in 0 1 0 1; do echo BALANCE $b; DBB_K1_BALANCE=$b timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --check-rows 16 2> gpurun_out/g2.err | python -c "$show"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairs_kernel --launch-skip 6 --launch-count 1 -o gpurun_out/r02_pairs_gthr -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --check-rows 8 > gpurun_out/ncu_pairs_gthr.log 2>&1
ls -la gpurun_out/r02_pairs_gthr.ncu-rep
