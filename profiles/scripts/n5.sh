set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_merge_gpu.py tests/test_pairs_gpu.py -m gpu -x -q ) > gpurun_out/n5_tests.log 2>&1
( time timeout 600 python tests/perf_n5_merge.py ) > gpurun_out/n5_perf.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ll_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ll_ncu.log 2>&1
tail -n 4 gpurun_out/n5_tests.log gpurun_out/n5_perf.log
