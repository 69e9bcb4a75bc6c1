# round 2, call 1: atomic-format ceilings (mb4), sanity of the sum-to-1 check, bench with the oracle row-sample check
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 profiles/microbench/mb4 ) > gpurun_out/r02_mb4.txt 2>&1
( time timeout 900 python -m pytest tests/test_pairs_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q ) > gpurun_out/c1_tests.log 2>&1
( time timeout 600 python bench.py --steps 5 --no-cpu ) > gpurun_out/c1_bench_c2.json 2> gpurun_out/c1_bench_c2.err
( time timeout 900 python bench.py --workload C4 --scale 0.4 --steps 2 --no-cpu --no-e2e ) > gpurun_out/c1_bench_c4s.json 2> gpurun_out/c1_bench_c4s.err
( time timeout 900 python bench.py --workload C3 --steps 2 --no-cpu --no-e2e ) > gpurun_out/c1_bench_c3.json 2> gpurun_out/c1_bench_c3.err
tail -n 5 gpurun_out/r02_mb4.txt gpurun_out/c1_tests.log gpurun_out/c1_bench_c2.err gpurun_out/c1_bench_c4s.err gpurun_out/c1_bench_c3.err
