set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_pairs_gpu.py -m gpu -x -q ) > gpurun_out/pn_tests.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/pn_bench.json 2> gpurun_out/pn_bench.err
( timeout 600 python profiles/microbench/k2_prune_bench.py C2 1.0 8 ) > gpurun_out/pn_k2.log 2>&1
tail -n 25 gpurun_out/pn_tests.log; tail -n 5 gpurun_out/pn_bench.err gpurun_out/pn_k2.log
