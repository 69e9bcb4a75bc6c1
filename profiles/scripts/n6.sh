set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_merge_gpu.py tests/test_pairs_gpu.py -m gpu -x -q ) > gpurun_out/n6_tests.log 2>&1
tail -n 30 gpurun_out/n6_tests.log
