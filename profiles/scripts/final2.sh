set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/fin2_tests.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/fin2_smoke.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/fin2_bench.json 2> gpurun_out/fin2_bench.err
tail -n 12 gpurun_out/fin2_tests.log; tail -n 3 gpurun_out/fin2_smoke.log gpurun_out/fin2_bench.err
