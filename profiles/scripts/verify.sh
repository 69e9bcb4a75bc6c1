set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_tests.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/v_smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/v_ref.json 2> gpurun_out/v_ref.err
tail -3 gpurun_out/v_tests.log gpurun_out/v_smoke.log gpurun_out/v_bench.err gpurun_out/v_ref.err
