# round-end check: what the driver runs (GPU tests, smoke, both bench arms) + the N6 timing script
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/fin_tests.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/fin_smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err
( time timeout 300 python tests/perf_n6_unbinned.py ) > gpurun_out/n6_perf.log 2>&1
tail -n 4 gpurun_out/fin_tests.log gpurun_out/fin_smoke.log gpurun_out/fin_bench.err gpurun_out/fin_ref.err gpurun_out/n6_perf.log
