# what the round-end driver does (GPU tests, smoke, bench, reference arm) + the ncu launch list of the bench command
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_tests.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/f_smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ll_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ll_ncu.log 2>&1

tail -n 4 gpurun_out/f_tests.log gpurun_out/f_smoke.log gpurun_out/f_bench.err gpurun_out/f_ref.err
