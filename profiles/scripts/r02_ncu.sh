cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --check-rows 8 > gpurun_out/ll_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --check-rows 8 > gpurun_out/ll_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel --launch-skip 9 --launch-count 3 -o gpurun_out/r02_count -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --check-rows 8 > gpurun_out/ncu_count.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairs_kernel --launch-skip 6 --launch-count 1 -o gpurun_out/r02_pairs -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --check-rows 8 > gpurun_out/ncu_pairs.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
