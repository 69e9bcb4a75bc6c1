cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel --launch-skip 9 --launch-count 3 -o gpurun_out/r02_count_final -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --check-rows 8 > gpurun_out/ncu_count_final.log 2>&1
ls -la gpurun_out/r02_count_final.ncu-rep
