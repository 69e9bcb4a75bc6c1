cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:count_kernel<.int.1," --launch-skip 3 --launch-count 1 -o gpurun_out/r02_s1_c5 -f python bench.py --workload C5 --scale 0.05 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_s1.log 2>&1
ls -la gpurun_out/r02_s1_c5.ncu-rep; tail -3 gpurun_out/ncu_s1.log | cut -c1-300
