cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 0 1; do
DBB_K1_DEBUG_SKIP_FLUSH=$f python profiles/microbench/k1_classes.py S3w2 96001
DBB_K1_DEBUG_SKIP_FLUSH=$f DBB_K1_T3=513 DBB_K1_WPI4=8 python profiles/microbench/k1_classes.py S4w8 96001
DBB_K1_DEBUG_SKIP_FLUSH=$f DBB_K1_T3=513 DBB_K1_WPI4=4 python profiles/microbench/k1_classes.py S4w4 96001
DBB_K1_DEBUG_SKIP_FLUSH=$f DBB_K1_T3=513 DBB_K1_WPI4=8 python profiles/microbench/k1_classes.py S4w8 48001
DBB_K1_DEBUG_SKIP_FLUSH=$f DBB_K1_T3=513 DBB_K1_WPI4=8 python profiles/microbench/k1_classes.py S4w8 250001
done 2>&1 | grep -v Warning
DBB_K1_T3=513 DBB_K1_WPI4=8 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:count_kernel<.int.4," --launch-skip 3 --launch-count 1 -o gpurun_out/r02_s4w8 -f python profiles/microbench/k1_classes.py S4w8 96001 > gpurun_out/ncu_s4.log 2>&1
ls -la gpurun_out/r02_s4w8.ncu-rep
