cd $GRAFT_REPO_ROOT
K=profiles/microbench/k1_classes.py
DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:count_kernel -c 2 -o gpurun_out/k1_s4 -f python $K S4 96001 > gpurun_out/k1_ncu_s4.log 2>&1
DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=3071 timeout 300 ncu --set full --clock-control none --import-source on -k regex:count_kernel -c 2 -o gpurun_out/k1_s3 -f python $K S3 24001 > gpurun_out/k1_ncu_s3.log 2>&1
DBB_K1_T1=4000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:count_kernel -c 2 -o gpurun_out/k1_s1 -f python $K S1 1501 > gpurun_out/k1_ncu_s1.log 2>&1
ls -la gpurun_out/*.ncu-rep
