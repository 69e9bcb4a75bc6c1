cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K=profiles/microbench/k1_classes.py
{
for L in 1501 3001 6001; do DBB_K1_T1=4000 timeout 120 python $K S1 $L; done
for L in 1501 3001 6001 12001 24001; do DBB_K1_T1=0 DBB_K1_T2=2047 timeout 120 python $K S2 $L; DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=2047 timeout 120 python $K S2 $L; done
for L in 6001 12001 24001 48001 96001; do DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=3071 timeout 120 python $K S3 $L; DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=3071 timeout 120 python $K S3 $L; done
for L in 24001 48001 96001 250001; do DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=0 timeout 120 python $K S4 $L; DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=0 timeout 120 python $K S4 $L; done
} 2>&1 | grep -v Warning | tee gpurun_out/k1_classes.txt
