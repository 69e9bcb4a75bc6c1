cd $GRAFT_REPO_ROOT
K=profiles/microbench/k1_classes.py
timeout 300 python -m pytest tests/test_signatures_gpu.py -m gpu -x -q 2>&1 | tail -3
{
for L in 1501 3001 6001; do for F in 0 1; do DBB_K1_DEBUG_SKIP_FLUSH=$F DBB_K1_T1=4000 timeout 120 python $K S1 $L; done; done
for W in 1 2; do for L in 3001 6001 12001; do for F in 0 1; do DBB_K1_WPI2=$W DBB_K1_DEBUG_SKIP_FLUSH=$F DBB_K1_T1=0 DBB_K1_T2=2047 timeout 120 python $K S2w$W $L; done; done; done
for W in 1 2 4; do for L in 6001 12001 24001 48001 96001; do for F in 0 1; do DBB_K1_WPI3=$W DBB_K1_DEBUG_SKIP_FLUSH=$F DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=3071 timeout 120 python $K S3w$W $L; done; done; done
for W in 2 4 8; do for L in 24001 48001 96001 250001; do for F in 0 1; do DBB_K1_WPI4=$W DBB_K1_DEBUG_SKIP_FLUSH=$F DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=0 timeout 120 python $K S4w$W $L; done; done; done
} 2>&1 | grep -v Warning | tee gpurun_out/k1_classes_e.txt
