cd $GRAFT_REPO_ROOT
K=profiles/microbench/k1_classes.py
{
for M in 400 40; do
DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=3071 timeout 120 python $K S3_${M}Mbp 48001 $M
DBB_K1_DEPTH3=4 DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=3071 timeout 120 python $K S3d4_${M}Mbp 48001 $M
DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=0 timeout 120 python $K S4_${M}Mbp 96001 $M
DBB_K1_DEPTH4=6 DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=0 DBB_K1_T2=0 DBB_K1_T3=0 timeout 120 python $K S4d6_${M}Mbp 96001 $M
DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=4000 timeout 120 python $K S1_${M}Mbp 6001 $M
DBB_K1_DEBUG_SKIP_FLUSH=1 DBB_K1_T1=4000 timeout 120 python $K S1_${M}Mbp 1501 $M
done
} 2>&1 | grep -v Warning | tee gpurun_out/k1_classes_c.txt
