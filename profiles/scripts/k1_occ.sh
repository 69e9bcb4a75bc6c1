# co-running the class kernels of stage 1: CTAs per SM of each class capped by DBB_K1_OCC<S>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/k1_occ.txt
run() {
  echo "OCC3=$1 OCC2=$2 OCC1=$3" >> gpurun_out/k1_occ.txt
  DBB_K1_OCC3=$1 DBB_K1_OCC2=$2 DBB_K1_OCC1=$3 timeout 120 python bench.py --no-cpu --no-e2e --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  signatures ms', round(d['stages_ms']['signatures'],4), 'value', round(d['value'],1), 'verified', d['verified_vs_oracle'])" >> gpurun_out/k1_occ.txt 2>&1
}
run 0 0 0
run 3 4 8
run 3 8 0
run 2 8 8
run 3 4 0
run 2 6 12
run 0 8 8
run 0 0 0
cat gpurun_out/k1_occ.txt
