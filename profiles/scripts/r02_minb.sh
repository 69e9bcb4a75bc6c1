cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_signatures_gpu.py -m gpu -q -x 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["stages_ms"]["signatures"], d["roofline"]["frac"], d["verified_vs_oracle"])'
for c in "12 12" "16 12" "12 16" "16 16" "12 12" "16 16"; do set -- $c; echo "MINB2 $1 MINB3 $2"; DBB_K1_MINB2=$1 DBB_K1_MINB3=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --check-rows 8 2> gpurun_out/minb.err | python -c "$show"; done
