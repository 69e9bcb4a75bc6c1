cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["stages_ms"]["signatures"], d["roofline"]["frac"], d["verified_vs_oracle"])'
for t in 32 24 16 8 0; do for w in 1 2; do echo "T1 $t WPI2 $w"; DBB_K1_T1=$t DBB_K1_WPI2=$w timeout 300 python bench.py --workload C5 --scale 0.1 --steps 5 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/t1.err | python -c "$show"; done; done
