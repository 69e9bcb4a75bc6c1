cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_signatures_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -2
DBB_K1_T3=513 timeout 600 python -m pytest tests/test_signatures_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "signat or count or exact or kernel_class or fullsize or c2 or C2" 2>&1 | tail -2
for L in 24001 48001 96001 250001; do DBB_K1_T3=300 DBB_K1_WPI4=8 python profiles/microbench/k1_classes.py S4w8 $L; done 2>&1 | grep -v Warn
DBB_K1_T3=300 DBB_K1_WPI4=4 python profiles/microbench/k1_classes.py S4w4 48001 2>&1 | grep -v Warn
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["stages_ms"]["signatures"], d["roofline"]["frac"], d["verified_vs_oracle"])'
for t in 1696 1280 1024 768 640 512 384; do echo "T3 $t"; DBB_K1_T3=$t timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --check-rows 8 2> gpurun_out/f4.err | python -c "$show"; done
