# K1 v2: correctness + timing sweeps
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_signatures_gpu.py tests/test_fullsize_gpu.py tests/test_preprocess_gpu.py -m gpu -x -q ) > gpurun_out/k1_tests.log 2>&1
tail -n 15 gpurun_out/k1_tests.log
run() { tag=$1; shift; ( env "$@" timeout 300 python bench.py --steps 10 --no-cpu --no-e2e --check-rows 8 ) > gpurun_out/k1_$tag.json 2> gpurun_out/k1_$tag.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/k1_$tag.json').read().strip().splitlines()[-1])
    print('$tag', '$*', round(d['stages_ms']['signatures'],4), round(d['roofline']['frac'],4))
except Exception as e:
    print('$tag FAILED', e)
PY
}
run base A=1
run t1_32 DBB_K1_T1=32
run t1_64 DBB_K1_T1=64
run t2_128 DBB_K1_T2=128
run t2_256 DBB_K1_T2=256
run t3_800 DBB_K1_T3=800
run t3_1200 DBB_K1_T3=1200
run t3_3071 DBB_K1_T3=3071
run w1 DBB_K1_WARPS1=1 DBB_K1_WARPS2=1
run skipflush DBB_K1_DEBUG_SKIP_FLUSH=1
( timeout 600 python bench.py --workload C5 --scale 0.1 --steps 5 --no-cpu ) > gpurun_out/k1_c5.json 2> gpurun_out/k1_c5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/k1_c5.json').read().strip().splitlines()[-1])
print('C5 x0.1', round(d['stages_ms']['signatures'],4), round(d['roofline']['frac'],4), d['value'])
PY
