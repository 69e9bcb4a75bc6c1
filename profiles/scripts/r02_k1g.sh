cd $GRAFT_REPO_ROOT
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
run t2_128 DBB_K1_T2=128
run t2_512 DBB_K1_T2=512
run t1_48 DBB_K1_T1=48
run t1_16 DBB_K1_T1=16
run w22 DBB_K1_WPI2=2
run w34 DBB_K1_WPI3=4
run t3_1000 DBB_K1_T3=1000
( timeout 600 python bench.py --workload C5 --scale 0.1 --steps 5 --no-cpu ) > gpurun_out/k1_c5.json 2> gpurun_out/k1_c5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/k1_c5.json').read().strip().splitlines()[-1])
print('C5 x0.1', round(d['stages_ms']['signatures'],4), round(d['roofline']['frac'],4), d['value'])
PY
