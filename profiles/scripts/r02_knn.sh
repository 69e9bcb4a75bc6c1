cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/knn_tests.log 2>&1
tail -n 25 gpurun_out/knn_tests.log
( time timeout 600 python bench.py --steps 5 --no-cpu ) > gpurun_out/knn_bench_c2.json 2> gpurun_out/knn_bench_c2.err
tail -n 5 gpurun_out/knn_bench_c2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/knn_bench_c2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','value2','ms_per_step','stages_ms','verified_vs_oracle','oracle_check_stage2','pruning','e2e','launches_per_step')})
PY
