cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_pairs_gpu.py tests/test_abi.py -m gpu -q 2>&1 | tail -3
for sg in 0 24 32 45 64 96; do
  echo SEG_TILES $sg
  DBB_PRUNE_SEG_TILES=$sg DBB_K2_STATS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --check-rows 8 2> gpurun_out/seg.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stages_ms']['knn'], d['pruning']['ms_sweep'], d['verified_vs_oracle'])"
  grep "over 148" gpurun_out/seg.err | sort | uniq -c | sort -rn | head -2
done
