# usage: r02_scale.sh N "C4 C3 C2 C5"
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
port=29600
for W in $2; do
  port=$((port+1))
  extra="--no-e2e"
  steps=3
  [ "$W" = "C2" ] && extra="" && steps=5
  [ "$W" = "C4" ] && steps=2
  if [ "$N" = "1" ]; then
    ( time timeout 900 python bench.py --gpus 1 --steps $steps --warmup 3 --no-cpu $extra --workload $W ) > gpurun_out/r02_bench_n${N}_${W}.json 2> gpurun_out/r02_bench_n${N}_${W}.err
  else
    ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps $steps --warmup 3 --no-cpu $extra --workload $W ) > gpurun_out/r02_bench_n${N}_${W}.json 2> gpurun_out/r02_bench_n${N}_${W}.err
  fi
  tail -n 4 gpurun_out/r02_bench_n${N}_${W}.err
done
