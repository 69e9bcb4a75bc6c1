cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/n2_c2.json 2> gpurun_out/n2_c2.err
tail -n 3 gpurun_out/n2_c2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --workload C3 ) > gpurun_out/n2_c3.json 2> gpurun_out/n2_c3.err
tail -n 3 gpurun_out/n2_c3.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu --no-e2e --workload C4 ) > gpurun_out/n2_c4.json 2> gpurun_out/n2_c4.err
tail -n 3 gpurun_out/n2_c4.err
