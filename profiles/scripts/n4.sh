set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --no-cpu --no-e2e ) > gpurun_out/pn_n4.json 2> gpurun_out/pn_n4.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --workload C3 --no-cpu --no-e2e --steps 3 ) > gpurun_out/pn_n4_c3.json 2> gpurun_out/pn_n4_c3.err
tail -n 4 gpurun_out/pn_n4.err gpurun_out/pn_n4_c3.err
