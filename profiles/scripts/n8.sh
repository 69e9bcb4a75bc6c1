set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu ) > gpurun_out/pn_n8.json 2> gpurun_out/pn_n8.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload C4 --no-cpu --steps 3 ) > gpurun_out/pn_n8_c4.json 2> gpurun_out/pn_n8_c4.err
tail -n 6 gpurun_out/pn_n8.err gpurun_out/pn_n8_c4.err
