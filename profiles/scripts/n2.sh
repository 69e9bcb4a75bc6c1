set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu ) > gpurun_out/pn_n2.json 2> gpurun_out/pn_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C3 --no-cpu --steps 3 ) > gpurun_out/pn_n2_c3.json 2> gpurun_out/pn_n2_c3.err
tail -n 6 gpurun_out/pn_n2.err gpurun_out/pn_n2_c3.err
