import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dbb_b200 import device, synthetic
from dbb_b200.distributions import PairTables
from dbb_b200.data import load_td_dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.RandomState(0)
meta = synthetic.make_metagenome(seed=2, n_scaffolds=n, min_len=1000, max_len=100000, n_genomes=200)
# signatures: random dirichlet around genome centroids (no counting needed for this timing experiment)
cent = rng.dirichlet(np.ones(136) * 5, 200)
sig = (cent[meta.genome] * rng.gamma(50, 1 / 50.0, (n, 136))); sig /= sig.sum(1, keepdims=True)
f = torch.zeros((n, 140), dtype=torch.float32, device='cuda'); f[:, :136] = torch.from_numpy(sig.astype(np.float32)).cuda()
dt = device.DevicePairTables(PairTables(meta.length, tdDist=load_td_dist(), tdDistPer=80))
def run(label, **env):
    for k_, v in env.items(): os.environ[k_] = v
    for _ in range(2): device.pairs(f, dt, k=20, want_count=True)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): device.pairs(f, dt, k=20, want_count=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('%-40s %.3f ms  %.1f G pairs/s' % (label, ms, n * n / ms / 1e6), flush=True)
run('default', DBB_K2_DEBUG_SKIP_EPILOGUE='0')
run('skip epilogue', DBB_K2_DEBUG_SKIP_EPILOGUE='1')
run('no epilogue, only warps 0-3 compute (4)', DBB_K2_DEBUG_SKIP_EPILOGUE='4')
run('default again', DBB_K2_DEBUG_SKIP_EPILOGUE='0')
