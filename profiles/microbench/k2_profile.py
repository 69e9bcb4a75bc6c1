"""Small driver for ncu captures of pairs_kernel: python k2_profile.py [n_scaffolds] [td|all]
(td = k-NN + TD count as in C2; all = TD + GC + coverage with filtered top-k as in C3).  Synthetic signatures around
genome centroids; no counting stage."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dbb_b200 import device, synthetic
from dbb_b200.distributions import PairTables
from dbb_b200.data import load_td_dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
mode = sys.argv[2] if len(sys.argv) > 2 else 'td'
rng = np.random.RandomState(0)
meta = synthetic.make_metagenome(seed=2, n_scaffolds=n, min_len=1000, max_len=100000, n_genomes=200)
cent = rng.dirichlet(np.ones(136) * 5, 200)
sig = (cent[meta.genome] * rng.gamma(50, 1 / 50.0, (n, 136))); sig /= sig.sum(1, keepdims=True)
f = torch.zeros((n, 140), dtype=torch.float32, device='cuda'); f[:, :136] = torch.from_numpy(sig.astype(np.float32)).cuda()
if mode == 'all':
    t = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=load_td_dist(), tdDistPer=80,
                   gcDist=synthetic.synthetic_gc_dist(), gcDistPer=80, covDist=synthetic.synthetic_cov_dist(seed=3), covDistPer=80)
    filt = device.FILTER_GC | device.FILTER_COV
else:
    t = PairTables(meta.length, tdDist=load_td_dist(), tdDistPer=80)
    filt = 0
dt = device.DevicePairTables(t)
for _ in range(3):
    device.pairs(f, dt, k=20, want_count=True, topk_filter=filt)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    device.pairs(f, dt, k=20, want_count=True, topk_filter=filt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print('%s n=%d %.3f ms  %.1f G pairs/s' % (mode, n, ms, n * n / ms / 1e6), flush=True)
