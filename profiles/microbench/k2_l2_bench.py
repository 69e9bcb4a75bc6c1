"""Timing of the Euclidean k-NN mode beside the Manhattan one (same synthetic signatures as k2_profile.py)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dbb_b200 import device, synthetic
from dbb_b200.distributions import PairTables
from dbb_b200.data import load_td_dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.RandomState(0)
meta = synthetic.make_metagenome(seed=2, n_scaffolds=n, min_len=1000, max_len=100000, n_genomes=200)
cent = rng.dirichlet(np.ones(136) * 5, 200)
sig = (cent[meta.genome] * rng.gamma(50, 1 / 50.0, (n, 136))); sig /= sig.sum(1, keepdims=True)
f = torch.zeros((n, 140), dtype=torch.float32, device='cuda'); f[:, :136] = torch.from_numpy(sig.astype(np.float32)).cuda()
dt = device.DevicePairTables(PairTables(meta.length, tdDist=load_td_dist(), tdDistPer=80))
dt0 = device.DevicePairTables(PairTables(meta.length))
for label, tab, metric, cnt in (('l1 knn + TD count', dt, 'l1', True), ('l2 knn + bounded count', dt, 'l2', True),
                                ('l1 knn only', dt0, 'l1', False), ('l2 knn only', dt0, 'l2', False)):
    for _ in range(3):
        device.pairs(f, tab, k=20, want_count=cnt, metric=metric)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        device.pairs(f, tab, k=20, want_count=cnt, metric=metric)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('%-26s n=%d %.3f ms  %.1f G pairs/s' % (label, n, ms, n * n / ms / 1e6), flush=True)
