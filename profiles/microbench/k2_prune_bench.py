"""pairs vs pairs_pruned on the C2 workload (real signatures from the counting stage): python k2_prune_bench.py [C2|C3|C4] [scale]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from dbb_b200 import device, synthetic, _lib
from dbb_b200.distributions import PairTables
from dbb_b200.data import load_td_dist
name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
meta, kw = bench.workload(1, scale, name)
d_ascii, off = device.synth_ascii(meta)
packed = device.pack_ascii(d_ascii, off); del d_ascii
plan = device.CountPlan(packed)
n = meta.n_scaffolds
freq = torch.zeros((n, _lib.SIG_STRIDE), dtype=torch.float32, device='cuda')
plan.run(None, freq)
if name == 'C3':
    t = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=load_td_dist(), tdDistPer=80,
                   gcDist=synthetic.synthetic_gc_dist(), gcDistPer=80, covDist=synthetic.synthetic_cov_dist(seed=3), covDistPer=80)
    filt = device.FILTER_GC | device.FILTER_COV
else:
    t = PairTables(meta.length, tdDist=load_td_dist(), tdDistPer=80)
    filt = 0
dt = device.DevicePairTables(t)


REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 3


def timed(fn, reps=REPS):
    for _ in range(1 if REPS == 1 else 2): r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

ms_f, full = timed(lambda: device.pairs(freq, dt, k=20, topk_filter=filt, want_count=True))
ncls = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ms_p, pr = timed(lambda: device.pairs_pruned(freq, dt, k=20, topk_filter=filt, want_count=True, n_classes=ncls))
tm = {}
device.pairs_pruned(freq, dt, k=20, topk_filter=filt, want_count=True, n_classes=ncls, timings=tm)
print('classes', ncls, {k_: (round(v, 3) if isinstance(v, float) else v) for k_, v in tm.items()})
got = device.unsort_rows(pr, n)
same = all(torch.equal(got[k_], full[k_]) for k_ in ('idx', 'dist', 'count', 'neighbour_bp'))
print('%s x%.2f n=%d: full sweep %.3f ms (%.1f G pairs/s), pruned %.3f ms (%.1f G pairs/s), tiles visited %d of %d (%.1f %%), identical results: %s'
      % (name, scale, n, ms_f, n * n / ms_f / 1e6, ms_p, n * n / ms_p / 1e6, pr['tiles_visited'], pr['tiles_total'],
         100.0 * pr['tiles_visited'] / pr['tiles_total'], same))
