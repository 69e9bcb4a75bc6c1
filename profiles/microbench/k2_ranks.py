"""Strong scaling of stage 2 emulated on ONE GPU: every rank's share of the pruned sweep (dbb_knn_dev with rank / world)
is run in turn and timed; a step at N GPUs takes as long as its slowest rank (plus the all-gather).
usage: python k2_ranks.py [C3|C4|C2] [scale] [worlds, e.g. 1,2,4,8]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dbb_b200 import device, synthetic
from dbb_b200.data import load_td_dist
from dbb_b200.distributions import PairTables
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C3'
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
worlds = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '1,2,4,8').split(',')]
meta = synthetic.make_config(cfg, scale)
d_ascii, seq_off = device.synth_ascii(meta)
packed = device.pack_ascii(d_ascii, seq_off)
del d_ascii
n = meta.n_scaffolds
freq = torch.zeros((n, 140), dtype=torch.float32, device='cuda')
device.CountPlan(packed).run(None, freq)
td = load_td_dist()
if cfg == 'C3':
    tables = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=synthetic.synthetic_gc_dist(),
                        gcDistPer=80, covDist=synthetic.synthetic_cov_dist(seed=3), covDistPer=80)
    filt = device.FILTER_GC | device.FILTER_COV
else:
    tables = PairTables(meta.length, tdDist=td, tdDistPer=80)
    filt = 0
dt = device.DevicePairTables(tables)
t1 = None
for world in worlds:
    rows = []
    for rank in range(world):
        best = None
        for rep in range(2):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            tm = {}
            e0.record()
            r = device.pairs_pruned(freq, dt, k=20, topk_filter=filt, want_count=True, part=(rank, world) if world > 1 else None, timings=tm)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if best is None or ms < best[0]:
                best = (ms, tm, r['tiles_visited'], r['tiles_second'], r['row_tiles_second'])
        rows.append(best)
    tot = [b[0] for b in rows]
    if t1 is None:
        t1 = max(tot) * world
    print('%s x%.2f n=%d world=%d: rank ms max %.2f mean %.2f (imbalance %.3f) | sweep max %.2f | second max %.2f mean %.2f | tiles/rank %s second tiles %s row tiles %s | efficiency vs first world %.3f'
          % (cfg, scale, n, world, max(tot), np.mean(tot), max(tot) / np.mean(tot), max(b[1]['sweep'] for b in rows),
             max(b[1]['second sweep + merge'] for b in rows), np.mean([b[1]['second sweep + merge'] for b in rows]),
             [b[2] for b in rows][:4], [b[3] for b in rows][:4], [b[4] for b in rows][:4], t1 / (max(tot) * world)), flush=True)
