"""Offline (CPU) estimate for stage 2: how many warp sub-tiles (16 rows x 128 columns) of the EXECUTED 128 x 128 tiles could
stop after 1..4 of the 5 K chunks (28 dims each), because every partial L1 sum already exceeds every bound it is tested
against (partial sums of |a - b| only grow, so the decision is exact).

Emulates the planner of csrc/prune.cu / knn.cu in numpy: 8 equal-count TD-bound classes, saturated-GC projection, sort by
(class, p), tile intervals, tile pair executed iff gap <= max(tmax_r, tmax_c) / 2 + margin.
Row-side bound of a pair = max(TD bound of the row, k-th best distance of the row) (k-th best computed exactly here),
column side = TD bound of the column -- the test of test_row_pair().

usage: python profiles/microbench/k2_early_exit_offline.py [scale of C2, default 0.2] [sampled tiles, default 1500]
Signatures are cached in /tmp/dbb_c2_sig_<scale>.npz.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from dbb_b200 import synthetic                                  # noqa: E402
from dbb_b200.data import load_td_dist                          # noqa: E402
from dbb_b200.distributions import PairTables                   # noqa: E402
from oracle import c_oracle                                     # noqa: E402  (offline analysis, not a product path)

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
TD_PER = 80
T, K, KCH = 128, 20, 28

cache = '/tmp/dbb_c2_sig_%g.npz' % scale
meta = synthetic.make_config('C2', scale)
n = meta.n_scaffolds
if os.path.exists(cache):
    sig = np.load(cache)['sig']
else:
    t0 = time.time()
    sig = np.empty((n, 136), dtype=np.float64)
    step = 512
    for b in range(0, n, step):
        ids = np.arange(b, min(n, b + step))
        flat, off = synthetic.concat(synthetic.generate_sequences(meta, ids))
        c = c_oracle.counts_batch(flat, off).astype(np.float64)
        sig[ids] = c / c.sum(axis=1, keepdims=True)
    np.savez(cache, sig=sig)
    print('signatures of %d scaffolds in %.0f s' % (n, time.time() - t0))
sig = np.nan_to_num(sig).astype(np.float32)

tables = PairTables(meta.length, tdDist=load_td_dist(), tdDistPer=TD_PER)
tdb = np.asarray(tables.td_bound, dtype=np.float32)
_, names = c_oracle.kmer_table()
w = np.array([min(1.0, max(0.0, (sum(ch in 'GC' for ch in nm) - 1) / 2.0)) for nm in names], dtype=np.float32)
p = sig @ w
rank = np.empty(n, dtype=np.int64)
rank[np.argsort(tdb, kind='stable')] = np.arange(n)
cls = (rank * 8) // n
order = np.lexsort((p, cls))
S = torch.from_numpy(sig[order])
tb = tdb[order]
ps = p[order]
n_t = (n + T - 1) // T

# exact k-th best distance per row (blocked cdist)
kth = np.empty(n, dtype=np.float32)
t0 = time.time()
for b in range(0, n, 512):
    d = torch.cdist(S[b:b + 512], S, p=1)
    d[torch.arange(d.shape[0]), torch.arange(b, b + d.shape[0])] = float('inf')
    kth[b:b + 512] = torch.kthvalue(d, K, dim=1).values.numpy()
print('k-th best of %d rows in %.0f s; median k-th best / TD bound = %.2f' % (n, time.time() - t0, np.median(kth / tb)))

pad = n_t * T - n
pmin = np.pad(ps, (0, pad), constant_values=np.inf).reshape(n_t, T).min(1)
pmax = np.pad(ps, (0, pad), constant_values=-np.inf).reshape(n_t, T).max(1)
tmax = np.pad(tb, (0, pad), constant_values=-np.inf).reshape(n_t, T).max(1)
gap = np.maximum(0, np.maximum(pmin[:, None] - pmax[None, :], pmin[None, :] - pmax[:, None]))
execd = 2 * gap <= np.maximum(tmax[:, None], tmax[None, :]) + 1e-5
print('n = %d, %d x %d tiles, executed by the TD plan: %.1f %%' % (n, n_t, n_t, 100 * execd.mean()))

rng = np.random.RandomState(0)
rt, ct = np.nonzero(execd)
pick = rng.choice(rt.size, min(n_sample, rt.size), replace=False)
u_row = np.maximum(tb, kth)

for label, perm in (('natural column order', np.arange(136)),
                    ('columns by decreasing variance', np.argsort(-sig.var(axis=0), kind='stable'))):
    Sp = torch.zeros((n, 140))
    Sp[:, :136] = S[:, perm]
    first_warp = np.zeros(6)      # [c] = warp sub-tiles whose first skippable point is after chunk c+1 (index 5: never)
    first_warp_cons = np.zeros(6)
    first_cta = np.zeros(6)
    for r, c in zip(rt[pick], ct[pick]):
        A = Sp[r * T:(r + 1) * T]
        B = Sp[c * T:(c + 1) * T]
        ur = torch.from_numpy(u_row[r * T:r * T + A.shape[0]])
        tc = torch.from_numpy(tb[c * T:c * T + B.shape[0]])
        bound = torch.maximum(ur[:, None], tc[None, :])                        # pair passes the hot test iff d <= bound
        part = torch.zeros((A.shape[0], B.shape[0]))
        ok_w = np.full((5, 8), False)
        ok_wc = np.full((5, 8), False)
        for kc in range(5):
            part = part + (A[:, None, kc * KCH:(kc + 1) * KCH] - B[None, :, kc * KCH:(kc + 1) * KCH]).abs().sum(-1)
            slack = part - bound
            for wq in range(8):
                rows = slice(16 * wq, min(16 * wq + 16, A.shape[0]))
                if rows.start >= A.shape[0]:
                    ok_w[kc, wq] = ok_wc[kc, wq] = True
                    continue
                ok_w[kc, wq] = bool((slack[rows] > 0).all())
                ok_wc[kc, wq] = bool(part[rows].min() > max(float(ur[rows].max()), float(tc.max())))
        for wq in range(8):
            f = np.nonzero(ok_w[:4, wq])[0]
            first_warp[f[0] if f.size else 5] += 1
            f = np.nonzero(ok_wc[:4, wq])[0]
            first_warp_cons[f[0] if f.size else 5] += 1
        f = np.nonzero(ok_w[:4].all(axis=1))[0]
        first_cta[f[0] if f.size else 5] += 1
    for nm, fw in (('per warp, exact per-pair bounds', first_warp), ('per warp, one bound (max row, max column)', first_warp_cons),
                   ('whole CTA, exact per-pair bounds', first_cta)):
        fw = fw / fw.sum()
        saved = sum(fw[c] * (4 - c) / 5.0 for c in range(4))
        print('%-32s %-44s stop after chunk 1 / 2 / 3 / 4 / never: %s  -> K-loop work saved %.1f %%'
              % (label, nm, ' '.join('%.3f' % v for v in (fw[0], fw[1], fw[2], fw[3], fw[5])), 100 * saved))
