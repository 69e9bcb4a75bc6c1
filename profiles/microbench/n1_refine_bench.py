"""N1 measurement: dbb_refine_dev on device-resident inputs, U unbinned contigs x B cores (FP64)."""
import ctypes, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dbb_b200 import _lib, synthetic
from dbb_b200.data import load_td_dist
U, B = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.RandomState(0)
td, gcd, cvd = load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=1)
cent = rng.dirichlet(np.ones(136) * 5, B)
c_gc = rng.uniform(0.3, 0.7, B); c_cov = rng.lognormal(3, 1, B)
own = rng.randint(0, B, U)
u_sig = cent[own] * rng.gamma(200, 1 / 200.0, (U, 136)); u_sig /= u_sig.sum(1, keepdims=True)
u_len = rng.randint(2500, 50000, U)
u_gc = c_gc[own] + rng.normal(0, 0.01, U); u_cov = c_cov[own] * (1 + rng.normal(0, 0.05, U))
from dbb_b200.distributions import Distributions, nearest_index, _sorted_keys
d = Distributions(gcd, 99, td, 99, cvd, 99)
tl = _sorted_keys(td); means = _sorted_keys(gcd); gl = _sorted_keys(gcd[means[0]]); cl = _sorted_keys(cvd)
T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
t = dict(u_sig=T(u_sig, np.float64), c_sig=T(cent, np.float64), u_gc=T(u_gc, np.float64), u_cov=T(u_cov, np.float64), c_gc=T(c_gc, np.float64),
         c_cov=T(c_cov, np.float64), u_tdb=T(np.array([td[l][d.tdKey] for l in tl])[nearest_index(tl, u_len)], np.float64),
         u_gcl=T(nearest_index(gl, u_len), np.int32), u_cl=T(nearest_index(cl, u_len), np.int32), c_gm=T(nearest_index(means, c_gc), np.int32),
         glo=T([[gcd[m][l][d.gcLowerBoundKey] for l in gl] for m in means], np.float64), ghi=T([[gcd[m][l][d.gcUpperBoundKey] for l in gl] for m in means], np.float64),
         clo=T([cvd[l][d.covLowerBoundKey] for l in cl], np.float64), chi=T([cvd[l][d.covUpperBoundKey] for l in cl], np.float64))
closest = torch.empty((U, 3), dtype=torch.int32, device='cuda'); assigned = torch.empty(U, dtype=torch.int32, device='cuda')
p = lambda x: ctypes.c_void_p(x.data_ptr())
inp = _lib.RefineInput(U, B, p(t['u_sig']), p(t['c_sig']), p(t['u_gc']), p(t['u_cov']), p(t['c_gc']), p(t['c_cov']), p(t['u_tdb']), p(t['u_gcl']),
                       p(t['u_cl']), p(t['c_gm']), p(t['glo']), p(t['ghi']), len(means), len(gl), p(t['clo']), p(t['chi']), len(cl))
for want_within in (False, True):
    within = torch.empty((U, B), dtype=torch.uint8, device='cuda') if want_within else None
    out = _lib.RefineOutput(p(closest), None, p(assigned), p(within) if want_within else None)
    L = _lib.lib()
    for _ in range(2): _lib.check(L.dbb_refine_dev(ctypes.byref(inp), ctypes.byref(out), None))
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _lib.check(L.dbb_refine_dev(ctypes.byref(inp), ctypes.byref(out), None))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('refine U=%d B=%d within=%s: %.3f ms  %.2f G (contig,core) pairs/s  assigned %.1f%%' % (U, B, want_within, ms, U * B / ms / 1e6, 100.0 * (assigned >= 0).float().mean().item()), flush=True)
