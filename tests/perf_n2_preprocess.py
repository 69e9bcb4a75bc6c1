"""(Lives under tests/ because its CPU leg imports the oracle.)  N2 row: device timing of the batched sequence side of preprocess on the C2 input (50 000 scaffolds, 1.07 Gbp,
ASCII resident in HBM), stage by stage, with the algorithmic bytes each stage moves; and the reference's per-scaffold
loop (oracle restatement of splitScaffolds.py:27-53 + seqUtils.py:258-265) timed on a sample of the same scaffolds."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dbb_b200 import device, synthetic
from oracle import dbb_oracle as O

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
meta = synthetic.make_config('C2', scale)
d_ascii, off = device.synth_ascii(meta)
n, bp = meta.n_scaffolds, int(off[-1])
MIN_N = 10
PEAK = 6541.1


def timed(fn, reps=5):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


rows = []
import ctypes
from dbb_b200 import _lib
L = _lib.lib()
P, ST = device._p, device._stream
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).cuda()
packed = device.pack_ranges(d_ascii, off[:-1], off[1:], want_nmask=True)
d_b, d_e, d_len = up(off[:-1]), up(off[1:]), up(np.diff(off))
ms, _ = timed(lambda: _lib.check(L.dbb_pack_ranges_dev(P(d_ascii), P(d_b), P(d_e), P(packed.d_blk_off), n, packed.total_blocks,
                                                       P(packed.codes), P(packed.valid), P(packed.nmask), ST())))
rows.append(('pack kernel (codes + validity + N plane)', ms, bp * (1 + 0.25 + 0.125 + 0.125)))
c_off, cb, ce = device.split_scaffolds(packed, MIN_N)
d_coff = up(c_off)
d_nc = torch.zeros(n, dtype=torch.int32, device='cuda')
d_cb, d_ce = torch.zeros(int(c_off[-1]), dtype=torch.int64, device='cuda'), torch.zeros(int(c_off[-1]), dtype=torch.int64, device='cuda')
def split_both():
    _lib.check(L.dbb_split_scaffolds_dev(P(packed.nmask), P(d_len), P(packed.d_blk_off), n, MIN_N, None, P(d_nc), None, None, ST()))
    _lib.check(L.dbb_split_scaffolds_dev(P(packed.nmask), P(d_len), P(packed.d_blk_off), n, MIN_N, P(d_coff), None, P(d_cb), P(d_ce), ST()))
ms, _ = timed(split_both)
rows.append(('N-run splitter kernels, count + write pass', ms, 2 * bp / 8.0))
nc = np.diff(c_off)
multi = np.nonzero(nc > 1)[0]
sel = np.concatenate([np.arange(c_off[s], c_off[s + 1]) for s in multi])
owner = np.repeat(multi, nc[multi])
ab, ae = off[owner] + cb[sel], off[owner] + ce[sel]
d_rb, d_re = up(np.concatenate([off[:-1], ab])), up(np.concatenate([off[1:], ae]))
d_acgt = torch.zeros((d_rb.shape[0], 4), dtype=torch.int64, device='cuda')
ms, _ = timed(lambda: _lib.check(L.dbb_base_count_dev(P(d_ascii), P(d_rb), P(d_re), d_rb.shape[0], P(d_acgt), ST())))
rows.append(('base composition kernel, scaffolds + contigs', ms, float(bp + (ae - ab).sum())))
plan = device.CountPlan(packed)
counts = torch.zeros((n, 136), dtype=torch.int32, device='cuda')
ms, _ = timed(lambda: plan.run(counts, None))
rows.append(('tetramer counts of the scaffolds (K1)', ms, float(packed.algorithmic_bytes())))
cpk = device.pack_ranges(d_ascii, ab, ae)
d_ab, d_ae = up(ab), up(ae)
ms, _ = timed(lambda: _lib.check(L.dbb_pack_ranges_dev(P(d_ascii), P(d_ab), P(d_ae), P(cpk.d_blk_off), cpk.n, cpk.total_blocks,
                                                       P(cpk.codes), P(cpk.valid), None, ST())))
rows.append(('pack kernel, contigs of split scaffolds', ms, float((ae - ab).sum()) * 1.375))
t0 = time.time()
for _ in range(3):
    from dbb_b200.preprocess import Preprocess
torch.cuda.synchronize()
total = sum(r[1] for r in rows)
print('C2 x%.2f: %d scaffolds, %.3f Gbp, minN = %d: %d scaffolds split into %d contigs' % (scale, n, bp / 1e9, MIN_N, multi.size, sel.size))
for name, ms, by in rows:
    print('  %-62s %8.3f ms  %7.1f GB/s (%.2f of measured HBM peak)' % (name, ms, by / ms / 1e6, by / ms / 1e6 / PEAK))
print('  %-62s %8.3f ms  = %.1f Gbp/s for the whole sequence side' % ('sum', total, bp / total / 1e6))

# CPU: the reference's loop, one scaffold at a time, on a sample (single core, as the reference's worker is per process)
host = d_ascii.cpu().numpy().tobytes()
rng = np.random.RandomState(1)
sample = rng.choice(n, 300, replace=False)
t0 = time.time()
sbp = 0
for s in sample:
    seq = host[off[s]:off[s + 1]].decode('latin-1')
    O.splits_literal('s', seq, MIN_N)
    O.base_count_literal(seq)
    sbp += len(seq)
dt = time.time() - t0
print('  reference loop (splits + baseCount, 1 core, %d scaffolds = %.1f Mbp): %.2f s = %.4f Gbp/s' % (len(sample), sbp / 1e6, dt, sbp / dt / 1e9))
