"""Parity at BASELINE.json sizes through the device-resident pipeline (synth -> pack -> count -> pairs):
exact counts for a scaffold sample, size-independent properties for all scaffolds, and oracle rows for a
row sample of the all-pairs stage with all three predicates enabled."""
import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _pipeline(cfg, scale):
    import torch
    from dbb_b200 import device, synthetic
    meta = synthetic.make_config(cfg, scale)
    d_ascii, seq_off = device.synth_ascii(meta)
    packed = device.pack_ascii(d_ascii, seq_off)
    plan = device.CountPlan(packed)
    counts = torch.zeros((meta.n_scaffolds, 136), dtype=torch.int32, device='cuda')
    freq = torch.zeros((meta.n_scaffolds, 140), dtype=torch.float32, device='cuda')
    plan.run(counts, freq)
    torch.cuda.synchronize()
    return meta, d_ascii, seq_off, counts, freq


def test_c2_full_size_counts_and_properties():
    import torch
    meta, d_ascii, seq_off, counts, freq = _pipeline('C2', 1.0)
    n = meta.n_scaffolds
    assert n == 50000 and 1.0e9 < meta.total_bp < 1.15e9
    c = counts.cpu().numpy().astype(np.int64)
    f = freq.cpu().numpy()
    # property: every scaffold without N / lowercase has exactly L - 3 windows
    clean = (meta.n_start < 0) & (meta.lc_start < 0)
    assert np.array_equal(c.sum(axis=1)[clean], meta.length[clean] - 3)
    assert (c.sum(axis=1)[~clean] < meta.length[~clean] - 3).all()
    # property: frequencies are counts / total within 1 ulp, padding columns are zero
    tot = c.sum(axis=1, keepdims=True).astype(np.float64)
    f64 = c / tot
    assert (np.abs(f[:, :136] - f64) <= util.FREQ_RTOL * f64).all() and (f[:, 136:] == 0).all()
    # exact counts of ALL 50 000 scaffolds against the C restatement of the oracle
    from oracle import c_oracle as C
    assert np.array_equal(c, C.counts_batch(d_ascii.cpu().numpy(), seq_off))
    # and against the numpy oracle: a seeded sample + every scaffold with N or lowercase in the first 5000
    rng = np.random.RandomState(0)
    sample = set(rng.choice(n, 150, replace=False).tolist()) | set(np.nonzero(~clean[:5000])[0].tolist())
    sample |= {int(np.argmax(meta.length)), int(np.argmin(meta.length))}
    for i in sorted(sample):
        s = d_ascii[seq_off[i]:seq_off[i + 1]].cpu().numpy()
        assert np.array_equal(c[i], O.seq_counts_np(s)), 'scaffold %d (len %d)' % (i, meta.length[i])


def test_c3_like_pairs_with_all_filters_row_sample():
    """20 000 scaffolds of the C3 law, TD + GC + coverage predicates, k = 20 restricted by GC and coverage."""
    import torch
    from dbb_b200 import device, synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.distributions import PairTables
    meta, d_ascii, seq_off, counts, freq = _pipeline('C3', 0.08)
    n = meta.n_scaffolds
    td, gcd, cvd = load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=3)
    for fixed in (None, 10000):
        tables = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd,
                            gcDistPer=80, covDist=cvd, covDistPer=80, fixedSeqLen=fixed)
        res = device.pairs(freq, device.DevicePairTables(tables), k=20, topk_filter=device.FILTER_GC | device.FILTER_COV,
                           want_count=True)
        torch.cuda.synchronize()
        rows = np.random.RandomState(1).choice(n, 96, replace=False).tolist()
        sig32 = freq[:, :136].cpu().numpy().astype(np.float64)
        ot = O.DistTables(gcd, 80, td, 80, cvd, 80)
        with np.errstate(invalid='ignore'):
            d64 = O.l1_rows_np(sig32, rows)
            o_td = O.td_rows_np(sig32, meta.length, rows, ot, fixed)
        o_gc = O.gc_rows_np(meta.length, meta.gc_obs, rows, ot, fixed)
        o_cov = O.cov_rows_np(meta.length, meta.coverage, rows, ot, fixed)
        idx = res['idx'].cpu().numpy()[rows]; dist = res['dist'].cpu().numpy()[rows].astype(np.float64)
        util.check_topk(idx, dist, d64, rows, 20, allowed=o_gc * o_cov)
        allm = (o_td * o_gc * o_cov).astype(bool)
        bound = O.td_bound_rows_np(meta.length, rows, ot, fixed)
        near = (np.abs(d64 - bound) <= util.DIST_ATOL) & (o_gc * o_cov).astype(bool)
        cnt = res['count'].cpu().numpy()[rows]; bp = res['neighbour_bp'].cpu().numpy()[rows]
        assert (np.abs(cnt - allm.sum(axis=1)) <= near.sum(axis=1)).all()
        exact = near.sum(axis=1) == 0
        assert np.array_equal(cnt[exact], allm.sum(axis=1)[exact])
        assert np.array_equal(bp[exact], (allm * meta.length[None, :]).sum(axis=1)[exact])


def test_c5_like_skewed_lengths_counting():
    """Short 1-2 kb contigs plus multi-megabase scaffolds (the chunked S = 4 class and the finalize kernel)."""
    import torch
    from dbb_b200 import device, synthetic
    meta = synthetic.make_metagenome(seed=5, n_scaffolds=20000, min_len=1000, max_len=2000, n_genomes=50, law='uniform',
                                     extra_lengths=(5000000, 3000001, 1048576 + 64))
    d_ascii, seq_off = device.synth_ascii(meta)
    packed = device.pack_ascii(d_ascii, seq_off)
    plan = device.CountPlan(packed)
    counts = torch.zeros((meta.n_scaffolds, 136), dtype=torch.int32, device='cuda')
    plan.run(counts, None)
    torch.cuda.synchronize()
    c = counts.cpu().numpy().astype(np.int64)
    clean = (meta.n_start < 0) & (meta.lc_start < 0)
    assert np.array_equal(c.sum(axis=1)[clean], meta.length[clean] - 3)
    for i in list(range(meta.n_scaffolds - 3, meta.n_scaffolds)) + [0, 1, 2, 777]:
        s = d_ascii[seq_off[i]:seq_off[i + 1]].cpu().numpy()
        assert np.array_equal(c[i], O.seq_counts_np(s)), i


def test_c1_whole_config_against_oracle():
    """BASELINE.json configs[0] (2 000 scaffolds, 1-50 kb): every scaffold's counts, the full TD / GC / coverage
    masks through the drop-in classes' kernel path and k = 20 for every row, against the oracle."""
    from dbb_b200 import _pairs, synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.distributions import PairTables
    from dbb_b200.genomicSignatures import GenomicSignatures
    meta = synthetic.make_config('C1')
    seqs = synthetic.generate_sequences(meta)
    counts, freq = GenomicSignatures(4, 1).signatures(seqs)
    for i, s in enumerate(seqs):
        assert np.array_equal(counts[i].astype(np.int64), O.seq_counts_np(s)), i
    lit = np.random.RandomState(3).choice(meta.n_scaffolds, 12, replace=False)
    for i in lit:                                                   # the literal (authoritative) form on a sample
        assert O.seq_counts_literal(seqs[i].tobytes().decode('ascii')) == counts[i].tolist()
    n = meta.n_scaffolds
    td, gcd, cvd = load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=1)
    sig64 = counts.astype(np.float64) / counts.sum(axis=1, keepdims=True)
    tables = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd,
                        gcDistPer=80, covDist=cvd, covDistPer=80)
    res = _pairs.run_pairs_host(sig64, tables, k=20, want_count=True, want_td_mask=True, want_gc_mask=True, want_cov_mask=True)
    rows = list(range(n))
    ot = O.DistTables(gcd, 80, td, 80, cvd, 80)
    s32 = sig64.astype(np.float32).astype(np.float64)
    d64 = O.l1_rows_np(s32, rows)
    assert np.array_equal(util.unpack_mask(res['gc_mask'], n), O.gc_rows_np(meta.length, meta.gc_obs, rows, ot, None))
    assert np.array_equal(util.unpack_mask(res['cov_mask'], n), O.cov_rows_np(meta.length, meta.coverage, rows, ot, None))
    flips = util.check_mask_against_oracle(util.unpack_mask(res['td_mask'], n), O.td_rows_np(s32, meta.length, rows, ot, None),
                                           d64, O.td_bound_rows_np(meta.length, rows, ot, None))
    assert flips <= 4
    util.check_topk(res['idx'], res['dist'].astype(np.float64), d64, rows, 20)
    # literal per-pair loop of tdNeighbours.py:41-51 on two rows
    cs = [O.Contig('c', int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), s32[i]) for i in range(n)]
    dT = O.Distributions(None, None, td, 80, None, None)
    for i in (0, 1234):
        lit_row = O.td_row_literal(cs[i], cs, dT, None)
        assert np.array_equal(lit_row, O.td_rows_np(s32, meta.length, [i], ot, None)[0])
