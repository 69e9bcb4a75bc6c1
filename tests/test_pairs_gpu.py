"""Parity of stage 2 (all-pairs distance, neighbour predicates, fused top-k) on the GPU against the oracle
and the golden rows produced by the reference's own __workerThread loops.  GC / coverage masks: exact.
TD mask: exact outside |d64 - bound| <= 2e-6.  Distances: |d32 - d64| <= 2e-6."""
import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util

pytestmark = pytest.mark.gpu

CASES = (('fixedNone_p80', None, 80), ('fixed10000_p80', 10000, 80), ('fixedNone_p99', None, 99))


def _golden_contigs():
    nb = util.golden_npz('ref_neighbours.npz')
    cs = [O.Contig('c%d' % i, int(nb['length'][i]), float(nb['GC'][i]), float(nb['coverage'][i]), nb['tetraSig'][i])
          for i in range(len(nb['length']))]
    return nb, cs


def _dists():
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    return load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=7)


@pytest.mark.parametrize('tag,fixed,per', CASES)
def test_golden_rows_through_dropin_classes(tag, fixed, per):
    from dbb_b200.tdNeighbours import TdNeighbours
    from dbb_b200.gcNeighbours import GcNeighbours
    from dbb_b200.coverageNeighbours import CoverageNeighbours
    nb, cs = _golden_contigs()
    td, gcd, cvd = _dists()
    assert TdNeighbours(fixed).run(cs, td, per, 4) is None
    GcNeighbours(fixed).run(cs, gcd, per, 4)
    CoverageNeighbours(fixed).run(cs, cvd, per, 4)
    got_td = np.array([c.tdNeighbours for c in cs]); got_gc = np.array([c.gcNeighbours for c in cs])
    got_cov = np.array([c.covNeighbours for c in cs])
    assert got_td.dtype == np.float64 and got_td.shape == nb['td_' + tag].shape
    assert np.array_equal(got_gc, nb['gc_' + tag])
    assert np.array_equal(got_cov, nb['cov_' + tag])
    tables = O.DistTables(tdDist=td, tdDistPer=per)
    bound = O.td_bound_rows_np(nb['length'], list(range(len(cs))), tables, fixed)
    with np.errstate(invalid='ignore'):
        util.check_mask_against_oracle(got_td, nb['td_' + tag], nb['distance'], bound)
    # the expressions greedy.py:409-414 applies to the attributes
    i = 3
    assert np.array_equal(np.where(cs[i].tdNeighbours == 1)[0], np.where(got_td[i] == 1)[0])


def _metagenome(seed, n, gmax=30000):
    from dbb_b200 import synthetic
    from dbb_b200.genomicSignatures import GenomicSignatures
    meta = synthetic.make_metagenome(seed=seed, n_scaffolds=n, min_len=1000, max_len=gmax, n_genomes=6)
    seqs = synthetic.generate_sequences(meta)
    counts, _ = GenomicSignatures(4, 1).signatures(seqs, want_freq=False)
    sig64 = counts.astype(np.float64) / counts.sum(axis=1, keepdims=True)
    return meta, sig64


def test_masks_counts_topk_vs_oracle_all_predicates():
    from dbb_b200 import _pairs
    from dbb_b200.distributions import PairTables
    meta, sig64 = _metagenome(21, 700)
    sig64[17] = np.nan                                    # NaN signature: neighbour of nobody
    meta.coverage[5] = 0.0
    td, gcd, cvd = _dists()
    n = meta.n_scaffolds
    rows = list(range(n))
    for fixed in (None, 10000):
        tables = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd,
                            gcDistPer=80, covDist=cvd, covDistPer=80, fixedSeqLen=fixed)
        res = _pairs.run_pairs_host(sig64, tables, k=20, topk_filter=_pairs.FILTER_GC | _pairs.FILTER_COV,
                                    want_count=True, want_td_mask=True, want_gc_mask=True, want_cov_mask=True)
        ot = O.DistTables(gcd, 80, td, 80, cvd, 80)
        with np.errstate(invalid='ignore'):
            d64 = O.l1_rows_np(sig64.astype(np.float32).astype(np.float64), rows)
            o_td = O.td_rows_np(sig64.astype(np.float32).astype(np.float64), meta.length, rows, ot, fixed)
        o_gc = O.gc_rows_np(meta.length, meta.gc_obs, rows, ot, fixed)
        o_cov = O.cov_rows_np(meta.length, meta.coverage, rows, ot, fixed)
        g_td, g_gc, g_cov = (util.unpack_mask(res[k_], n) for k_ in ('td_mask', 'gc_mask', 'cov_mask'))
        assert np.array_equal(g_gc, o_gc) and np.array_equal(g_cov, o_cov)
        bound = O.td_bound_rows_np(meta.length, rows, ot, fixed)
        with np.errstate(invalid='ignore'):
            flips = util.check_mask_against_oracle(g_td, o_td, d64, bound)
        allm = (g_td & g_gc & g_cov).astype(bool)
        assert np.array_equal(res['count'], allm.sum(axis=1))
        assert np.array_equal(res['neighbour_bp'], (allm * meta.length[None, :]).sum(axis=1))
        util.check_topk(res['idx'], res['dist'].astype(np.float64), d64, rows, 20, allowed=(o_gc * o_cov))
        assert flips <= 2


def test_plain_knn_and_row_block_invariance():
    """k-NN without predicates; results for a row block are identical to the same rows of the full run
    (what makes the multi-GPU row sharding exact)."""
    from dbb_b200 import _pairs
    from dbb_b200.distributions import PairTables
    meta, sig64 = _metagenome(22, 1000, gmax=20000)
    n = meta.n_scaffolds
    td = _dists()[0]
    tables = PairTables(meta.length, tdDist=td, tdDistPer=80)
    full = _pairs.run_pairs_host(sig64, tables, k=20, want_count=True)
    d64 = O.l1_rows_np(sig64.astype(np.float32).astype(np.float64), list(range(n)))
    util.check_topk(full['idx'], full['dist'].astype(np.float64), d64, list(range(n)), 20)
    for r0, r1 in ((0, 130), (130, 515), (515, n)):
        part = _pairs.run_pairs_host(sig64, tables, k=20, want_count=True, row0=r0, row1=r1)
        for key in ('idx', 'dist', 'count', 'neighbour_bp'):
            assert np.array_equal(part[key], full[key][r0:r1]), key
    for k in (1, 5, 32):
        r = _pairs.run_pairs_host(sig64, tables, k=k)
        util.check_topk(r['idx'], r['dist'].astype(np.float64), d64, list(range(n)), k)


def test_small_and_degenerate_inputs():
    from dbb_b200 import _pairs
    from dbb_b200.distributions import PairTables
    td = _dists()[0]
    rng = np.random.RandomState(9)
    for n in (1, 2, 19, 128, 129):
        sig = rng.dirichlet(np.ones(136), n)
        length = rng.randint(1000, 50000, n)
        tables = PairTables(length, tdDist=td, tdDistPer=80)
        r = _pairs.run_pairs_host(sig, tables, k=20, want_count=True, want_td_mask=True)
        d64 = O.l1_rows_np(sig.astype(np.float32).astype(np.float64), list(range(n)))
        util.check_topk(r['idx'], r['dist'].astype(np.float64), d64, list(range(n)), 20)
        m = util.unpack_mask(r['td_mask'], n)
        assert (np.diag(m) == 1).all()                     # i == j is always a TD neighbour (0 < bound)
        assert np.array_equal(r['count'], m.sum(axis=1))


def test_column_split_items_merge_to_identical_results(monkeypatch):
    """Row tiles of the last partial wave are split into column segments and merged (csrc/pairs.cu
    plan_items / merge_kernel); forcing a 3-way split on a small input must not change any output."""
    from dbb_b200 import _pairs
    from dbb_b200.distributions import PairTables
    meta, sig64 = _metagenome(23, 1500, gmax=12000)
    td, gcd, cvd = _dists()
    tables = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd,
                        gcDistPer=80, covDist=cvd, covDistPer=80)
    kw = dict(k=20, topk_filter=_pairs.FILTER_GC, want_count=True, want_td_mask=True)
    monkeypatch.setenv('DBB_K2_SEG', '1')
    whole = _pairs.run_pairs_host(sig64, tables, **kw)
    for seg in ('2', '3', '5'):
        monkeypatch.setenv('DBB_K2_SEG', seg)
        split = _pairs.run_pairs_host(sig64, tables, **kw)
        for key in whole:
            assert np.array_equal(whole[key], split[key]), (seg, key)


def test_refine_closest_cores_vs_reference_rows_and_oracle():
    """N1: dbb_b200.refineBins against the reference's own __workerThread rows (golden) and, on a larger random
    instance, against the oracle's literal loop."""
    from tests.test_oracle_golden import _refine_setup
    from dbb_b200.refineBins import RefineBins, coreStatistics
    nb, cs, (td, gcd, cvd) = _refine_setup()
    cores = coreStatistics(cs[:30])
    for n, k in enumerate(sorted(cores)):
        assert cores[k].length == nb['refine_core_len'][n] and cores[k].GC == nb['refine_core_gc'][n]
        assert np.array_equal(cores[k].tetraSig, nb['refine_core_sig'][n])
    unb = cs[30:]
    rows = RefineBins().refine(unb, cores, gcd, 99, td, 99, cvd, 99)
    got = np.array([[r[0], r[1], r[2], int(r[3]), int(r[4]), int(r[5])] for r in rows])
    assert np.array_equal(got, nb['refine_rows'])
    assert [c.binId for c in unb] == nb['refine_rows'][:, 1].tolist()
    # larger instance: 3 000 unbinned x 57 cores, with the per-pair flag matrix
    from dbb_b200 import synthetic
    from dbb_b200.genomicSignatures import GenomicSignatures
    meta = synthetic.make_metagenome(seed=31, n_scaffolds=3600, min_len=1000, max_len=20000, n_genomes=57)
    counts, _ = GenomicSignatures(4, 1).signatures(synthetic.generate_sequences(meta), want_freq=False)
    sig = counts / counts.sum(axis=1, keepdims=True)
    allc = [O.Contig('c%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), sig[i]) for i in range(3600)]
    for i, c in enumerate(allc):
        c.binId = int(meta.genome[i])
    allc[5].coverage = 0.0
    cores = coreStatistics(allc[:600])
    ocores = O.core_statistics_literal(allc[:600])
    unb = allc[600:]
    closest, mind, assigned, within = RefineBins().closestCores(unb, cores, gcd, 99, td, 99, cvd, 99, want_within=True)
    dist = O.Distributions(gcd, 99, td, 99, cvd, 99)
    order = sorted(ocores)
    n_assigned = 0
    for i in list(range(0, 3000, 7)) + [2999]:
        cb, md, newBin, bA, bF, bI = O.refine_worker_literal(unb[i], ocores, dist)
        exp = [-1 if b is None else order.index(b) for b in cb]
        assert closest[i].tolist() == exp, i
        assert np.allclose(mind[i], md, rtol=0, atol=1e-12)
        assert (assigned[i] >= 0) == bA and (not bA or order[assigned[i]] == newBin)
        n_assigned += bA
        for b, k in enumerate(order):
            f = within[i, b]
            assert bool(f & 1) == dist.withinDistGC(unb[i], ocores[k]) and bool(f & 2) == dist.withinDistCov(unb[i], ocores[k])
            assert bool(f & 4) == dist.witinDistTD(O.distance(unb[i].tetraSig, np.asarray(ocores[k].tetraSig)), unb[i])
    assert n_assigned > 100


def test_queue_path_equals_mask_path_with_coverages_on_the_bounds():
    """The count / top-k path (queued pairs, two-pass drain, FP32-screened coverage predicate) must agree with the
    every-pair mask path and with the FP64 oracle when coverage differences sit exactly on, and one ulp either
    side of, the percentile bounds (coverageNeighbours.py:42-50 is inclusive on both ends)."""
    from dbb_b200 import _pairs
    from dbb_b200.distributions import PairTables
    meta, sig64 = _metagenome(23, 900)
    td, gcd, cvd = _dists()
    n = meta.n_scaffolds
    rows = list(range(n))
    probe = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd,
                       gcDistPer=80, covDist=cvd, covDistPer=80)
    cov = np.full(n, 5.0)
    hi = probe.cov_hi[probe.cov_lbin]
    lo = probe.cov_lo[probe.cov_lbin]
    on_hi, on_lo = 5.0 * (1.0 + hi / 100.0), 5.0 * (1.0 + lo / 100.0)
    sel = np.arange(n) % 8
    cov = np.where(sel == 1, on_hi, cov)
    cov = np.where(sel == 2, np.nextafter(on_hi, np.inf), cov)
    cov = np.where(sel == 3, np.nextafter(on_hi, -np.inf), cov)
    cov = np.where(sel == 4, on_lo, cov)
    cov = np.where(sel == 5, np.nextafter(on_lo, -np.inf), cov)
    cov = np.where(sel == 6, np.nextafter(on_lo, np.inf), cov)
    cov[7] = 0.0; cov[15] = -3.0; cov[23] = 1e-45; cov[31] = 1e36; cov[39] = 1e300
    tables = PairTables(meta.length, GC=meta.gc_obs, coverage=cov, tdDist=td, tdDistPer=80, gcDist=gcd,
                        gcDistPer=80, covDist=cvd, covDistPer=80)
    filt = _pairs.FILTER_GC | _pairs.FILTER_COV
    slow = _pairs.run_pairs_host(sig64, tables, k=20, topk_filter=filt, want_count=True, want_cov_mask=True)
    fast = _pairs.run_pairs_host(sig64, tables, k=20, topk_filter=filt, want_count=True)
    ot = O.DistTables(gcd, 80, td, 80, cvd, 80)
    with np.errstate(over='ignore', invalid='ignore'):
        o_cov = O.cov_rows_np(meta.length, cov, rows, ot, None)
    assert np.array_equal(util.unpack_mask(slow['cov_mask'], n), o_cov)
    assert 0.05 < o_cov.mean() < 0.95
    for key in ('count', 'neighbour_bp', 'idx', 'dist'):
        assert np.array_equal(fast[key], slow[key]), key
    cov_only = PairTables(meta.length, coverage=cov, covDist=cvd, covDistPer=80)
    r = _pairs.run_pairs_host(sig64, cov_only, k=0, want_count=True)
    assert np.array_equal(r['count'], o_cov.sum(axis=1))


def test_euclidean_metric_knn_and_bounded_count():
    """metric='l2' (an addition; the reference only has Manhattan): top-k and the bounded neighbourhood size against
    float64 numpy, tolerance 2e-6 absolute on the distance as for L1."""
    from dbb_b200 import _pairs, _lib
    from dbb_b200.distributions import PairTables
    meta, sig64 = _metagenome(24, 1100, gmax=20000)
    n = meta.n_scaffolds
    rows = list(range(n))
    td = _dists()[0]
    tables = PairTables(meta.length, tdDist=td, tdDistPer=80)
    s = sig64.astype(np.float32).astype(np.float64)
    from scipy.spatial.distance import cdist
    d64 = cdist(s, s, 'euclidean')
    r = _pairs.run_pairs_host(sig64, tables, k=20, want_count=True, metric='l2')
    util.check_topk(r['idx'], r['dist'].astype(np.float64), d64, rows, 20)
    shorter_is_i = meta.length[:, None] <= meta.length[None, :]          # ties -> i (tdNeighbours.py:43-47)
    bound = np.where(shorter_is_i, tables.td_bound[:, None], tables.td_bound[None, :]).astype(np.float64)
    exact = d64 < bound
    near = np.abs(d64 - bound) <= util.DIST_ATOL
    lo = (exact & ~near).sum(axis=1)
    hi = (exact | near).sum(axis=1)
    assert ((r['count'] >= lo) & (r['count'] <= hi)).all()
    assert (r['count'] == exact.sum(axis=1)).mean() > 0.99
    part = _pairs.run_pairs_host(sig64, tables, k=5, metric='l2', row0=130, row1=700)
    full5 = _pairs.run_pairs_host(sig64, tables, k=5, metric='l2')
    assert np.array_equal(part['idx'], full5['idx'][130:700])
    plain = _pairs.run_pairs_host(sig64, PairTables(meta.length), k=20, metric='l2')     # no bound at all
    util.check_topk(plain['idx'], plain['dist'].astype(np.float64), d64, rows, 20)
    with pytest.raises(_lib.DbbError):
        _pairs.run_pairs_host(sig64, PairTables(meta.length, GC=meta.gc_obs, gcDist=_dists()[1], gcDistPer=80), k=20, metric='l2')
    with pytest.raises(ValueError):
        _pairs.run_pairs_host(sig64, tables, k=20, metric='cosine')


def test_pruned_sweep_equals_full_sweep():
    """device.pairs_pruned (sorted by a projection, tiles that cannot matter skipped, second launch for unfiltered
    top-k) returns exactly what the full sweep returns: indices, distances bit for bit, counts, base pairs."""
    import torch
    from dbb_b200 import device, synthetic
    from dbb_b200.distributions import PairTables
    from dbb_b200.genomicSignatures import GenomicSignatures
    meta = synthetic.make_metagenome(seed=51, n_scaffolds=2600, min_len=1000, max_len=60000, n_genomes=12)
    seqs = synthetic.generate_sequences(meta)
    _, freq = GenomicSignatures(4, 1).signatures(seqs)
    freq[100] = freq[101]                                               # exact distance ties: order by original id
    freq[2000] = freq[101]
    sig = device.pad_signatures(torch.from_numpy(freq).cuda())
    td, gcd, cvd = _dists()
    n = meta.n_scaffolds
    cases = [
        (PairTables(meta.length, tdDist=td, tdDistPer=80), 0),
        (PairTables(meta.length, tdDist=td, tdDistPer=80), device.FILTER_TD),
        (PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd, gcDistPer=80,
                    covDist=cvd, covDistPer=80), device.FILTER_GC | device.FILTER_COV),
        (PairTables(meta.length), 0),                                   # no bound at all: plain k-NN
    ]
    skipped = []
    for tables, filt in cases:
        dt = device.DevicePairTables(tables)
        cnt = tables.td_bound is not None                              # without a bound every pair counts: nothing to skip
        full = device.pairs(sig, dt, k=20, topk_filter=filt, want_count=cnt)
        pr = device.pairs_pruned(sig, dt, k=20, topk_filter=filt, want_count=cnt, n_classes=(1 if filt == device.FILTER_TD else 3))
        got = device.unsort_rows(pr, n)
        for key in ('idx', 'dist') + (('count', 'neighbour_bp') if cnt else ()):
            assert torch.equal(got[key], full[key]), (key, filt)
        skipped.append(1.0 - pr['tiles_visited'] / pr['tiles_total'])
    assert skipped[1] > 0.05                                           # the TD-filtered case really skips tiles
    with pytest.raises(ValueError):
        device.pairs_pruned(sig, device.DevicePairTables(cases[3][0]), k=20, want_count=True)
    # row tiles dealt to three ranks by work (what each GPU takes): together they are the full result
    dt3 = device.DevicePairTables(cases[0][0])
    full3 = device.pairs(sig, dt3, k=20, want_count=True)
    seen = torch.zeros(n, dtype=torch.bool, device='cuda')
    for rank in range(3):
        part = device.pairs_pruned(sig, dt3, k=20, want_count=True, part=(rank, 3))
        ids = part['row_ids']
        assert not bool(seen[ids].any())
        seen[ids] = True
        for key in ('idx', 'dist', 'count', 'neighbour_bp'):
            assert torch.equal(part[key], full3[key][ids]), (key, rank)
    assert bool(seen.all())


def test_pruned_sweep_with_nan_signatures_and_ranks_beyond_the_row_tiles():
    """Scaffolds without a valid window have NaN signatures (genomicSignatures.py:147-148): the planner leaves them out of
    the tile intervals, and a whole row tile of them still writes its (empty) results.  Also more ranks than row tiles."""
    import torch
    from dbb_b200 import device, synthetic
    from dbb_b200.distributions import PairTables
    from dbb_b200.genomicSignatures import GenomicSignatures
    meta = synthetic.make_metagenome(seed=52, n_scaffolds=900, min_len=1000, max_len=30000, n_genomes=5)
    seqs = synthetic.generate_sequences(meta)
    _, freq = GenomicSignatures(4, 1).signatures(seqs)
    freq[7] = np.nan
    freq[300:450] = np.nan                                              # more than one whole tile of the sorted order
    sig = device.pad_signatures(torch.from_numpy(freq).cuda())
    td, _, _ = _dists()
    n = meta.n_scaffolds
    for filt, ncls in ((0, 1), (device.FILTER_TD, 8)):                   # one class: the 151 NaN rows fill the first tile
        dt = device.DevicePairTables(PairTables(meta.length, tdDist=td, tdDistPer=80))
        full = device.pairs(sig, dt, k=20, topk_filter=filt, want_count=True)
        got = device.unsort_rows(device.pairs_pruned(sig, dt, k=20, topk_filter=filt, want_count=True, n_classes=ncls), n)
        for key in ('idx', 'dist', 'count', 'neighbour_bp'):
            assert torch.equal(got[key], full[key]), (key, filt)
        assert int(full['count'][7]) == 0 and int(full['idx'][7, 0]) == -1
    dt = device.DevicePairTables(PairTables(meta.length, tdDist=td, tdDistPer=80))
    full = device.pairs(sig, dt, k=20, want_count=True)
    seen = 0
    for rank in range(10):                                              # 8 row tiles, 10 ranks: ranks 8 and 9 get nothing
        part = device.pairs_pruned(sig, dt, k=20, want_count=True, part=(rank, 10))
        ids = part['row_ids']
        seen += int(ids.numel())
        for key in ('idx', 'dist', 'count', 'neighbour_bp'):
            assert torch.equal(part[key], full[key][ids]), (key, rank)
    assert seen == n


def test_tdneighbours_knn_public_api_pruned_and_plain_agree_with_the_oracle():
    """TdNeighbours.knn (the array-native addition of SURVEY 8b): the default (pruned sweep) and prune=False (plain sweep
    through dbb_pairs_host) return the same arrays, and those match the oracle's k-NN and TD neighbourhood sizes."""
    from dbb_b200.tdNeighbours import TdNeighbours
    meta, sig64 = _metagenome(25, 1500, gmax=40000)
    td, gcd, cvd = _dists()
    n = meta.n_scaffolds
    rows = list(range(n))
    a = TdNeighbours(None).knn(sig64, meta.length, k=20, tdDist=td, tdDistPer=80)
    b = TdNeighbours(None).knn(sig64, meta.length, k=20, tdDist=td, tdDistPer=80, prune=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    s32 = sig64.astype(np.float32).astype(np.float64)
    d64 = O.l1_rows_np(s32, rows)
    util.check_topk(a[0], a[1].astype(np.float64), d64, rows, 20)
    ot = O.DistTables(tdDist=td, tdDistPer=80)
    o_td = O.td_rows_np(s32, meta.length, rows, ot, None)
    assert np.abs(a[2] - o_td.sum(axis=1)).max() <= 1
    c = TdNeighbours(None).knn(sig64, meta.length, GC=meta.gc_obs, cov=meta.coverage, k=10, tdDist=td, tdDistPer=80, gcDist=gcd,
                               gcDistPer=80, covDist=cvd, covDistPer=80)
    d = TdNeighbours(None).knn(sig64, meta.length, GC=meta.gc_obs, cov=meta.coverage, k=10, tdDist=td, tdDistPer=80, gcDist=gcd,
                               gcDistPer=80, covDist=cvd, covDistPer=80, prune=False)
    for x, y in zip(c, d):
        assert np.array_equal(x, y)


def test_pruned_sweep_refuses_signatures_that_do_not_sum_to_one():
    """The tile-skipping bound L1 >= 2|p_a - p_b| needs rows that sum to 1.  Raw counts (or rows zeroed by nan_to_num)
    break it: the library counts such rows on the device and fails the call (prune=1 / prune=True), or runs the exact
    full sweep instead (prune=-1 / the default of TdNeighbours.knn) -- never a silently wrong neighbour list."""
    import torch
    from dbb_b200 import _lib, device
    from dbb_b200.distributions import PairTables
    from dbb_b200.tdNeighbours import TdNeighbours
    meta, sig64 = _metagenome(61, 1200, gmax=30000)
    td, _, _ = _dists()
    bad = sig64.copy()
    bad[5] *= 1.5                                                       # sums to 1.5
    bad[700] = 0.0                                                      # a NaN row that somebody zeroed
    tables = PairTables(meta.length, tdDist=td, tdDistPer=80)
    dt = device.DevicePairTables(tables)
    sig = device.pad_signatures(torch.from_numpy(bad.astype(np.float32)).cuda())
    with pytest.raises(_lib.DbbError, match='do not sum to 1'):
        device.pairs_pruned(sig, dt, k=20, want_count=True)
    full = device.pairs(sig, dt, k=20, want_count=True)
    auto = device.pairs_pruned(sig, dt, k=20, want_count=True, prune=-1)
    assert not auto['pruned']
    for key in ('idx', 'dist', 'count', 'neighbour_bp'):
        assert torch.equal(auto[key], full[key]), key
    with pytest.raises(_lib.DbbError, match='do not sum to 1'):
        TdNeighbours(None).knn(bad, meta.length, k=20, tdDist=td, tdDistPer=80, prune=True)
    a = TdNeighbours(None).knn(bad, meta.length, k=20, tdDist=td, tdDistPer=80)                  # auto: full sweep
    b = TdNeighbours(None).knn(bad, meta.length, k=20, tdDist=td, tdDistPer=80, prune=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # rows that sum to 1 within float32 rounding pass, NaN rows are exempt
    ok = sig64.copy()
    ok[9] = np.nan
    good = device.pairs_pruned(device.pad_signatures(torch.from_numpy(ok.astype(np.float32)).cuda()), dt, k=20, want_count=True)
    assert good['pruned']


def test_c4_law_eight_ranks_against_the_oracle():
    """BASELINE.json configs[3] at reduced scale (30 000 scaffolds of the C4 length law, 1-14 kb, 2 000 genomes): the
    pruned sweep dealt to 8 ranks (each emulated on this GPU) covers every row exactly once, and a row sample of every
    rank agrees with the oracle's float64 rows (tdNeighbours.py:41-51) -- top-k within the tie band, counts, base pairs."""
    import torch
    from dbb_b200 import device, synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.distributions import PairTables
    from oracle import checks
    meta = synthetic.make_config('C4', 0.03)
    d_ascii, seq_off = device.synth_ascii(meta)
    packed = device.pack_ascii(d_ascii, seq_off)
    n = meta.n_scaffolds
    freq = torch.zeros((n, 140), dtype=torch.float32, device='cuda')
    device.CountPlan(packed).run(None, freq)
    td = load_td_dist()
    dt = device.DevicePairTables(PairTables(meta.length, tdDist=td, tdDistPer=80))
    ot = O.DistTables(tdDist=td, tdDistPer=80)
    sig32 = freq[:, :136].cpu().numpy()
    seen = torch.zeros(n, dtype=torch.int32, device='cuda')
    rng = np.random.RandomState(4)
    for rank in range(8):
        part = device.pairs_pruned(freq, dt, k=20, want_count=True, part=(rank, 8))
        ids = part['row_ids']
        seen[ids] += 1
        pos = np.sort(rng.choice(int(ids.numel()), 12, replace=False))
        tp = torch.from_numpy(pos).cuda()
        checks.check_knn_sample(ids[tp].cpu().numpy(), part['idx'][tp].cpu().numpy(), part['dist'][tp].cpu().numpy(),
                                part['count'][tp].cpu().numpy(), part['neighbour_bp'][tp].cpu().numpy(), sig32, meta.length, 20, ot)
        assert part['tiles_visited'] < part['tiles_total']
    assert bool((seen == 1).all())


def test_eight_warp_instance_passes_the_same_tests():
    """The pair kernel has two instances (csrc/pairs.cu, Cfg<WARPS>): two 4-warp CTAs per SM on 64-row half tiles while the
    signature matrix fits L2, one 8-warp CTA per SM on 128-row tiles beyond.  The inputs of this file are small, so the
    8-warp instance is forced here for the whole file in a child process."""
    import os
    import subprocess
    import sys
    if os.environ.get('DBB_K2_WARPS'):
        pytest.skip('already a forced-instance run')
    env = dict(os.environ, DBB_K2_WARPS='8')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-m', 'gpu', '-q', '-x', '-k', 'not eight_warp'],
                       env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
