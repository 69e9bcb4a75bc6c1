"""Row-sample parity checks of stage 2 against the oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used by tests/ and by bench.py's untimed parity check (at every GPU count and workload, incl. the 1 M-scaffold
config): a seeded sample of rows of the all-pairs problem is recomputed by the oracle against ALL N columns --
float64 L1 distances by the C restatement of tdNeighbours.py:41-51 (oracle/dbb_oracle.c::oracle_td_rows), the GC and
coverage predicates by the numpy restatement of gcNeighbours.py:43-52 / coverageNeighbours.py:41-50 -- and the GPU
rows must agree within the tolerances SURVEY.md section 8d states (written below).
"""
import numpy as np

from . import dbb_oracle as O

DIST_ATOL = 2e-6          # |d32 - d64| <= 2e-6 absolute; also the tie band of the TD bound and of the k-th distance


def check_topk(idx, dist, d64_rows, rows, k, allowed=None, atol=DIST_ATOL):
    """GPU top-k (idx, dist) vs float64 distance rows: distances within atol of float64, sorted, no duplicates, and the
    index set equal to the oracle's up to swaps among candidates within atol of the oracle's k-th distance."""
    for r, i in enumerate(rows):
        d = d64_rows[r]
        ok = ~np.isnan(d)
        ok[i] = False
        if allowed is not None:
            ok &= allowed[r].astype(bool)
        n_cand = int(ok.sum())
        m = min(k, n_cand)
        got = idx[r]
        assert (got[m:] == -1).all() and np.isinf(dist[r][m:]).all(), 'row %d: padding' % i
        got = got[:m]
        assert len(set(got.tolist())) == m and ok[got].all(), 'row %d: invalid/duplicate neighbours' % i
        assert np.abs(dist[r][:m] - d[got]).max(initial=0) <= atol, 'row %d: distance error' % i
        assert (np.diff(dist[r][:m]) >= 0).all(), 'row %d: not sorted' % i
        if m == 0:
            continue
        dm = np.where(ok, d, np.inf)
        kth = np.partition(dm, m - 1)[m - 1]                      # the oracle's m-th smallest candidate distance
        must = set(np.nonzero(dm < kth - atol)[0].tolist())        # clearly inside
        may_mask = dm <= kth + atol                                # inside or in the tie band
        g = set(got.tolist())
        assert must <= g and may_mask[got].all(), 'row %d: neighbour set differs beyond the tie band' % i


def check_knn_sample(rows, idx, dist, count, neighbour_bp, sig32, length, k, td_tables, fixedSeqLen=None,
                     GC=None, coverage=None, topk_filter_gc_cov=False, atol=DIST_ATOL):
    """rows: sampled row ids (into the N scaffolds); idx/dist/count/neighbour_bp: the GPU results OF THOSE ROWS;
    sig32 [N,136]: the float32 signatures the GPU used (widened here); td_tables: oracle DistTables (with GC / coverage
    tables when those predicates were enabled).  Asserts parity; returns a dict of what was checked."""
    from . import c_oracle
    rows = [int(r) for r in rows]
    length = np.asarray(length, dtype=np.int64)
    n = length.shape[0]
    sig64 = np.ascontiguousarray(sig32, dtype=np.float64)
    if fixedSeqLen is not None:
        per = np.full(n, td_tables.tdBound[O.nearest_index_sorted(td_tables.tdLens, np.array([fixedSeqLen]))[0]])
    else:
        per = td_tables.tdBound[O.nearest_index_sorted(td_tables.tdLens, length)]
    d64, _ = c_oracle.td_rows(sig64, length, per, rows, want_dist=True, want_mask=False)
    li = length[rows][:, None]
    bound = np.where(li > length[None, :], per[None, :], per[rows][:, None])         # shorter scaffold's bound, ties -> i
    with np.errstate(invalid='ignore'):
        td = d64 < bound
    allowed = None
    other = np.ones_like(td)
    if GC is not None:
        other &= O.gc_rows_np(length, GC, rows, td_tables, fixedSeqLen).astype(bool)
    if coverage is not None:
        other &= O.cov_rows_np(length, coverage, rows, td_tables, fixedSeqLen).astype(bool)
    if topk_filter_gc_cov:
        allowed = other
    check_topk(np.asarray(idx), np.asarray(dist, dtype=np.float64), d64, rows, k, allowed=allowed, atol=atol)
    allm = td & other
    near = (np.abs(d64 - bound) <= atol) & other
    cnt_o = allm.sum(axis=1)
    n_near = near.sum(axis=1)
    count = np.asarray(count).astype(np.int64)
    bp = np.asarray(neighbour_bp).astype(np.int64)
    assert (np.abs(count - cnt_o) <= n_near).all(), 'neighbour count differs beyond the tie band'
    exact = n_near == 0
    assert np.array_equal(count[exact], cnt_o[exact]), 'neighbour count differs'
    assert np.array_equal(bp[exact], (allm * length[None, :]).sum(axis=1)[exact]), 'neighbour_bp differs'
    return {'rows': len(rows), 'columns': int(n), 'rows_exact_count': int(exact.sum()), 'pairs_in_tie_band': int(n_near.sum()),
            'mean_neighbours': float(cnt_o.mean())}
