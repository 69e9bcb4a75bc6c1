"""Host-buffer driver of the pair kernel (``dbb_pairs_host``), shared by tdNeighbours / gcNeighbours /
coverageNeighbours, plus the lazy row view that stands in for the reference's dense float64 rows."""
import ctypes

import numpy as np

from . import _lib
from .distributions import PairTables

FILTER_TD, FILTER_GC, FILTER_COV = 1, 2, 4


class NeighbourRows(object):
    """Bit-packed N x N neighbour mask (uint32 words, bit j%32 of word j/32)."""

    def __init__(self, bits, n):
        self.bits = bits                # uint32[rows, words]
        self.n = int(n)

    def dense_row(self, i):
        """The exact object the reference stores: float64[N] of 0.0 / 1.0 (tdNeighbours.py:41,51)."""
        b = np.unpackbits(self.bits[i].view(np.uint8), bitorder='little')[:self.n]
        return b.astype(np.float64)

    def row(self, i):
        return NeighbourRow(self, i)

    def counts(self):
        return np.array([int(np.unpackbits(self.bits[i].view(np.uint8), bitorder='little')[:self.n].sum())
                         for i in range(self.bits.shape[0])])


class NeighbourRow(object):
    """Lazy stand-in for one dense row: supports ``row == 1``, ``np.where(row == 1)``, ``np.asarray``,
    ``len`` and indexing (greedy.py:409-411), materialising float64 only on demand."""
    __slots__ = ('_rows', '_i')

    def __init__(self, rows, i):
        self._rows, self._i = rows, i

    def __array__(self, dtype=None, copy=None):
        a = self._rows.dense_row(self._i)
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self._rows.n

    def __getitem__(self, j):
        return self._rows.dense_row(self._i)[j]

    def __eq__(self, other):
        return self._rows.dense_row(self._i) == other

    def __ne__(self, other):
        return self._rows.dense_row(self._i) != other

    __hash__ = None


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run_pairs_host(sig, tables, k=0, topk_filter=0, want_count=False, want_td_mask=False,
                   want_gc_mask=False, want_cov_mask=False, row0=0, row1=None, metric='l1'):
    """sig: float[N,136] (any float dtype, converted to float32).  Returns a dict of numpy outputs.
    metric: 'l1' (Manhattan, the reference's) or 'l2' (Euclidean; k-NN with optional TD-style bound only)."""
    if metric not in _lib.METRICS:
        raise ValueError("metric must be 'l1' or 'l2'")
    L = _lib.lib()
    sig32 = np.ascontiguousarray(sig, dtype=np.float32)
    n = tables.n
    if sig32.shape != (n, _lib.NUM_KMERS):
        raise ValueError('sig must be [N,136]')
    if not 0 <= k <= _lib.MAX_K:
        raise ValueError('k must be in 0..%d' % _lib.MAX_K)
    row1 = n if row1 is None else row1
    rows = row1 - row0
    inp = _lib.PairsInput()
    inp.n = n
    inp.sig = _vp(sig32)
    inp.length = _vp(tables.length)
    inp.td_bound = _vp(tables.td_bound)
    inp.metric = _lib.METRICS[metric]
    if tables.gc is not None:
        inp.gc, inp.gc_mbin, inp.gc_lbin = _vp(tables.gc), _vp(tables.gc_mbin), _vp(tables.gc_lbin)
        inp.gc_lo, inp.gc_hi = _vp(tables.gc_lo), _vp(tables.gc_hi)
        inp.gc_nm, inp.gc_nl = tables.gc_lo.shape
    if tables.cov is not None:
        inp.cov, inp.cov_lbin = _vp(tables.cov), _vp(tables.cov_lbin)
        inp.cov_lo, inp.cov_hi = _vp(tables.cov_lo), _vp(tables.cov_hi)
        inp.cov_nl = tables.cov_lo.shape[0]
    out = _lib.PairsOutput()
    res = {}
    out.k = k
    out.topk_filter = topk_filter
    if k > 0:
        res['idx'] = np.empty((rows, k), dtype=np.int32)
        res['dist'] = np.empty((rows, k), dtype=np.float32)
        out.knn_idx, out.knn_dist = _vp(res['idx']), _vp(res['dist'])
    if want_count:
        res['count'] = np.empty(rows, dtype=np.int32)
        res['neighbour_bp'] = np.empty(rows, dtype=np.int64)
        out.count, out.neighbour_bp = _vp(res['count']), _vp(res['neighbour_bp'])
    words = (n + 31) // 32
    out.mask_words = words
    for key, want, field in (('td_mask', want_td_mask, 'td_mask'), ('gc_mask', want_gc_mask, 'gc_mask'),
                             ('cov_mask', want_cov_mask, 'cov_mask')):
        if want:
            res[key] = np.zeros((rows, words), dtype=np.uint32)
            setattr(out, field, _vp(res[key]))
    if n and rows:
        _lib.check(L.dbb_pairs_host(ctypes.byref(inp), row0, row1, ctypes.byref(out)))
    return res


def run_knn_host(sig, tables, k=20, topk_filter=0, prune=-1, n_classes=8, margin=1e-5, rank=0, world=1, ctx=None):
    """``dbb_knn_host``: sig float[N,136] and the PairTables on the host -> dict of numpy outputs (idx, dist, count,
    neighbour_bp, row_ids, plus the statistics of the call).  prune: 1 tile-skipping sweep, 0 full sweep, -1 auto."""
    L = _lib.lib()
    sig32 = np.ascontiguousarray(sig, dtype=np.float32)
    n = tables.n
    if sig32.shape != (n, _lib.NUM_KMERS):
        raise ValueError('sig must be [N,136]')
    if not 1 <= k <= _lib.MAX_K:
        raise ValueError('k must be in 1..%d' % _lib.MAX_K)
    inp = _lib.PairsInput()
    inp.n = n
    inp.sig, inp.length, inp.td_bound = _vp(sig32), _vp(tables.length), _vp(tables.td_bound)
    inp.metric = _lib.METRICS['l1']
    if tables.gc is not None:
        inp.gc, inp.gc_mbin, inp.gc_lbin = _vp(tables.gc), _vp(tables.gc_mbin), _vp(tables.gc_lbin)
        inp.gc_lo, inp.gc_hi = _vp(tables.gc_lo), _vp(tables.gc_hi)
        inp.gc_nm, inp.gc_nl = tables.gc_lo.shape
    if tables.cov is not None:
        inp.cov, inp.cov_lbin = _vp(tables.cov), _vp(tables.cov_lbin)
        inp.cov_lo, inp.cov_hi = _vp(tables.cov_lo), _vp(tables.cov_hi)
        inp.cov_nl = tables.cov_lo.shape[0]
    n_rt = (n + 127) // 128
    cap = max(1, n if (world <= 1 or not prune) else min(n, 128 * len(range(rank, n_rt, world))))
    par = _lib.KnnParams(k, topk_filter, prune, n_classes, float(margin), rank, world)
    res = {'idx': np.empty((cap, k), dtype=np.int32), 'dist': np.empty((cap, k), dtype=np.float32),
           'count': np.empty(cap, dtype=np.int32), 'neighbour_bp': np.empty(cap, dtype=np.int64),
           'row_ids': np.empty(cap, dtype=np.int64)}
    out = _lib.KnnOutput()
    out.cap_rows = cap
    out.row_ids, out.knn_idx, out.knn_dist = _vp(res['row_ids']), _vp(res['idx']), _vp(res['dist'])
    out.count, out.neighbour_bp = _vp(res['count']), _vp(res['neighbour_bp'])
    if n:
        _lib.check(L.dbb_knn_host(ctx, ctypes.byref(inp), ctypes.byref(par), ctypes.byref(out)))
    rows = int(out.n_rows)
    for key in list(res):
        res[key] = res[key][:rows]
    res.update(tiles_visited=int(out.tiles_visited), tiles_total=int(out.tiles_total), pruned=bool(out.pruned),
               launches=int(out.launches), ms_plan=float(out.ms_plan), ms_sweep=float(out.ms_sweep), ms_second=float(out.ms_second))
    return res


def objects_to_arrays(seqs):
    """The reference passes a list of objects with .tetraSig/.length/.GC/.coverage (greedy.py:133-141)."""
    n = len(seqs)
    sig = np.empty((n, _lib.NUM_KMERS), dtype=np.float32)
    for i, s in enumerate(seqs):
        sig[i] = s.tetraSig
    length = np.array([s.length for s in seqs], dtype=np.int64)
    return sig, length


def attach_rows(seqs, attr, bits, dense_limit):
    """``seqs[i].<attr> = row_i``: a real float64 array for small N (exactly the reference's object),
    a lazy NeighbourRow beyond ``dense_limit`` scaffolds (the dense rows are 8*N^2 bytes)."""
    n = len(seqs)
    rows = NeighbourRows(bits, n)
    for i, s in enumerate(seqs):
        setattr(s, attr, rows.dense_row(i) if n <= dense_limit else rows.row(i))
    return rows
