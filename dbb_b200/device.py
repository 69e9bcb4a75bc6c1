"""Device-resident pipeline: torch tensors for memory and streams, libdbbcuda kernels for the work.

torch is plumbing here (allocation, the current CUDA stream, torch.distributed); every kernel that
touches sequence or signature data is one of ours, launched through the C ABI on torch's current stream.
"""
import ctypes

import numpy as np
import torch

from . import _lib

FILTER_TD, FILTER_GC, FILTER_COV = 1, 2, 4


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.DbbError('no CUDA device: dbb_b200 has no CPU fallback')
    _lib.lib()


def block_offsets(lengths):
    """int64[N+1] first 64-base block of every scaffold (host)."""
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    off = np.zeros(lengths.shape[0] + 1, dtype=np.int64)
    _lib.check(_lib.lib().dbb_block_offsets(_lib.ptr(lengths), lengths.shape[0], _lib.ptr(off)))
    return off


class PackedBatch(object):
    """2-bit code plane + validity plane of a batch of scaffolds in HBM (layout: include/dbb_cuda.h)."""

    def __init__(self, lengths, device=None):
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        self.n = int(self.lengths.shape[0])
        self.blk_off = block_offsets(self.lengths)
        self.total_blocks = int(self.blk_off[-1])
        dev = device or torch.device('cuda', torch.cuda.current_device())
        self.codes = torch.empty(max(self.total_blocks, 1) * 4 + 4, dtype=torch.int32, device=dev)
        self.valid = torch.empty(max(self.total_blocks, 1) + 2, dtype=torch.int64, device=dev)
        self.d_blk_off = torch.from_numpy(self.blk_off).to(dev)
        self.total_bp = int(self.lengths.sum())

    def algorithmic_bytes(self):
        """ceil(L/4) + ceil(L/8) + 136*4 + 136*4 + 16 per scaffold."""
        L = self.lengths
        return int(((L + 3) // 4).sum() + ((L + 7) // 8).sum() + self.n * (136 * 4 + 136 * 4 + 16))


def pack_ascii(d_ascii, seq_off, packed=None):
    """d_ascii: uint8 cuda tensor holding the scaffolds back to back; seq_off: host int64[N+1]."""
    require_cuda()
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    lengths = np.diff(seq_off)
    if packed is None:
        packed = PackedBatch(lengths, d_ascii.device)
    d_off = torch.from_numpy(seq_off).to(d_ascii.device)
    _lib.check(_lib.lib().dbb_pack_ascii_dev(_p(d_ascii), _p(d_off), _p(packed.d_blk_off), packed.n,
                                             packed.total_blocks, _p(packed.codes), _p(packed.valid), _stream()))
    return packed


class CountPlan(object):
    """Device work list for one PackedBatch (build once, run many times)."""

    def __init__(self, packed):
        require_cuda()
        self.packed = packed
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().dbb_count_plan_create(_lib.ptr(packed.blk_off), packed.n, ctypes.byref(h)))
        self._h = h
        self.launches = int(_lib.lib().dbb_count_plan_launches(h))

    def run(self, counts=None, freq=None):
        """counts: int32/uint32-viewed cuda tensor [N,136] or None; freq: float32 cuda tensor [N,stride>=136]."""
        stride = 136 if freq is None else int(freq.stride(0))
        _lib.check(_lib.lib().dbb_tetra_count_dev(self._h, _p(self.packed.codes), _p(self.packed.valid),
                                                  _p(counts), _p(freq), stride, _stream()))

    def close(self):
        if self._h:
            _lib.lib().dbb_count_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_ascii(meta, device=None):
    """Generate the metagenome's ASCII on the device (same bytes as synthetic.generate_sequences).
    Returns (uint8 cuda tensor, host seq_off int64[N+1])."""
    require_cuda()
    dev = device or torch.device('cuda', torch.cuda.current_device())
    lengths = np.ascontiguousarray(meta.length, dtype=np.int64)
    seq_off = np.zeros(lengths.shape[0] + 1, dtype=np.int64)
    seq_off[1:] = np.cumsum(lengths)
    blk_off = block_offsets(lengths)
    to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_ascii = torch.empty(max(int(seq_off[-1]), 1), dtype=torch.uint8, device=dev)
    args = [to(seq_off, np.int64), to(blk_off, np.int64)]
    tail = [to(meta.genome, np.int32), to(meta.cum.astype(np.int64), np.int64).to(torch.int32),
            to(meta.n_start, np.int64), to(meta.n_len, np.int64), to(meta.lc_start, np.int64), to(meta.lc_len, np.int64)]
    _lib.check(_lib.lib().dbb_synth_ascii_dev(ctypes.c_uint64(meta.seed), lengths.shape[0], _p(args[0]), _p(args[1]),
                                              int(blk_off[-1]), _p(tail[0]), _p(tail[1]), _p(tail[2]), _p(tail[3]),
                                              _p(tail[4]), _p(tail[5]), _p(d_ascii), _stream()))
    torch.cuda.current_stream().synchronize()
    return d_ascii, seq_off


class DevicePairTables(object):
    """distributions.PairTables uploaded once."""

    def __init__(self, tables, device=None):
        dev = device or torch.device('cuda', torch.cuda.current_device())
        up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.n = tables.n
        self.length = up(tables.length)
        self.td_bound = up(tables.td_bound)
        self.gc, self.gc_mbin, self.gc_lbin = up(tables.gc), up(tables.gc_mbin), up(tables.gc_lbin)
        self.gc_lo, self.gc_hi = up(tables.gc_lo), up(tables.gc_hi)
        self.cov, self.cov_lbin = up(tables.cov), up(tables.cov_lbin)
        self.cov_lo, self.cov_hi = up(tables.cov_lo), up(tables.cov_hi)


def pad_signatures(dense):
    """[N,136] float32 cuda -> [N,140] padded copy (the layout the pair kernel's TMA boxes expect)."""
    n = dense.shape[0]
    out = torch.empty((n, _lib.SIG_STRIDE), dtype=torch.float32, device=dense.device)
    _lib.check(_lib.lib().dbb_pad_signatures_dev(_p(dense.contiguous()), n, _p(out), _stream()))
    return out


def pairs(sig_padded, dt, row0=0, row1=None, k=0, topk_filter=0, want_count=True, masks=(), metric='l1', sparse=None):
    """Rows [row0,row1) of the all-pairs problem on the current device/stream.
    sig_padded: float32 cuda [N,140]; dt: DevicePairTables.  Returns a dict of cuda tensors.
    sparse: optional (col_id int32[N], item_order int32[row tiles], tile_off int64[row tiles + 1], tile_list int32[...],
    split_all, n_active_row_tiles) -- the sparse sweep of include/dbb_cuda.h (see pairs_pruned)."""
    require_cuda()
    n = dt.n
    row1 = n if row1 is None else row1
    rows = row1 - row0
    dev = sig_padded.device
    assert sig_padded.shape == (n, _lib.SIG_STRIDE) and sig_padded.is_contiguous() and sig_padded.dtype == torch.float32
    inp = _lib.PairsInput()
    inp.n = n
    inp.sig, inp.length, inp.td_bound = _p(sig_padded), _p(dt.length), _p(dt.td_bound)
    inp.metric = _lib.METRICS[metric]
    if sparse is not None:
        inp.col_id, inp.item_order, inp.tile_off, inp.tile_list = (_p(t) for t in sparse[:4])
        inp.split_all, inp.n_active_row_tiles = int(sparse[4]), int(sparse[5])
        inp.split_tiles = int(sparse[6]) if len(sparse) > 6 else 0
    if dt.gc is not None:
        inp.gc, inp.gc_mbin, inp.gc_lbin, inp.gc_lo, inp.gc_hi = _p(dt.gc), _p(dt.gc_mbin), _p(dt.gc_lbin), _p(dt.gc_lo), _p(dt.gc_hi)
        inp.gc_nm, inp.gc_nl = dt.gc_lo.shape
    if dt.cov is not None:
        inp.cov, inp.cov_lbin, inp.cov_lo, inp.cov_hi = _p(dt.cov), _p(dt.cov_lbin), _p(dt.cov_lo), _p(dt.cov_hi)
        inp.cov_nl = dt.cov_lo.shape[0]
    out = _lib.PairsOutput()
    res = {}
    out.k, out.topk_filter = k, topk_filter
    if k > 0:
        res['idx'] = torch.empty((rows, k), dtype=torch.int32, device=dev)
        res['dist'] = torch.empty((rows, k), dtype=torch.float32, device=dev)
        out.knn_idx, out.knn_dist = _p(res['idx']), _p(res['dist'])
    if want_count:
        res['count'] = torch.empty(rows, dtype=torch.int32, device=dev)
        res['neighbour_bp'] = torch.empty(rows, dtype=torch.int64, device=dev)
        out.count, out.neighbour_bp = _p(res['count']), _p(res['neighbour_bp'])
    words = (n + 31) // 32
    out.mask_words = words
    for name in masks:
        res[name + '_mask'] = torch.zeros((rows, words), dtype=torch.int32, device=dev)
        setattr(out, name + '_mask', _p(res[name + '_mask']))
    if rows and n:
        _lib.check(_lib.lib().dbb_pairs_dev(ctypes.byref(inp), row0, row1, ctypes.byref(out), _stream()))
    return res


def pairs_pruned(sig_padded, dt, k=20, topk_filter=0, want_count=True, metric='l1', margin=1e-5, n_classes=8, timings=None,
                 part=None, prune=1, ctx=None):
    """k nearest neighbours + neighbourhood sizes of all rows through ONE C call, ``dbb_knn_dev`` (csrc/knn.cu): same
    results as ``pairs`` (no masks), skipping 128 x 128 tiles of the pair matrix that cannot matter.

    Signatures sum to 1, so for ANY weights w_k in [0, 1] and p = sum_k w_k sig_k:  L1(a, b) >= 2 |p_a - p_b|
    (sum_k (w_k - 1/2)(a_k - b_k) equals p_a - p_b and is at most L1 / 2 in magnitude).  With the scaffolds sorted by
    (TD-bound class, p) every tile has a p interval; a tile pair whose intervals are farther apart than half the largest
    TD bound of its rows and columns holds no TD neighbour (the bound of a pair is the bound of its shorter scaffold),
    hence nothing that counts and no candidate of a TD-filtered top-k.  For a top-k that is not TD-filtered a second
    launch visits the tiles whose lower bound does not exceed the k-th best distance found so far of some row of the row
    tile, and the lists are merged; results are identical to the unpruned sweep, ties included.  The library checks on
    the device that every finite signature row sums to 1 (the bound's precondition) and refuses the call otherwise
    (``prune=1``) or runs the full sweep (``prune=-1``).
    part=(rank, world): multi-GPU -- the row tiles of the sorted order are dealt to the ranks by decreasing work; the
    rank's rows come back in ascending sorted position.  The dict carries ``row_ids`` (original ids of the returned
    rows), ``tiles_visited`` / ``tiles_total`` / ``launches`` and the device times of the phases."""
    require_cuda()
    if metric != 'l1':
        raise ValueError('pairs_pruned: the projection bound is for the Manhattan distance')
    if want_count and dt.td_bound is None:
        raise ValueError('pairs_pruned: without a TD bound every pair may count, nothing can be skipped (use pairs)')
    n = dt.n
    dev = sig_padded.device
    assert sig_padded.shape == (n, _lib.SIG_STRIDE) and sig_padded.is_contiguous() and sig_padded.dtype == torch.float32
    rank, world = part if part is not None else (0, 1)
    n_rt = (n + 127) // 128
    cap = max(1, n if world <= 1 else min(n, 128 * len(range(rank, n_rt, world))))   # (a rank beyond the row tiles gets no rows)
    inp = _lib.PairsInput()
    inp.n = n
    inp.sig, inp.length, inp.td_bound = _p(sig_padded), _p(dt.length), _p(dt.td_bound)
    inp.metric = _lib.METRICS[metric]
    if dt.gc is not None:
        inp.gc, inp.gc_mbin, inp.gc_lbin, inp.gc_lo, inp.gc_hi = _p(dt.gc), _p(dt.gc_mbin), _p(dt.gc_lbin), _p(dt.gc_lo), _p(dt.gc_hi)
        inp.gc_nm, inp.gc_nl = dt.gc_lo.shape
    if dt.cov is not None:
        inp.cov, inp.cov_lbin, inp.cov_lo, inp.cov_hi = _p(dt.cov), _p(dt.cov_lbin), _p(dt.cov_lo), _p(dt.cov_hi)
        inp.cov_nl = dt.cov_lo.shape[0]
    par = _lib.KnnParams(k, topk_filter, prune, n_classes, float(margin), rank, world)
    res = {'idx': torch.empty((cap, k), dtype=torch.int32, device=dev), 'dist': torch.empty((cap, k), dtype=torch.float32, device=dev),
           'row_ids': torch.empty(cap, dtype=torch.int64, device=dev)}
    out = _lib.KnnOutput()
    out.cap_rows = cap
    out.row_ids, out.knn_idx, out.knn_dist = _p(res['row_ids']), _p(res['idx']), _p(res['dist'])
    if want_count:
        res['count'] = torch.empty(cap, dtype=torch.int32, device=dev)
        res['neighbour_bp'] = torch.empty(cap, dtype=torch.int64, device=dev)
        out.count, out.neighbour_bp = _p(res['count']), _p(res['neighbour_bp'])
    if n:
        _lib.check(_lib.lib().dbb_knn_dev(ctx, ctypes.byref(inp), ctypes.byref(par), ctypes.byref(out), _stream()))
    rows = int(out.n_rows)
    for key in list(res):
        res[key] = res[key][:rows]
    res['tiles_visited'], res['tiles_total'] = int(out.tiles_visited), int(out.tiles_total)
    res['launches'], res['pruned'] = int(out.launches), bool(out.pruned)
    res['tiles_second'], res['row_tiles_second'] = int(out.tiles_second), int(out.row_tiles_second)
    ms = {'plan': float(out.ms_plan), 'sweep': float(out.ms_sweep), 'second sweep + merge': float(out.ms_second)}
    if isinstance(timings, list):
        timings.append(ms)
    elif timings is not None:
        timings.update(ms)
    return res


def unsort_rows(res, n):
    """Results of a pairs_pruned call scattered to the rows named by ``row_ids`` (single GPU: already the caller's
    order; multi-GPU: the rank's rows at their original positions of an [n, ...] array)."""
    out = {}
    ids = res['row_ids']
    for key in ('idx', 'dist', 'count', 'neighbour_bp'):
        if key in res:
            t = torch.empty((n,) + tuple(res[key].shape[1:]), dtype=res[key].dtype, device=res[key].device)
            t[ids] = res[key]
            out[key] = t
    return out


# ------------------------------------------------------------------------------------------------
# N2: sequence side of preprocess (split at N runs, base composition, contigs as slices)
# ------------------------------------------------------------------------------------------------

def pack_ranges(d_ascii, begin, end, want_codes=True, want_nmask=False):
    """Sequences that are slices ascii[begin[i], end[i]) of one device buffer (host int64 arrays).
    Returns a PackedBatch; ``.nmask`` (int64 cuda tensor, one word per 64-base block) when requested."""
    require_cuda()
    begin = np.ascontiguousarray(begin, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    packed = PackedBatch(end - begin, d_ascii.device)
    packed.nmask = torch.empty(max(packed.total_blocks, 1) + 2, dtype=torch.int64, device=d_ascii.device) if want_nmask else None
    d_b, d_e = torch.from_numpy(begin).to(d_ascii.device), torch.from_numpy(end).to(d_ascii.device)
    _lib.check(_lib.lib().dbb_pack_ranges_dev(_p(d_ascii), _p(d_b), _p(d_e), _p(packed.d_blk_off), packed.n,
                                              packed.total_blocks, _p(packed.codes) if want_codes else None,
                                              _p(packed.valid) if want_codes else None, _p(packed.nmask), _stream()))
    return packed


def split_scaffolds(packed, min_n):
    """SplitScaffolds.splits (splitScaffolds.py:27-53) for every sequence of a PackedBatch that carries an N plane.
    Returns (contig_off int64[N+1], begin int64[M], end int64[M]) on the host; positions are relative to the start of
    their scaffold, contig k of scaffold s sits at contig_off[s] + k."""
    require_cuda()
    if packed.nmask is None:
        raise ValueError('split_scaffolds needs pack_ranges(..., want_nmask=True)')
    dev = packed.nmask.device
    d_len = torch.from_numpy(packed.lengths).to(dev)
    n_contigs = torch.zeros(max(packed.n, 1), dtype=torch.int32, device=dev)
    L = _lib.lib()
    _lib.check(L.dbb_split_scaffolds_dev(_p(packed.nmask), _p(d_len), _p(packed.d_blk_off), packed.n, int(min_n), None,
                                         _p(n_contigs), None, None, _stream()))
    off = np.zeros(packed.n + 1, dtype=np.int64)
    off[1:] = np.cumsum(n_contigs[:packed.n].cpu().numpy().astype(np.int64))
    m = int(off[-1])
    d_off = torch.from_numpy(off).to(dev)
    d_begin = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    d_end = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    _lib.check(L.dbb_split_scaffolds_dev(_p(packed.nmask), _p(d_len), _p(packed.d_blk_off), packed.n, int(min_n), _p(d_off),
                                         None, _p(d_begin), _p(d_end), _stream()))
    return off, d_begin[:m].cpu().numpy(), d_end[:m].cpu().numpy()


def base_counts(d_ascii, begin, end):
    """baseCount (seqUtils.py:258-265) of the slices ascii[begin[i], end[i]): int64[n,4] = A, C, G, T+U on the host."""
    require_cuda()
    begin = np.ascontiguousarray(begin, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    n = int(begin.shape[0])
    out = torch.zeros((max(n, 1), 4), dtype=torch.int64, device=d_ascii.device)
    d_b, d_e = torch.from_numpy(begin).to(d_ascii.device), torch.from_numpy(end).to(d_ascii.device)
    _lib.check(_lib.lib().dbb_base_count_dev(_p(d_ascii), _p(d_b), _p(d_e), n, _p(out), _stream()))
    return out[:n].cpu().numpy()
