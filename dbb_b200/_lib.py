"""ctypes binding of libdbbcuda.so (include/dbb_cuda.h).  There is NO CPU fallback: if the library is
missing or CUDA is unavailable, every compute call raises."""
import ctypes
import os

import numpy as np

NUM_KMERS = 136
SIG_STRIDE = 140
BLOCK_BASES = 64
MAX_K = 32

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdbbcuda.so')

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_vp = ctypes.c_void_p


class DbbError(RuntimeError):
    pass


class PairsInput(ctypes.Structure):
    _fields_ = [('n', c_i64), ('sig', c_vp), ('length', c_vp), ('td_bound', c_vp),
                ('gc', c_vp), ('gc_mbin', c_vp), ('gc_lbin', c_vp), ('gc_lo', c_vp), ('gc_hi', c_vp),
                ('gc_nm', c_i32), ('gc_nl', c_i32),
                ('cov', c_vp), ('cov_lbin', c_vp), ('cov_lo', c_vp), ('cov_hi', c_vp), ('cov_nl', c_i32),
                ('metric', c_i32), ('col_id', c_vp), ('item_order', c_vp), ('tile_off', c_vp), ('tile_list', c_vp),
                ('split_all', c_i32), ('n_active_row_tiles', c_i32), ('split_tiles', c_i32), ('reserved0', c_i32)]


METRICS = {'l1': 0, 'manhattan': 0, 'l2': 1, 'euclidean': 1}


class PairsOutput(ctypes.Structure):
    _fields_ = [('k', c_i32), ('topk_filter', c_i32), ('knn_idx', c_vp), ('knn_dist', c_vp),
                ('count', c_vp), ('neighbour_bp', c_vp),
                ('td_mask', c_vp), ('gc_mask', c_vp), ('cov_mask', c_vp), ('mask_words', c_i64)]


class RefineInput(ctypes.Structure):
    _fields_ = [('n_unbinned', c_i64), ('n_cores', c_i64), ('u_sig', c_vp), ('c_sig', c_vp), ('u_gc', c_vp), ('u_cov', c_vp),
                ('c_gc', c_vp), ('c_cov', c_vp), ('u_td_bound', c_vp), ('u_gc_lbin', c_vp), ('u_cov_lbin', c_vp),
                ('c_gc_mbin', c_vp), ('gc_lo', c_vp), ('gc_hi', c_vp), ('gc_nm', c_i32), ('gc_nl', c_i32),
                ('cov_lo', c_vp), ('cov_hi', c_vp), ('cov_nl', c_i32)]


class GreedyInput(ctypes.Structure):
    _fields_ = [('n', c_i64), ('sig', c_vp), ('length', c_vp), ('gc', c_vp), ('cov', c_vp), ('td_bound', c_vp),
                ('gc_lbin', c_vp), ('cov_lbin', c_vp), ('gc_mean_keys', c_vp), ('gc_lo', c_vp), ('gc_hi', c_vp),
                ('gc_nm', c_i32), ('gc_nl', c_i32), ('cov_lo', c_vp), ('cov_hi', c_vp), ('cov_nl', c_i32)]


class PruneInput(ctypes.Structure):
    _fields_ = [('n', c_i64), ('sig', c_vp), ('weights', c_vp), ('cls', c_vp), ('length', c_vp), ('td_bound', c_vp),
                ('gc', c_vp), ('gc_mbin', c_vp), ('gc_lbin', c_vp), ('cov', c_vp), ('cov_lbin', c_vp),
                ('row0', c_i64), ('row1', c_i64), ('rank', c_i32), ('world', c_i32), ('margin', ctypes.c_float)]


class PruneBuffers(ctypes.Structure):
    _fields_ = [('sig_sorted', c_vp), ('length', c_vp), ('td_bound', c_vp), ('gc', c_vp), ('gc_mbin', c_vp),
                ('gc_lbin', c_vp), ('cov', c_vp), ('cov_lbin', c_vp), ('col_id', c_vp), ('p_sorted', c_vp),
                ('item_order', c_vp), ('tile_off', c_vp), ('tile_list', c_vp), ('tile_count', c_vp), ('mine', c_vp),
                ('info', c_vp), ('scratch', c_vp), ('scratch_bytes', c_i64)]


class KnnParams(ctypes.Structure):
    _fields_ = [('k', c_i32), ('topk_filter', c_i32), ('prune', c_i32), ('n_classes', c_i32), ('margin', ctypes.c_float),
                ('rank', c_i32), ('world', c_i32)]


class KnnOutput(ctypes.Structure):
    _fields_ = [('cap_rows', c_i64), ('row_ids', c_vp), ('knn_idx', c_vp), ('knn_dist', c_vp), ('count', c_vp),
                ('neighbour_bp', c_vp), ('n_rows', c_i64), ('tiles_visited', c_i64), ('tiles_total', c_i64), ('tiles_second', c_i64), ('row_tiles_second', c_i64),
                ('pruned', c_i32), ('launches', c_i32), ('ms_plan', ctypes.c_float), ('ms_sweep', ctypes.c_float),
                ('ms_second', ctypes.c_float)]


class MergeInput(ctypes.Structure):
    _fields_ = [('search', RefineInput), ('own', c_vp)]


class UnbinnedInput(ctypes.Structure):
    _fields_ = [('search', RefineInput), ('group', c_vp), ('n_groups', c_i64)]


class RefineOutput(ctypes.Structure):
    _fields_ = [('closest', c_vp), ('min_dist', c_vp), ('assigned', c_vp), ('within', c_vp), ('dist', c_vp)]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    'dbb_last_error': (ctypes.c_char_p, []),
    'dbb_version': (ctypes.c_int, []),
    'dbb_device_info': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    'dbb_ctx_create': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    'dbb_ctx_destroy': (ctypes.c_int, [c_vp]),
    'dbb_kmer_lut': (ctypes.c_int, [c_vp]),
    'dbb_kmer_names': (ctypes.c_int, [c_vp]),
    'dbb_packed_blocks': (c_i64, [c_i64]),
    'dbb_block_offsets': (ctypes.c_int, [c_vp, c_i64, c_vp]),
    'dbb_pack_ascii_dev': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    'dbb_pack_ascii_host': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    'dbb_pack_ranges_dev': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'dbb_split_scaffolds_dev': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'dbb_base_count_dev': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    'dbb_count_plan_create': (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(c_vp)]),
    'dbb_count_plan_destroy': (ctypes.c_int, [c_vp]),
    'dbb_count_plan_launches': (ctypes.c_int, [c_vp]),
    'dbb_count_class_limits': (ctypes.c_int, [c_vp]),
    'dbb_tetra_count_dev': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    'dbb_tetra_signatures_host': (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    'dbb_tetra_signatures_host_ptrs': (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    'dbb_host_upload_stats': (ctypes.c_int, [c_vp, c_vp, c_vp]),
    'dbb_pairs_dev': (ctypes.c_int, [ctypes.POINTER(PairsInput), c_i64, c_i64, ctypes.POINTER(PairsOutput), c_vp]),
    'dbb_pairs_scratch_bytes': (c_i64, [c_i64, c_i64, c_i32]),
    'dbb_pairs_launches': (ctypes.c_int, [c_i64, c_i64]),
    'dbb_pairs_host': (ctypes.c_int, [ctypes.POINTER(PairsInput), c_i64, c_i64, ctypes.POINTER(PairsOutput)]),
    'dbb_prune_scratch_bytes': (c_i64, [c_i64, c_i64, c_i32]),
    'dbb_prune_plan_dev': (ctypes.c_int, [ctypes.POINTER(PruneInput), ctypes.POINTER(PruneBuffers), c_vp]),
    'dbb_prune_second_dev': (ctypes.c_int, [ctypes.POINTER(PruneInput), ctypes.POINTER(PruneBuffers), c_vp, c_i64, c_vp]),
    'dbb_prune_merge_dev': (ctypes.c_int, [ctypes.POINTER(PruneBuffers), c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'dbb_knn_dev': (ctypes.c_int, [c_vp, ctypes.POINTER(PairsInput), ctypes.POINTER(KnnParams), ctypes.POINTER(KnnOutput), c_vp]),
    'dbb_knn_host': (ctypes.c_int, [c_vp, ctypes.POINTER(PairsInput), ctypes.POINTER(KnnParams), ctypes.POINTER(KnnOutput)]),
    'dbb_refine_dev': (ctypes.c_int, [ctypes.POINTER(RefineInput), ctypes.POINTER(RefineOutput), c_vp]),
    'dbb_refine_host': (ctypes.c_int, [ctypes.POINTER(RefineInput), ctypes.POINTER(RefineOutput)]),
    'dbb_merge_table_dev': (ctypes.c_int, [ctypes.POINTER(MergeInput), c_vp, c_vp, c_vp]),
    'dbb_merge_table_host': (ctypes.c_int, [ctypes.POINTER(MergeInput), c_vp, c_vp]),
    'dbb_unbinned_table_dev': (ctypes.c_int, [ctypes.POINTER(UnbinnedInput), c_vp, c_vp]),
    'dbb_unbinned_table_host': (ctypes.c_int, [ctypes.POINTER(UnbinnedInput), c_vp]),
    'dbb_greedy_host': (ctypes.c_int, [ctypes.POINTER(GreedyInput), c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'dbb_pca_gram_host': (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    'dbb_pca_project_host': (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp]),
    'dbb_pad_signatures_dev': (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp]),
    'dbb_synth_ascii_dev': (ctypes.c_int, [ctypes.c_uint64, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                           c_vp, c_vp, c_vp]),
}

_LIB = None


def lib():
    """Load the shared library once per process (lazily: reference callers fork after constructing
    GenomicSignatures, preprocess.py:395-413, so nothing CUDA may happen at import/__init__ time)."""
    global _LIB
    if _LIB is None:
        path = os.environ.get('DBB_LIB_PATH', LIB_PATH)             # override: A/B runs of two builds
        if not os.path.exists(path):
            raise DbbError('%s is not built (run `python -m dbb_b200.build`); dbb_b200 has no CPU fallback' % path)
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().dbb_last_error()
        raise DbbError('libdbbcuda error %d: %s' % (rc, msg.decode('utf-8', 'replace') if msg else '?'))


def ptr(a):
    """void* of a numpy array (C-contiguous) or None."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(c_vp)


def kmer_table():
    """(names[136], lut256) straight from the library (no GPU needed)."""
    L = lib()
    names = ctypes.create_string_buffer(NUM_KMERS * 5)
    check(L.dbb_kmer_names(names))
    lut = np.zeros(256, dtype=np.uint8)
    check(L.dbb_kmer_lut(ptr(lut)))
    raw = names.raw
    return [raw[5 * i:5 * i + 4].decode('ascii') for i in range(NUM_KMERS)], lut
