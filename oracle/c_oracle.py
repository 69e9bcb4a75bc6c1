"""ctypes loader of oracle/_build/libdbb_oracle.so (the C restatement; test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libdbb_oracle.so')
_L = None


def build():
    subprocess.run(['make', '-C', HERE, '--no-print-directory'], check=True, capture_output=True)
    return LIB


def lib():
    global _L
    if _L is None:
        if not os.path.exists(LIB):
            build()
        _L = ctypes.CDLL(LIB)
    return _L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def kmer_table():
    t = np.zeros(256, dtype=np.uint8)
    names = ctypes.create_string_buffer(680)
    n = lib().oracle_kmer_table(_p(t), names)
    return t, [names.raw[5 * i:5 * i + 4].decode() for i in range(n)]


def counts_batch(flat, off):
    flat = np.ascontiguousarray(flat, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = off.shape[0] - 1
    out = np.zeros((n, 136), dtype=np.int64)
    lib().oracle_counts_batch(_p(flat), _p(off), ctypes.c_int64(n), _p(out))
    return out


def td_rows(sig, length, td_bound, rows, want_dist=True, want_mask=True):
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    length = np.ascontiguousarray(length, dtype=np.int64)
    td_bound = np.ascontiguousarray(td_bound, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    n = sig.shape[0]
    dist = np.empty((rows.shape[0], n), dtype=np.float64) if want_dist else None
    mask = np.empty((rows.shape[0], n), dtype=np.uint8) if want_mask else None
    lib().oracle_td_rows(_p(sig), _p(length), _p(td_bound), ctypes.c_int64(n), _p(rows), ctypes.c_int64(rows.shape[0]),
                         _p(dist) if want_dist else None, _p(mask) if want_mask else None)
    return dist, mask
