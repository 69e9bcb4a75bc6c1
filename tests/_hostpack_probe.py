"""Helper of test_hostpack.py: prints digests of what the host packer / the host entry point produce under the
environment it was started with (the switches DBB_HOST_PACK, DBB_HOST_PACK_SCALAR are read once per process)."""
import ctypes
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbb_b200 import _lib  # noqa: E402
from tests import util  # noqa: E402


def workload(seed, n_seq, max_len):
    rng = np.random.RandomState(seed)
    lens = rng.randint(0, max_len, n_seq).astype(np.int64)
    lens[:8] = [0, 1, 3, 63, 64, 65, 127, 4097]
    seqs = [util.random_dna(rng, int(l), 0.01 if i % 7 == 0 else 0.0) for i, l in enumerate(lens)]
    off = np.zeros(n_seq + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    return np.concatenate(seqs), off


def main():
    mode = sys.argv[1]
    L = _lib.lib()
    if mode == 'planes':
        flat, off = workload(5, 300, 20000)
        blk = np.zeros(off.shape[0], dtype=np.int64)
        _lib.check(L.dbb_block_offsets(_lib.ptr(np.diff(off)), off.shape[0] - 1, _lib.ptr(blk)))
        codes = np.zeros(int(blk[-1]) * 4, dtype=np.uint32)
        valid = np.zeros(int(blk[-1]), dtype=np.uint64)
        _lib.check(L.dbb_pack_ascii_host(_lib.ptr(flat), _lib.ptr(off), _lib.ptr(blk), off.shape[0] - 1, _lib.ptr(codes), _lib.ptr(valid)))
        print(hashlib.sha256(codes.tobytes() + valid.tobytes()).hexdigest())
    else:                                      # 'signatures': the host entry point on ~40 MB in 16 MB batches
        flat, off = workload(9, 4000, 20000)
        n = off.shape[0] - 1
        counts = np.zeros((n, 136), dtype=np.uint32)
        freq = np.zeros((n, 136), dtype=np.float32)
        _lib.check(L.dbb_tetra_signatures_host(_lib.ptr(flat), _lib.ptr(off), n, _lib.ptr(counts), _lib.ptr(freq)))
        a, p, t = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
        _lib.check(L.dbb_host_upload_stats(ctypes.byref(a), ctypes.byref(p), ctypes.byref(t)))
        print(hashlib.sha256(counts.tobytes() + freq.tobytes()).hexdigest(), a.value, p.value, t.value)


if __name__ == '__main__':
    main()
