"""Host-side packer (csrc/hostpack.cpp) and the two upload routes of dbb_tetra_signatures_host (csrc/host_api.cu)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dbb_b200 import _lib
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))


def _probe(mode, **env):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(HERE, '_hostpack_probe.py'), mode], env=e, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    return out.stdout.strip().splitlines()[-1].split()


def _reference_planes(flat, off, blk):
    """The packed layout of include/dbb_cuda.h written out base by base (numpy)."""
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(b'ACGT'):
        lut[ch] = i
    nb = int(blk[-1])
    codes = np.zeros(nb * 4, dtype=np.uint32)
    valid = np.zeros(nb, dtype=np.uint64)
    for s in range(off.shape[0] - 1):
        x = lut[flat[off[s]:off[s + 1]]]
        pos = np.nonzero(x != 255)[0]
        b = blk[s] + pos // 64
        p = pos % 64
        np.bitwise_or.at(codes, 4 * b + p // 16, (x[pos].astype(np.uint32) << (30 - 2 * (p % 16)).astype(np.uint32)))
        np.bitwise_or.at(valid, b, np.uint64(1) << (63 - p).astype(np.uint64))
    return codes, valid


def test_pack_ascii_host_matches_the_layout_and_both_code_paths_agree():
    rng = np.random.RandomState(3)
    lens = np.array([0, 1, 3, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4097, 70001], dtype=np.int64)
    alphabet = np.frombuffer(b'ACGTACGTACGTACGTNacgtnRYU*-\x00\xff', dtype=np.uint8)
    flat = alphabet[rng.randint(0, len(alphabet), int(lens.sum()))].copy()
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    blk = np.zeros(len(lens) + 1, dtype=np.int64)
    blk[1:] = np.cumsum((lens + 63) // 64)
    codes = np.full(int(blk[-1]) * 4, 0xdeadbeef, dtype=np.uint32)
    valid = np.full(int(blk[-1]), 0x5555, dtype=np.uint64)
    L = _lib.lib()
    _lib.check(L.dbb_pack_ascii_host(_lib.ptr(flat), _lib.ptr(off), _lib.ptr(blk), len(lens), _lib.ptr(codes), _lib.ptr(valid)))
    rc, rv = _reference_planes(flat, off, blk)
    assert np.array_equal(codes, rc) and np.array_equal(valid, rv)
    assert L.dbb_pack_ascii_host(None, _lib.ptr(off), _lib.ptr(blk), len(lens), _lib.ptr(codes), _lib.ptr(valid)) == -1          # DBB_ERR_INVALID
    # vector path (when the CPU has AVX2 + BMI2) against the table loop
    assert _probe('planes')[0] == _probe('planes', DBB_HOST_PACK_SCALAR='1')[0]


@pytest.mark.gpu
def test_host_planes_equal_the_pack_kernel():
    import torch
    from dbb_b200 import device
    rng = np.random.RandomState(11)
    lens = rng.randint(0, 9000, 500).astype(np.int64)
    seqs = [util.random_dna(rng, int(l), 0.02 if i % 5 == 0 else 0.0) for i, l in enumerate(lens)]
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    flat = np.concatenate(seqs)
    packed = device.pack_ascii(torch.from_numpy(flat).cuda(), off)
    nb = packed.total_blocks
    codes = np.zeros(nb * 4, dtype=np.uint32)
    valid = np.zeros(nb, dtype=np.uint64)
    _lib.check(_lib.lib().dbb_pack_ascii_host(_lib.ptr(flat), _lib.ptr(off), _lib.ptr(packed.blk_off), len(lens), _lib.ptr(codes), _lib.ptr(valid)))
    assert np.array_equal(packed.codes[:nb * 4].cpu().numpy().view(np.uint32), codes)
    assert np.array_equal(packed.valid[:nb].cpu().numpy().view(np.uint64), valid)


@pytest.mark.gpu
def test_signatures_host_is_the_same_by_either_upload_route():
    """ASCII only, host-packed only and the mixed default give identical counts and frequencies (16 MB batches of a 40 MB
    input, so every batch takes the hybrid path)."""
    env = dict(DBB_HOST_BATCH_MB='16')
    ascii_only = _probe('signatures', DBB_HOST_PACK='0', **env)
    packed_only = _probe('signatures', DBB_HOST_PACK='all', DBB_HOST_PACK_THREADS='4', **env)
    mixed = _probe('signatures', DBB_HOST_PACK_THREADS='3', **env)
    assert ascii_only[0] == packed_only[0] == mixed[0]
    assert int(ascii_only[2]) == 0 and int(ascii_only[1]) > 0          # nothing went up as planes
    assert int(packed_only[1]) == 0 and int(packed_only[2]) > 0        # nothing went up as ASCII
    assert int(mixed[1]) + int(mixed[2]) > 0


@pytest.mark.gpu
def test_class_call_on_a_list_of_separate_buffers_needs_no_concatenation():
    """GenomicSignatures.signatures(list) -> dbb_tetra_signatures_host_ptrs: str, bytes, uint8 arrays, empty and one-base
    sequences, each at its own address; same counts as the back-to-back entry point and as the oracle."""
    from dbb_b200.genomicSignatures import GenomicSignatures
    from oracle import dbb_oracle as O
    rng = np.random.RandomState(21)
    lens = [0, 1, 3, 4, 63, 64, 65, 700, 5000, 130000] + rng.randint(0, 40000, 60).tolist()
    arrs = [util.random_dna(rng, int(l), 0.02 if i % 4 == 0 else 0.0) for i, l in enumerate(lens)]
    mixed = [a if i % 3 == 0 else (a.tobytes() if i % 3 == 1 else a.tobytes().decode('latin-1')) for i, a in enumerate(arrs)]
    gs = GenomicSignatures(4, 1)
    counts, freq = gs.signatures(mixed)
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    c2, f2 = gs.signatures((np.concatenate(arrs), off))
    assert np.array_equal(counts, c2)
    assert np.array_equal(freq, f2, equal_nan=True)
    for i in (0, 1, 3, 4, 7, 9, 20):
        assert np.array_equal(counts[i].astype(np.int64), O.seq_counts_np(arrs[i]))
    # a batch large enough for the packer threads (several MB) through the same call
    big = [util.random_dna(rng, 50000) for _ in range(120)]
    cb, _ = gs.signatures(big, want_freq=False)
    offb = np.arange(121, dtype=np.int64) * 50000
    cb2, _ = gs.signatures((np.concatenate(big), offb), want_freq=False)
    assert np.array_equal(cb, cb2)
