"""CPU: the C-ABI library loads without a GPU, exports every symbol include/dbb_cuda.h declares, its
GPU-free entry points work, and the compute entry points FAIL LOUDLY (no CPU fallback) when no CUDA device
is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from dbb_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'dbb_cuda.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dbb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run python -m dbb_b200.build'
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), 'missing export ' + n
    assert sorted(_lib.SIGNATURES) == names, 'ctypes table and header disagree'


def test_table_entry_points_without_gpu():
    from oracle import dbb_oracle as O
    names, lut = _lib.kmer_table()
    assert names == O.KMER_COLS
    assert np.array_equal(lut, O.code_lut256())
    L = _lib.lib()
    assert L.dbb_version() >= 100
    assert [L.dbb_packed_blocks(x) for x in (0, 1, 63, 64, 65, 128)] == [0, 1, 1, 1, 2, 2]
    lens = np.array([0, 1, 64, 65, 1000], dtype=np.int64)
    off = np.zeros(6, dtype=np.int64)
    _lib.check(L.dbb_block_offsets(_lib.ptr(lens), 5, _lib.ptr(off)))
    assert off.tolist() == [0, 0, 1, 2, 4, 20]
    bad = np.array([5, -1], dtype=np.int64)
    assert L.dbb_block_offsets(_lib.ptr(bad), 2, _lib.ptr(off)) == -1 and b'negative' in L.dbb_last_error()


def test_dropin_surface_and_k_restriction():
    from dbb_b200.genomicSignatures import GenomicSignatures
    import pickle
    g = GenomicSignatures(4, 8)
    assert g.K == 4 and g.totalThreads == 8 and g.numKmers() == 136 and g.canonicalKmerOrder()[0] == 'AAAA'
    assert 'ACGT'.translate(g.compl) == 'TGCA' and len(g.kmerToCanonicalIndex) == 256
    assert g.kmerToCanonicalIndex['TTTT'] == g.kmerToCanonicalIndex['AAAA'] == 0
    pickle.loads(pickle.dumps(g))                                   # preprocess.py:395-413 forks/pickles it
    a, b = np.full(136, 1 / 136.0), np.zeros(136); b[0] = 1.0
    assert abs(g.distance(a, b) - (2 - 2 / 136.0)) < 1e-12
    with pytest.raises(NotImplementedError):
        GenomicSignatures(5, 1)
    for mod, cls in (('tdNeighbours', 'TdNeighbours'), ('gcNeighbours', 'GcNeighbours'), ('coverageNeighbours', 'CoverageNeighbours')):
        m = __import__('dbb_b200.' + mod, fromlist=[cls])
        obj = getattr(m, cls)(10000)
        assert obj.fixedSeqLen == 10000 and hasattr(obj, 'run') and hasattr(obj, 'logger')


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from dbb_b200.genomicSignatures import GenomicSignatures
    with pytest.raises(_lib.DbbError):
        GenomicSignatures(4, 1).seqSignature('ACGTACGT')
    from dbb_b200 import _pairs
    from dbb_b200.distributions import PairTables
    with pytest.raises(_lib.DbbError):
        _pairs.run_pairs_host(np.zeros((4, 136), np.float32), PairTables(np.array([1000, 2000, 3000, 4000])), k=2)
    # the callers around the path (N1, N3, N5, N6): same rule
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.merge import Contig, Merge
    from dbb_b200.unbinned import Scaffold, Unbinned
    from dbb_b200.refineBins import RefineBins, coreStatistics
    from dbb_b200.greedy import Greedy
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    cs = []
    for i in range(4):
        c = Contig('c%d' % i, 5000 + i, 0.5, 10.0, np.full(136, 1 / 136.0))
        c.binId = 'b%d' % (i % 2)
        cs.append(c)
    args = (gcDist, 99, td, 99, covDist, 99)
    with pytest.raises(_lib.DbbError):
        Merge().mergeTable(cs, *args)
    sc = Scaffold('s')
    sc.add(cs[0])
    with pytest.raises(_lib.DbbError):
        Unbinned().unbinnedTable([sc], cs[1:], *args)
    with pytest.raises(_lib.DbbError):
        RefineBins().closestCores(cs[:1], coreStatistics(cs[1:]), *args)
    with pytest.raises(_lib.DbbError):
        Greedy().greedy(cs, 1000, 99, gcDist, td, covDist)
    # host-side argument checks of the C ABI answer before any CUDA call
    L = _lib.lib()
    assert L.dbb_merge_table_host(None, None, None) != 0 and b'NULL' in L.dbb_last_error()
    assert L.dbb_unbinned_table_host(None, None) != 0
    assert L.dbb_prune_plan_dev(None, None, None) != 0 and L.dbb_prune_merge_dev(None, 0, 0, None, None, None, None, None) != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'dbb_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                assert 'oracle' not in open(os.path.join(dirpath, f)).read(), f


@pytest.mark.gpu
def test_knn_host_from_plain_numpy_and_contexts():
    """The default kNN path through the C ABI alone: numpy buffers in, numpy buffers out, no torch (``dbb_knn_host``), on
    the default context and on an explicit one; a context refuses to run with another device current."""
    import ctypes
    from dbb_b200 import _lib, _pairs, synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.distributions import PairTables
    from dbb_b200.genomicSignatures import GenomicSignatures
    from oracle import dbb_oracle as O, checks
    meta = synthetic.make_metagenome(seed=71, n_scaffolds=1500, min_len=1000, max_len=30000, n_genomes=10)
    counts, freq = GenomicSignatures(4, 1).signatures(synthetic.generate_sequences(meta))
    td = load_td_dist()
    tables = PairTables(meta.length, tdDist=td, tdDistPer=80)
    a = _pairs.run_knn_host(freq, tables, k=20, prune=1)
    b = _pairs.run_knn_host(freq, tables, k=20, prune=0)
    assert a['pruned'] and not b['pruned'] and 0 < a['tiles_visited'] <= a['tiles_total']   # (12 tiles: the unfiltered top-k visits them all)
    for key in ('idx', 'dist', 'count', 'neighbour_bp', 'row_ids'):
        assert np.array_equal(a[key], b[key]), key
    rows = list(range(0, meta.n_scaffolds, 25))
    checks.check_knn_sample(rows, a['idx'][rows], a['dist'][rows], a['count'][rows], a['neighbour_bp'][rows], freq, meta.length, 20,
                            O.DistTables(tdDist=td, tdDistPer=80))
    L = _lib.lib()
    ctx = ctypes.c_void_p()
    _lib.check(L.dbb_ctx_create(-1, ctypes.byref(ctx)))
    try:
        c = _pairs.run_knn_host(freq, tables, k=20, prune=1, ctx=ctx)
        for key in ('idx', 'dist', 'count', 'neighbour_bp'):
            assert np.array_equal(a[key], c[key]), key
        # two ranks through the host entry point: disjoint rows, together everything
        parts = [_pairs.run_knn_host(freq, tables, k=20, prune=1, rank=r, world=2, ctx=ctx) for r in range(2)]
        ids = np.concatenate([p_['row_ids'] for p_ in parts])
        assert np.array_equal(np.sort(ids), np.arange(meta.n_scaffolds))
        for p_ in parts:
            assert np.array_equal(p_['idx'], a['idx'][p_['row_ids']]) and np.array_equal(p_['count'], a['count'][p_['row_ids']])
    finally:
        _lib.check(L.dbb_ctx_destroy(ctx))
    assert L.dbb_ctx_create(10 ** 6, ctypes.byref(ctx)) != 0 and b'no such device' in L.dbb_last_error()
