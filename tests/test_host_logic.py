"""CPU: host-side logic -- per-scaffold table resolution, the Distributions mirror, lazy neighbour rows,
the seeded generator, LPT sharding and the all-gather plumbing (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dists():
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    return load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=7)


def test_pair_tables_equal_oracle_tables():
    from dbb_b200.distributions import PairTables
    td, gcd, cvd = _dists()
    rng = np.random.RandomState(4)
    length = np.concatenate([rng.randint(500, 1200000, 500), [550, 12500, 150000, 500, 1000000, 5000000]])   # incl. midpoints
    GC = rng.uniform(0.1, 0.9, length.size); cov = rng.lognormal(3, 1, length.size)
    for fixed in (None, 10000):
        for per in (80, 99, 37):
            t = PairTables(length, GC=GC, coverage=cov, tdDist=td, tdDistPer=per, gcDist=gcd, gcDistPer=per, covDist=cvd,
                           covDistPer=per, fixedSeqLen=fixed)
            o = O.DistTables(gcd, per, td, per, cvd, per)
            look = length if fixed is None else np.full(length.shape, fixed)
            assert np.array_equal(t.td_bound64, o.tdBound[O.nearest_index_sorted(o.tdLens, look)])
            assert np.array_equal(t.gc_mbin, O.nearest_index_sorted(o.gcMeans, GC))
            assert np.array_equal(t.gc_lbin, O.nearest_index_sorted(o.gcLens, look))
            assert np.array_equal(t.cov_lbin, O.nearest_index_sorted(o.covLens, look))
            assert np.array_equal(t.gc_lo, o.gcLo) and np.array_equal(t.gc_hi, o.gcHi)
            assert np.array_equal(t.cov_lo, o.covLo) and np.array_equal(t.cov_hi, o.covHi)
    # frozen tie rule: lower key wins at an exact midpoint
    assert PairTables(np.array([12500]), tdDist=td, tdDistPer=80).td_bound64[0] == td[10000][80.0]


def test_distributions_mirror_equals_oracle_on_golden_contigs():
    from dbb_b200.distributions import Distributions
    td, gcd, cvd = _dists()
    nb = util.golden_npz('ref_neighbours.npz')
    cs = [O.Contig('c', int(nb['length'][i]), float(nb['GC'][i]), float(nb['coverage'][i]), nb['tetraSig'][i]) for i in range(20)]
    a = Distributions(gcd, 80, td, 80, cvd, 80)
    b = O.Distributions(gcd, 80, td, 80, cvd, 80)
    assert (a.tdKey, a.gcLowerBoundKey, a.gcUpperBoundKey, a.covLowerBoundKey, a.covUpperBoundKey) == (80.0, 10.0, 90.0, 10.0, 90.0)
    for fixed in (None, 10000):
        for x in cs:
            for y in cs:
                assert a.boundsDistGC(x, y, fixed) == b.boundsDistGC(x, y, fixed)
                assert a.withinDistGC(x, y, fixed) == b.withinDistGC(x, y, fixed)
                assert a.boundsDistCov(x, y, fixed) == b.boundsDistCov(x, y, fixed)
                assert a.withinDistCov(x, y, fixed) == b.withinDistCov(x, y, fixed)
            assert a.boundDistTD(x, fixed) == b.boundDistTD(x, fixed)
            assert a.witinDistTD(0.1, x, fixed) == b.witinDistTD(0.1, x, fixed)


def test_lazy_neighbour_rows_behave_like_dense_rows():
    from dbb_b200._pairs import NeighbourRows
    rng = np.random.RandomState(1)
    n = 77
    dense = (rng.uniform(size=(n, n)) < 0.2).astype(np.float64)
    bits = np.packbits(dense.astype(np.uint8), axis=1, bitorder='little')
    bits = np.ascontiguousarray(np.pad(bits, ((0, 0), (0, (-bits.shape[1]) % 4)))).view(np.uint32)
    rows = NeighbourRows(bits, n)
    for i in (0, 5, n - 1):
        r = rows.row(i)
        assert np.array_equal(rows.dense_row(i), dense[i]) and rows.dense_row(i).dtype == np.float64
        assert np.array_equal(np.where(r == 1)[0], np.where(dense[i] == 1)[0])       # greedy.py:409-411
        assert np.array_equal(np.asarray(r), dense[i]) and len(r) == n and r[3] == dense[i][3]
    assert np.array_equal(rows.counts(), dense.sum(axis=1))


def test_generator_is_deterministic_and_shaped():
    from dbb_b200 import synthetic
    m1 = synthetic.make_metagenome(seed=5, n_scaffolds=40, min_len=1000, max_len=9000, n_genomes=3, frac_n=0.2, frac_lower=0.2)
    m2 = synthetic.make_metagenome(seed=5, n_scaffolds=40, min_len=1000, max_len=9000, n_genomes=3, frac_n=0.2, frac_lower=0.2)
    s1, s2 = synthetic.generate_sequences(m1), synthetic.generate_sequences(m2, block_bp=20000)
    assert all(np.array_equal(a, b) for a, b in zip(s1, s2))
    assert [len(s) for s in s1] == m1.length.tolist()
    assert not set(m1.length.tolist()) & synthetic._MIDPOINTS
    sub = synthetic.generate_sequences(m1, [7, 3])
    assert np.array_equal(sub[0], s1[7]) and np.array_equal(sub[1], s1[3])
    alphabet = set(np.concatenate(s1).tolist())
    assert alphabet <= set(b'ACGTNacgt') and ord('N') in alphabet
    # same-genome scaffolds are closer in tetranucleotide space than different-genome ones
    sig = np.array([O.seq_signature_np(s) for s in s1])
    d = np.abs(sig[:, None, :] - sig[None, :, :]).sum(axis=2)
    same = m1.genome[:, None] == m1.genome[None, :]
    assert np.nanmean(d[same & ~np.eye(40, dtype=bool)]) < 0.7 * np.nanmean(d[~same])
    flat, off = synthetic.concat(s1)
    assert off[-1] == flat.size == m1.total_bp


def test_lpt_partition_and_row_blocks():
    from dbb_b200 import parallel
    rng = np.random.RandomState(2)
    lengths = np.concatenate([rng.randint(1000, 2001, 5000), [5000000] * 8])
    for world in (1, 2, 4, 8):
        parts = parallel.lpt_partition(lengths, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(lengths.size))
        assert parallel.imbalance(lengths, parts) < 1.02
        if world == 8:
            assert all((lengths[p] == 5000000).sum() == 1 for p in parts)          # one 5 Mb scaffold each
        order, inv = parallel.gathered_order(parts)
        assert np.array_equal(order[inv], np.arange(lengths.size))
        blocks = [parallel.row_block(lengths.size, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == lengths.size
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_all_gather_rows_gloo_world2():
    script = os.path.join(ROOT, 'tests', 'gloo_worker.py')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29611', PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r), WORLD_SIZE='2'),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), '\n'.join(outs)
    assert all('gloo ok' in o for o in outs)


def test_merge_and_unbinned_host_sides_reproduce_the_reference_files_from_their_own_tables():
    """Host logic of the N5 / N6 mirrors without a GPU: the running bin / scaffold statistics and the report formatting
    of dbb_b200.merge / dbb_b200.unbinned, fed with the integer tables parsed back from the reference's own output
    files, reproduce those files byte for byte."""
    import re
    from dbb_b200.merge import Merge
    from dbb_b200.unbinned import Scaffold, Unbinned
    from tests.test_oracle_golden import merge_case, unbinned_case
    for tag in 'abc':
        g, contigs, pers, _, _, _ = merge_case(tag)
        text = str(g[tag + '_merge_tsv'])
        bins = Merge.binStatistics([c for c in contigs if c.binId != -1])
        pos = dict((b.binId, k) for k, b in enumerate(bins))
        table = np.zeros((len(bins), len(bins), 6), dtype=np.int32)
        for row in text.strip().split('\n')[1:]:
            f = row.split('\t')
            table[pos[f[0]], pos[f[3]]] = [int(x) for x in f[7:13]]
        assert Merge().formatTable(bins, table) == text, tag
    for tag in 'ab':
        g, contigs, pr, _, _, _ = unbinned_case(tag)
        text = str(g[tag + '_unbinned_tsv'])
        minScaffoldLen, minContigLen, allScaffolds = pr[3], pr[4], bool(pr[5])
        scaffolds, binned = {}, []
        for c in contigs:                                              # unbinned.py:72-107
            m = re.search('_c[0-9]*$', c.contigId)
            sid = c.contigId[:m.start()] if m else c.contigId
            if c.length < minContigLen:
                continue
            if c.binId != -1:
                binned.append(c)
            if c.binId == -1 or allScaffolds:
                scaffolds.setdefault(sid, Scaffold(sid)).add(c)
        scaffolds = list(scaffolds.values())
        bins = Merge.binStatistics(binned)
        spos = dict((s.scaffoldId, k) for k, s in enumerate(scaffolds))
        bpos = dict((b.binId, k) for k, b in enumerate(bins))
        table = np.zeros((len(scaffolds), len(bins), 7), dtype=np.int32)
        for row in text.strip().split('\n')[1:]:
            f = row.split('\t')
            s = scaffolds[spos[f[0]]]
            score = int(round(float(f[14]) * len(s.contigs)))
            table[spos[f[0]], bpos[f[4]]] = [int(x) for x in f[8:14]] + [score]
        assert Unbinned().formatTable(scaffolds, bins, table, minScaffoldLen) == text, tag


def test_fasta_reader_parses_like_the_reference(tmp_path):
    """``GenomicSignatures.calculate`` reads FASTA exactly as seqUtils.readFasta does (golden: the reference's own function
    text, tests/golden/make_golden_fasta.py): ids up to the first space (tabs kept), the last character of every line
    dropped (a CRLF's '\\r' stays, an unterminated last line loses a base), a repeated id replaces the earlier record."""
    import io
    from dbb_b200.genomicSignatures import _read_fasta
    g = util.golden_json('ref_fasta.json')
    path = str(tmp_path / 'in.fasta')
    with io.open(path, 'w', newline='') as f:
        f.write(g['fasta'])
    got = dict(_read_fasta(path))
    assert got == g['records']
    assert '\r' in got['seq2'] and got['last'].endswith('AG') and got['seq1'].startswith('TTTT') and '' in got
    with pytest.raises(SystemExit):
        _read_fasta(str(tmp_path / 'missing.fasta'))
