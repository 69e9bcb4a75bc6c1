"""N5 / N6 rows: the bin-to-bin merge table (merge.py) and the statistics of unbinned scaffolds (unbinned.py) against the
reference's own output files (golden) and the oracle's literal loops."""
import ast
import os

import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests.test_oracle_golden import merge_case, unbinned_case

pytestmark = pytest.mark.gpu


def _write_inputs(d, contigs, gcDist, tdDist, covDist):
    from dbb_b200.genomicSignatures import GenomicSignatures
    with open(os.path.join(d, 'contigs.seq_stats.tsv'), 'w') as f:
        f.write('Contig Id\tLength\tGC\tCoverage\n')
        for c in contigs:
            f.write('%s\t%d\t%r\t%r\n' % (c.contigId, c.length, c.GC, c.coverage))
    with open(os.path.join(d, 'contigs.tetra.tsv'), 'w') as f:
        f.write('Contig Id\t' + '\t'.join(GenomicSignatures(4, 1).canonicalKmerOrder()) + '\n')
        for c in contigs:
            f.write(c.contigId + '\t' + '\t'.join(repr(float(x)) for x in c.tetraSig) + '\n')
    with open(os.path.join(d, 'bins.tsv'), 'w') as f:
        f.write('# parameters\nContig Id\tBin Id\tLength\n')
        for c in contigs:
            f.write('%s\t%s\t%d\n' % (c.contigId, 'unbinned' if c.binId == -1 else c.binId, c.length))
    data = os.path.join(d, 'data')
    os.makedirs(data)
    plain = lambda t: dict((k, plain(v)) for k, v in t.items()) if isinstance(t, dict) else t
    for name, t in (('gc_dist.txt', gcDist), ('td_dist.txt', tdDist)):
        with open(os.path.join(data, name), 'w') as f:
            f.write(repr(plain(t)))
    with open(os.path.join(d, 'coverage_dist.txt'), 'w') as f:
        f.write(repr(plain(covDist)))
    return data


def test_merge_run_writes_the_reference_file(tmp_path):
    """Merge.run through real files (the reference's call, merge.py:35) == the file the reference's own Merge.run wrote."""
    from dbb_b200.merge import Merge
    for tag in 'abc':
        g, contigs, pers, gcDist, td, covDist = merge_case(tag)
        d = str(tmp_path / tag)
        os.makedirs(d)
        data = _write_inputs(d, contigs, gcDist, td, covDist)
        assert ast.literal_eval(open(os.path.join(data, 'td_dist.txt')).read()) == dict((k, dict(v)) for k, v in td.items())
        out = os.path.join(d, 'merge.tsv')
        Merge().run(d, os.path.join(d, 'bins.tsv'), pers[0], pers[1], pers[2], out, dataDir=data)
        assert open(out).read() == str(g[tag + '_merge_tsv']), tag
    with pytest.raises(SystemExit):
        Merge().run(str(tmp_path / 'missing'), 'x', 99, 99, 99, 'y')


def test_merge_table_random_instance_vs_literal_loop():
    """700 binned contigs, 5 genomes cut into 15 bins, mixed percentiles; per-contig closest bins and the whole table
    against the literal loop; a bin with a single contig and exact ties (two bins with identical statistics: the one
    visited first must win) are planted."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.genomicSignatures import GenomicSignatures
    from dbb_b200.merge import Merge
    n = 700
    meta = synthetic.make_metagenome(seed=51, n_scaffolds=n, min_len=1500, max_len=25000, n_genomes=5)
    seqs = synthetic.generate_sequences(meta)
    counts, _ = GenomicSignatures(4, 1).signatures(seqs, want_freq=False)
    sig = counts.astype(np.float64) / counts.sum(axis=1, keepdims=True)
    rng = np.random.RandomState(51)
    bin_of = meta.genome.astype(np.int64) * 3 + rng.randint(0, 3, size=n)
    moved = rng.rand(n) < 0.1
    bin_of[moved] = rng.randint(0, 15, size=int(moved.sum()))
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    contigs = []
    for i in range(n):
        c = O.Contig('c%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), sig[i])
        c.binId = 'bin%d' % bin_of[i]
        contigs.append(c)
    # bin15: one contig; bin16 / bin17: twins of each other (same single member statistics -> exact ties everywhere)
    for b, src in ((15, contigs[10]), (16, contigs[3]), (17, contigs[3])):
        c = O.Contig('x%d' % b, src.length, src.GC, src.coverage, src.tetraSig.copy())
        c.binId = 'bin%d' % b
        contigs.append(c)
    m = Merge()
    bins, table, closest = m.mergeTable(contigs, gcDist, 95, td, 90, covDist, 99, want_closest=True)
    dist = O.Distributions(gcDist, 95, td, 90, covDist, 99)
    assert m.formatTable(bins, table) == O.merge_table_literal(contigs, dist)
    assert len(bins) == 18 and np.all(table[np.arange(18), np.arange(18)] == 0)
    pos = dict((b.binId, k) for k, b in enumerate(bins))
    # ties: wherever one twin is the closest bin of a contig outside both, it is the twin visited first
    first, second = sorted((pos['bin16'], pos['bin17']))
    outside = np.array([pos[c.binId] not in (first, second) for c in contigs])
    assert not np.any(closest[outside] == second) and np.any(closest[outside] == first)
    # per-contig closest bins against the N1 oracle with the own bin removed
    cores = dict((k, b) for k, b in enumerate(bins))
    for u in range(0, len(contigs), 7):
        own = pos[contigs[u].binId]
        cb, _, _, _, _, _ = O.refine_worker_literal(contigs[u], dict((k, b) for k, b in cores.items() if k != own), dist)
        assert [(-1 if x is None else x) for x in cb] == closest[u].tolist()


def test_merge_table_degenerate_inputs():
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.merge import Merge
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    bins, table = Merge().mergeTable([], gcDist, 99, td, 99, covDist, 99)
    assert bins == [] and table.shape == (0, 0, 6)
    c = O.Contig('a', 5000, 0.5, 10.0, np.full(136, 1 / 136.0))
    c.binId = 'only'
    bins, table = Merge().mergeTable([c], gcDist, 99, td, 99, covDist, 99)
    assert len(bins) == 1 and table.tolist() == [[[0] * 6]]
    assert Merge().formatTable(bins, table) == O.MERGE_HEADER


def test_unbinned_run_writes_the_reference_file(tmp_path):
    """Unbinned.run through real files (unbinned.py:38) == the file the reference's own Unbinned.run wrote."""
    from dbb_b200.unbinned import Unbinned
    for tag in 'ab':
        g, contigs, pr, gcDist, td, covDist = unbinned_case(tag)
        d = str(tmp_path / tag)
        os.makedirs(d)
        data = _write_inputs(d, contigs, gcDist, td, covDist)
        out = os.path.join(d, 'unbinned.tsv')
        Unbinned().run(d, os.path.join(d, 'bins.tsv'), pr[0], pr[1], pr[2], pr[3], pr[4], bool(pr[5]), out, dataDir=data)
        assert open(out).read() == str(g[tag + '_unbinned_tsv']), tag


def test_unbinned_table_random_instance_vs_literal_loop():
    """900 contigs in scaffolds of 1-5 contigs, 4 genomes cut into 12 bins, a third unbinned; the whole file against the
    literal loop, for unbinned scaffolds only and for all scaffolds."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.genomicSignatures import GenomicSignatures
    from dbb_b200.unbinned import Unbinned
    n = 900
    meta = synthetic.make_metagenome(seed=71, n_scaffolds=n, min_len=1000, max_len=25000, n_genomes=4)
    seqs = synthetic.generate_sequences(meta)
    counts, _ = GenomicSignatures(4, 1).signatures(seqs, want_freq=False)
    sig = counts.astype(np.float64) / counts.sum(axis=1, keepdims=True)
    rng = np.random.RandomState(71)
    order = np.argsort(meta.genome, kind='stable')
    contigs, pos, sc = [None] * n, 0, 0
    while pos < n:
        members = order[pos:pos + int(rng.randint(1, 6))]
        unb = rng.rand() < 0.33
        for k, i in enumerate(members):
            c = O.Contig('s%d_c%d' % (sc, k) if len(members) > 1 else 's%d' % sc, int(meta.length[i]), float(meta.gc_obs[i]),
                         float(meta.coverage[i]), sig[i])
            c.binId = -1 if unb else 'bin%d' % (int(meta.genome[i]) * 3 + int(rng.randint(0, 3)))
            contigs[i] = c
        pos += len(members)
        sc += 1
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    dist = O.Distributions(gcDist, 95, td, 99, covDist, 99)
    import tempfile
    for allScaffolds, minS, minC in ((False, 5000, 1500), (True, 0, 0)):
        with tempfile.TemporaryDirectory() as d:
            data = _write_inputs(d, contigs, gcDist, td, covDist)
            out = os.path.join(d, 'u.tsv')
            Unbinned().run(d, os.path.join(d, 'bins.tsv'), 95, 99, 99, minS, minC, allScaffolds, out, dataDir=data)
            got = open(out).read()
        want = O.unbinned_table_literal(contigs, dist, minS, minC, allScaffolds)
        assert got == want, allScaffolds
        assert len(want.split('\n')) > 500


def test_compare_run_prints_the_reference_report(tmp_path, capsys):
    """Compare.run through real files (compare.py:39) prints what the reference's own Compare.run printed."""
    from dbb_b200.compare import Compare
    from tests import util
    cases = util.golden_json('ref_compare.json')
    dirs = {}
    for c in cases:
        g, contigs, pr, gcDist, td, covDist = unbinned_case(c['case'])
        if c['case'] not in dirs:
            d = str(tmp_path / c['case'])
            os.makedirs(d)
            dirs[c['case']] = (d, _write_inputs(d, contigs, gcDist, td, covDist))
        d, data = dirs[c['case']]
        capsys.readouterr()
        Compare().run(d, os.path.join(d, 'bins.tsv'), c['bin'], c['scaffold'], c['pers'][0], c['pers'][1], c['pers'][2], dataDir=data)
        assert capsys.readouterr().out == c['stdout'], (c['case'], c['scaffold'])
