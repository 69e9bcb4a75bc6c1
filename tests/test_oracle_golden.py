"""CPU: the oracle against the golden vectors that tests/golden/make_golden.py produced by executing the
reference's own source files (this is what pins the oracle), plus literal-vs-vectorised self checks."""
import hashlib

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import dbb_oracle as O
from tests import util


def test_column_table_matches_reference_and_survey_hashes():
    t = util.golden_json('ref_table.json')
    assert t['kmerCols'] == O.KMER_COLS and t['numKmers'] == 136 == O.NUM_KMERS
    assert t['kmerToCanonicalIndex'] == O.KMER_TO_INDEX
    assert O.KMER_COLS == sorted(O.KMER_COLS)
    assert sum(1 for k in O.KMER_COLS if O.rev_comp(k) == k) == 16
    assert hashlib.sha256('\t'.join(O.KMER_COLS).encode()).hexdigest() == \
        'e50eaf21f505fba2ce9262fc8b7e053c1e29a54a276be01e6227f2d1c3503422'
    assert hashlib.sha256(bytes(O.code_lut256())).hexdigest() == \
        '9a57f94a839828cfdd42650170aa4db6e761932f7724c8056ab4ea818d0bc72f'


def test_signatures_match_reference_bit_for_bit():
    kat = util.golden_json('ref_kat_seqs.json')
    ref = util.golden_npz('ref_signatures.npz')['sigs']
    assert len(kat) == ref.shape[0] >= 50
    for s, g in zip(kat, ref):
        assert np.array_equal(O.seq_signature_literal(s), g, equal_nan=True)
        assert np.array_equal(O.seq_signature_np(s), g, equal_nan=True)


def test_known_answers_by_hand():
    c = O.seq_counts_literal('ACGT'); assert sum(c) == 1 and c[O.KMER_COLS.index('ACGT')] == 1
    assert O.seq_counts_literal('AAAA') == O.seq_counts_literal('TTTT')
    assert O.seq_counts_literal('AAAAA')[O.KMER_COLS.index('AAAA')] == 2
    assert sum(O.seq_counts_literal('ACGNACGT')) == 1 and sum(O.seq_counts_literal('acgtACGT')) == 1
    for s in ('', 'ACG', 'NNNN'):
        assert sum(O.seq_counts_literal(s)) == 0 and np.isnan(O.seq_signature_literal(s)).all()
    s = 'GATTACAGATTACACCGGTTAACC'
    assert O.seq_counts_literal(s) == O.seq_counts_literal(O.rev_comp(s))


@settings(max_examples=150, deadline=None)
@given(st.text(alphabet='ACGTNacgtRYU', min_size=0, max_size=300))
def test_literal_equals_vectorised(s):
    assert O.seq_counts_literal(s) == O.seq_counts_np(s).tolist()
    valid = sum(1 for i in range(len(s) - 3) if all(ch in 'ACGT' for ch in s[i:i + 4]))
    assert sum(O.seq_counts_literal(s)) == valid


def _golden_contigs():
    nb = util.golden_npz('ref_neighbours.npz')
    cs = [O.Contig('c%d' % i, int(nb['length'][i]), float(nb['GC'][i]), float(nb['coverage'][i]), nb['tetraSig'][i])
          for i in range(len(nb['length']))]
    return nb, cs


def test_neighbour_rows_match_reference_loops():
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    nb, cs = _golden_contigs()
    td, gcd, cvd = load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=7)
    n = len(cs)
    assert np.array_equal(np.array([[O.distance(a.tetraSig, b.tetraSig) for b in cs] for a in cs]), nb['distance'], equal_nan=True)
    for tag, fixed, per in (('fixedNone_p80', None, 80), ('fixed10000_p80', 10000, 80), ('fixedNone_p99', None, 99)):
        dT = O.Distributions(None, None, td, per, None, None)
        dG = O.Distributions(gcd, per, None, None, None, None)
        dC = O.Distributions(None, None, None, None, cvd, per)
        tb = O.DistTables(gcd, per, td, per, cvd, per)
        rows = list(range(n))
        with np.errstate(invalid='ignore'):
            assert np.array_equal(np.array([O.td_row_literal(c, cs, dT, fixed) for c in cs]), nb['td_' + tag])
            assert np.array_equal(O.td_rows_np(nb['tetraSig'], nb['length'], rows, tb, fixed), nb['td_' + tag])
        assert np.array_equal(np.array([O.gc_row_literal(c, cs, dG, fixed) for c in cs]), nb['gc_' + tag])
        assert np.array_equal(np.array([O.cov_row_literal(c, cs, dC, fixed) for c in cs]), nb['cov_' + tag])
        assert np.array_equal(O.gc_rows_np(nb['length'], nb['GC'], rows, tb, fixed), nb['gc_' + tag])
        assert np.array_equal(O.cov_rows_np(nb['length'], nb['coverage'], rows, tb, fixed), nb['cov_' + tag])
    # semantics the golden set was built to exercise
    td80 = nb['td_fixedNone_p80']
    assert (np.diag(td80)[np.arange(n) != 8] == 1).all()           # i == j is a neighbour ...
    assert td80[8].sum() == 0 and td80[:, 8].sum() == 0             # ... except for a NaN signature
    assert nb['length'][5] == nb['length'][6]                       # equal lengths: core = j in both orders


def test_td_dist_values_from_survey():
    from dbb_b200.data import load_td_dist
    td = load_td_dist()
    assert len(td) == 41 and len(td[500]) == 201
    assert td[10000][80.0] == 0.1542 and td[10000][99.0] == 0.351339 and td[500][95.0] == 0.549141


def test_tsv_format_is_py2_str():
    assert O.format_sig_value(0.0073529411764705881) == '0.00735294117647'
    assert O.format_sig_value(1.0) == '1.0' and O.format_sig_value(0.0) == '0.0'
    assert O.format_sig_value(float('nan')) == 'nan' and O.format_sig_value(1e-05) == '1e-05'
    assert O.tsv_header().startswith('Sequence Id\tAAAA\tAAAC') and O.tsv_header().count('\t') == 136


def _refine_setup():
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    nb = util.golden_npz('ref_neighbours.npz')
    sig = nb['tetraSig'].copy()
    sig[8] = nb['refine_sig8']
    cs = [O.Contig('c%d' % i, int(nb['length'][i]), float(nb['GC'][i]), float(nb['coverage'][i]), sig[i]) for i in range(sig.shape[0])]
    for c, b in zip(cs, nb['refine_bin_of_contig']):
        c.binId = int(b)
    return nb, cs, (load_td_dist(), synthetic.synthetic_gc_dist(), synthetic.synthetic_cov_dist(seed=7))


def test_refine_worker_matches_reference():
    """refineBins.py:42-96 and :288-304 against the rows the reference's own __workerThread produced."""
    nb, cs, (td, gcd, cvd) = _refine_setup()
    cores = O.core_statistics_literal(cs[:30])
    assert sorted(cores) == nb['refine_core_ids'].tolist()
    for n, k in enumerate(sorted(cores)):
        assert cores[k].length == nb['refine_core_len'][n] and cores[k].GC == nb['refine_core_gc'][n]
        assert cores[k].coverage == nb['refine_core_cov'][n] and np.array_equal(np.array(cores[k].tetraSig), nb['refine_core_sig'][n])
    dist = O.Distributions(gcd, 99, td, 99, cvd, 99)
    for row, c in zip(nb['refine_rows'], cs[30:]):
        _, _, newBin, bA, bF, bI = O.refine_worker_literal(c, cores, dist)
        assert [newBin, c.length, int(bA), int(bF), int(bI)] == row[1:].tolist()


def test_c_restatement_equals_literal_oracle():
    """oracle/dbb_oracle.c (used for full-size checks) against the Python literal forms."""
    from oracle import c_oracle as C
    t, names = C.kmer_table()
    assert names == O.KMER_COLS and np.array_equal(t, O.code_lut256())
    kat = util.golden_json('ref_kat_seqs.json')
    arrs = [np.frombuffer(s.encode(), dtype=np.uint8) for s in kat]
    off = np.zeros(len(arrs) + 1, np.int64); off[1:] = np.cumsum([a.size for a in arrs])
    c = C.counts_batch(np.concatenate(arrs), off)
    for i, s in enumerate(kat):
        assert c[i].tolist() == O.seq_counts_literal(s)
    nb, cs = _golden_contigs()
    from dbb_b200.data import load_td_dist
    tb = O.DistTables(tdDist=load_td_dist(), tdDistPer=80)
    sig = nb['tetraSig']
    bound = tb.tdBound[O.nearest_index_sorted(tb.tdLens, nb['length'])]
    with np.errstate(invalid='ignore'):
        dist, mask = C.td_rows(sig, nb['length'], bound, np.arange(len(cs)))
    assert np.array_equal(mask.astype(np.float64), nb['td_fixedNone_p80'])
    assert np.allclose(dist, nb["distance"], rtol=0, atol=1e-13, equal_nan=True)   # summation order differs from numpy


def test_split_and_base_count_oracle_matches_reference_outputs():
    """N2 row: the literal restatements of splitScaffolds.py:27-53 and seqUtils.py:258-265 reproduce what the reference's
    own code returned (tests/golden/make_golden_split.py)."""
    g = util.golden_json('ref_split.json')
    seqs = g['seqs']
    for case in g['splits']:
        i = case['seq']
        got = O.splits_literal('s%d' % i, seqs[i], case['minN'])
        assert got == case['contigs'], (i, case['minN'])
    for s, want in zip(seqs, g['baseCount']):
        assert list(O.base_count_literal(s)) == want
    assert any(v == [0, 0] for c in g['splits'] for v in c['contigs'].values())      # the empty leading contig quirk


class _GContig(object):
    def __init__(self, length, GC, coverage, tetraSig):
        self.binId = -1
        self.length, self.GC, self.coverage, self.tetraSig = length, GC, coverage, tetraSig


def greedy_case(tag):
    """Inputs and the reference's own outputs of one golden greedy case (tests/golden/make_golden_greedy.py)."""
    import ast
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    g = util.golden_npz('ref_greedy.npz')
    contigs = [_GContig(int(l), float(a), float(b), s.copy()) for l, a, b, s in
               zip(g[tag + '_length'], g[tag + '_gc'], g[tag + '_cov'], g[tag + '_sig'])]
    td = load_td_dist()
    return g, contigs, int(g[tag + '_params'][0]), int(g[tag + '_params'][1]), synthetic.synthetic_gc_dist(sorted(td.keys())), td, \
        synthetic.synthetic_cov_dist(seed=7)


def check_greedy_against_golden(g, tag, contigs, cores):
    assert np.array_equal(np.array([c.binId for c in contigs]), g[tag + '_binId'])
    assert len(cores) == len(g[tag + '_core_len'])
    for b, core in enumerate(cores):
        assert core.length == g[tag + '_core_len'][b] and core.GC == g[tag + '_core_gc'][b] and core.coverage == g[tag + '_core_cov'][b]
        assert np.array_equal(np.asarray(core.tetraSig), g[tag + '_core_sig'][b])        # same float64 operations, same order
        assert len(core.contigs) == g[tag + '_core_sizes'][b]


def test_greedy_oracle_matches_reference_outputs():
    """N3 row: the literal restatement of greedy.py:172-251 reproduces what the reference's own __greedy produced."""
    for tag in ('a', 'b'):
        g, contigs, per, minBin, gcDist, tdDist, covDist = greedy_case(tag)
        cores = O.greedy_literal(contigs, minBin, O.Distributions(gcDist, per, tdDist, per, covDist, per))
        check_greedy_against_golden(g, tag, contigs, cores)


def greedy_merge_case(tag):
    """Input cores (with dummy contigs that remember their core) and the reference's own outputs of one golden __merge
    case (tests/golden/make_golden_greedy_merge.py)."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    g = util.golden_npz('ref_greedy_merge.npz')
    cores = []
    for b in range(len(g[tag + '_in_len'])):
        core = O.Core(b, int(g[tag + '_in_len'][b]), float(g[tag + '_in_gc'][b]), float(g[tag + '_in_cov'][b]), g[tag + '_in_sig'][b].copy())
        core.contigs = [_GContig(1, 0.5, 1.0, None) for _ in range(int(g[tag + '_in_sizes'][b]))]
        for c in core.contigs:
            c.binId = b
        cores.append(core)
    td = load_td_dist()
    return g, cores, int(g[tag + '_params'][1]), synthetic.synthetic_gc_dist(sorted(td.keys())), td, synthetic.synthetic_cov_dist(seed=7)


def check_merge_against_golden(g, tag, cores, merged):
    assert np.array_equal(np.array([c.binId for c in cores]), g[tag + '_core_bin'])
    assert len(merged) == len(g[tag + '_out_len']) < len(cores)
    for b, m in enumerate(merged):
        assert m.binId == b
        assert m.length == g[tag + '_out_len'][b] and m.GC == g[tag + '_out_gc'][b] and m.coverage == g[tag + '_out_cov'][b]
        assert np.array_equal(np.asarray(m.tetraSig), g[tag + '_out_sig'][b])             # same float64 operations, same order
        assert len(m.contigs) == g[tag + '_out_sizes'][b] and all(c.binId == b for c in m.contigs)
    # (a seed core keeps the binId __greedy gave it -- upstream only renumbers the cores it absorbs, greedy.py:311 -- so
    # _core_bin holds old ids for the seeds; the contigs carry the merged ids, checked above)
    assert int(g[tag + '_out_sizes'].sum()) == int((g[tag + '_contig_bin'] >= 0).sum())


def test_greedy_merge_and_sort_oracle_match_reference_outputs():
    """Rest of the N3 row: the literal restatements of greedy.py:258-333 (__merge) and :406-420 (neighbourhood sizes of
    __sortContigs) reproduce what the reference's own source produced."""
    for tag in ('a', 'b'):
        g, cores, per, gcDist, tdDist, covDist = greedy_merge_case(tag)
        merged = O.merge_literal(cores, O.Distributions(gcDist, per, tdDist, per, covDist, per))
        check_merge_against_golden(g, tag, cores, merged)
    g = util.golden_npz('ref_greedy_merge.npz')
    n = len(g['s_length'])
    fixed, per = int(g['s_params'][0]), int(g['s_params'][1])
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    td = load_td_dist()
    ot = O.DistTables(synthetic.synthetic_gc_dist(sorted(td.keys())), per, td, per, synthetic.synthetic_cov_dist(seed=7), per)
    rows = list(range(n))
    contigs = [_GContig(int(g['s_length'][i]), float(g['s_gc'][i]), float(g['s_cov'][i]), g['s_sig'][i]) for i in rows]
    for i, c in enumerate(contigs):
        c.tdNeighbours = O.td_rows_np(g['s_sig'], g['s_length'], [i], ot, fixed)[0]
        c.gcNeighbours = O.gc_rows_np(g['s_length'], g['s_gc'], [i], ot, fixed)[0]
        c.covNeighbours = O.cov_rows_np(g['s_length'], g['s_cov'], [i], ot, fixed)[0]
    sizes, order = O.neighbourhood_sizes_literal(contigs)
    assert np.array_equal(np.array(sizes), g['s_sizes']) and np.array_equal(order, g['s_sorted'])


def merge_case(tag):
    """(golden, contigs in file order with binId set, (gcPer, tdPer, covPer), gcDist, tdDist, covDist) of the N5 golden
    vectors (tests/golden/make_golden_merge.py)."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    g = util.golden_npz('ref_merge.npz')
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    contigs = []
    for i in range(len(g[tag + '_length'])):
        c = O.Contig('c%d' % i, int(g[tag + '_length'][i]), float(g[tag + '_gc'][i]), float(g[tag + '_cov'][i]), g[tag + '_sig'][i])
        b = str(g[tag + '_bin'][i])
        c.binId = -1 if b == 'unbinned' else b
        contigs.append(c)
    return g, contigs, [int(x) for x in g[tag + '_pers']], gcDist, td, covDist


def test_merge_oracle_matches_reference_outputs():
    """N5 row: the literal restatement of merge.py:66-156 writes, byte for byte, the file the reference's own Merge.run
    wrote (three cases; in case c every genome is cut into three bins, so the three 'closest' columns differ)."""
    for tag in 'abc':
        g, contigs, pers, gcDist, td, covDist = merge_case(tag)
        dist = O.Distributions(gcDist, pers[0], td, pers[1], covDist, pers[2])
        assert O.merge_table_literal(contigs, dist) == str(g[tag + '_merge_tsv']), tag
    rows = [r.split('\t') for r in str(g['c_merge_tsv']).strip().split('\n')[1:]]
    assert any(r[10] != r[11] or r[11] != r[12] for r in rows)


def unbinned_case(tag):
    """(golden, contigs in file order with binId set, [gcPer, tdPer, covPer, minScaffoldLen, minContigLen, allScaffolds],
    gcDist, tdDist, covDist) of the N6 golden vectors (tests/golden/make_golden_unbinned.py)."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    g = util.golden_npz('ref_unbinned.npz')
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    contigs = []
    for i in range(len(g[tag + '_length'])):
        c = O.Contig(str(g[tag + '_id'][i]), int(g[tag + '_length'][i]), float(g[tag + '_gc'][i]), float(g[tag + '_cov'][i]),
                     g[tag + '_sig'][i])
        b = str(g[tag + '_bin'][i])
        c.binId = -1 if b == 'unbinned' else b
        contigs.append(c)
    return g, contigs, [int(x) for x in g[tag + '_params']], gcDist, td, covDist


def test_unbinned_oracle_matches_reference_outputs():
    """N6 row: the literal restatement of unbinned.py:70-186 writes, byte for byte, the file the reference's own
    Unbinned.run wrote (case a: length thresholds, unbinned scaffolds only; case b: all scaffolds)."""
    for tag in 'ab':
        g, contigs, pr, gcDist, td, covDist = unbinned_case(tag)
        dist = O.Distributions(gcDist, pr[0], td, pr[1], covDist, pr[2])
        assert O.unbinned_table_literal(contigs, dist, pr[3], pr[4], bool(pr[5])) == str(g[tag + '_unbinned_tsv']), tag


def test_compare_oracle_matches_reference_outputs():
    """compare.py: the literal restatement of :70-152 prints what the reference's own Compare.run printed (twelve
    scaffold / bin queries over the two N6 cases, tests/golden/make_golden_compare.py)."""
    cases = util.golden_json('ref_compare.json')
    assert len(cases) == 12
    for c in cases:
        g, contigs, pr, gcDist, td, covDist = unbinned_case(c['case'])
        dist = O.Distributions(gcDist, c['pers'][0], td, c['pers'][1], covDist, c['pers'][2])
        assert O.compare_text_literal(contigs, c['bin'], c['scaffold'], dist) == c['stdout'], (c['case'], c['scaffold'])
