"""N3 row: greedy core growth against the reference's own outputs (golden) and the oracle's literal loop."""
import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util
from tests.test_oracle_golden import (_GContig, check_greedy_against_golden, check_merge_against_golden, greedy_case,
                                      greedy_merge_case)

pytestmark = pytest.mark.gpu


def test_greedy_matches_reference_golden():
    from dbb_b200.greedy import Greedy
    for tag in ('a', 'b'):
        g, contigs, per, minBin, gcDist, tdDist, covDist = greedy_case(tag)
        gr = Greedy()
        cores = gr.greedy(contigs, minBin, per, gcDist, tdDist, covDist)
        check_greedy_against_golden(g, tag, contigs, cores)
        assert gr.numSteps >= len(contigs) - (np.asarray(g[tag + '_binId']) < 0).sum()
        assert all(c.bProcessed == (c.binId >= 0) for c in contigs)


def test_greedy_random_instance_vs_literal_loop():
    """600 contigs from 8 genomes, percentile 90, a bin-size threshold that releases some cores; exact ties in coverage
    (the later contig must win) are planted."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.genomicSignatures import GenomicSignatures
    from dbb_b200.greedy import Greedy
    meta = synthetic.make_metagenome(seed=41, n_scaffolds=600, min_len=1500, max_len=20000, n_genomes=8)
    seqs = synthetic.generate_sequences(meta)
    counts, _ = GenomicSignatures(4, 1).signatures(seqs, want_freq=False)
    sig = counts.astype(np.float64) / counts.sum(axis=1, keepdims=True)
    order = np.argsort(-meta.length, kind='stable')
    cov = meta.coverage.astype(np.float64).copy()
    same = np.nonzero(meta.genome == meta.genome[order[5]])[0]
    cov[same[:6]] = cov[same[0]]                                        # equal |cov - core.cov| among candidates
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    mk = lambda: [_GContig(int(meta.length[i]), float(meta.gc_obs[i]), float(cov[i]), sig[i].copy()) for i in order]
    a, b = mk(), mk()
    cores_o = O.greedy_literal(a, 60000, O.Distributions(gcDist, 90, td, 90, covDist, 90))
    cores_g = Greedy().greedy(b, 60000, 90, gcDist, td, covDist)
    assert [c.binId for c in a] == [c.binId for c in b]
    assert len(cores_o) == len(cores_g) and len(cores_o) >= 3
    assert any(c.binId < 0 for c in a)
    for co, cg in zip(cores_o, cores_g):
        assert co.length == cg.length and co.GC == cg.GC and co.coverage == cg.coverage
        assert np.array_equal(np.asarray(co.tetraSig), np.asarray(cg.tetraSig))
    assert Greedy().greedy([], 1000, 90, gcDist, td, covDist) == []


def test_merge_matches_reference_golden_and_literal_loop():
    """Greedy.merge (greedy.py:253-334) against the reference's own __merge outputs, then grow + merge on a random
    instance against the oracle's literal loops."""
    from dbb_b200.greedy import Greedy
    for tag in ('a', 'b'):
        g, cores, per, gcDist, tdDist, covDist = greedy_merge_case(tag)
        merged = Greedy().merge(cores, per, gcDist, tdDist, covDist)
        check_merge_against_golden(g, tag, cores, merged)
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.genomicSignatures import GenomicSignatures
    meta = synthetic.make_metagenome(seed=43, n_scaffolds=500, min_len=1500, max_len=20000, n_genomes=6)
    counts, _ = GenomicSignatures(4, 1).signatures(synthetic.generate_sequences(meta), want_freq=False)
    sig = counts.astype(np.float64) / counts.sum(axis=1, keepdims=True)
    order = np.argsort(-meta.length, kind='stable')
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    mk = lambda: [_GContig(int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), sig[i].copy()) for i in order]
    a, b = mk(), mk()
    cores_o = O.greedy_literal(a, 5000, O.Distributions(gcDist, 50, td, 50, covDist, 50))
    cores_g = Greedy().greedy(b, 5000, 50, gcDist, td, covDist)
    merged_o = O.merge_literal(cores_o, O.Distributions(gcDist, 99, td, 99, covDist, 99))
    merged_g = Greedy().merge(cores_g, 99, gcDist, td, covDist)
    assert len(merged_o) == len(merged_g) < len(cores_g)
    assert [c.binId for c in a] == [c.binId for c in b]
    for mo, mg in zip(merged_o, merged_g):
        assert mo.length == mg.length and mo.GC == mg.GC and mo.coverage == mg.coverage
        assert np.array_equal(np.asarray(mo.tetraSig), np.asarray(mg.tetraSig)) and len(mo.contigs) == len(mg.contigs)
    assert Greedy().merge([], 99, gcDist, td, covDist) == []


def test_sort_contigs_neighbourhood_sizes_match_reference():
    """Greedy.sortContigs (greedy.py:390-422): neighbourhood sizes and order equal to what the reference's own lines
    406-420 computed from its own worker-loop rows; attributes set as the three run() calls set them; upstream's
    dropped sort (F8) is opt-in."""
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    from dbb_b200.greedy import Greedy
    g = util.golden_npz('ref_greedy_merge.npz')
    n = len(g['s_length'])
    fixed, per = int(g['s_params'][0]), int(g['s_params'][1])
    td = load_td_dist()
    gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=7)
    contigs = [_GContig(int(g['s_length'][i]), float(g['s_gc'][i]), float(g['s_cov'][i]), g['s_sig'][i]) for i in range(n)]
    same, sizes, order = Greedy().sortContigs(contigs, fixed, per, gcDist, td, covDist)
    assert same is contigs
    assert np.array_equal(sizes, g['s_sizes']) and np.array_equal(order, g['s_sorted'])
    o_sizes, o_order = O.neighbourhood_sizes_literal(contigs)            # the reference's expressions on OUR attributes
    assert np.array_equal(np.array(o_sizes), g['s_sizes'])
    srt, _, _ = Greedy().sortContigs(contigs, fixed, per, gcDist, td, covDist, bSort=True)
    assert [id(c) for c in srt] == [id(contigs[i]) for i in g['s_sorted']]
