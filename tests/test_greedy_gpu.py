"""N3 row: greedy core growth against the reference's own outputs (golden) and the oracle's literal loop."""
import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util
from tests.test_oracle_golden import _GContig, check_greedy_against_golden, greedy_case

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
