"""(Lives under tests/ because its CPU leg imports the oracle.)  N3 row: time of the GPU-assisted greedy core growth
on a C2-like contig set, and of the reference's loop (oracle restatement of greedy.py:172-251, one core, as upstream)
on a smaller one.  python tests/perf_n3_greedy.py [n_gpu] [n_cpu]"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from dbb_b200 import synthetic
from dbb_b200.data import load_td_dist
from dbb_b200.genomicSignatures import GenomicSignatures
from dbb_b200.greedy import Greedy
from oracle import dbb_oracle as O

n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 1500


class C(object):
    def __init__(self, length, GC, coverage, tetraSig):
        self.binId = -1
        self.length, self.GC, self.coverage, self.tetraSig = length, GC, coverage, tetraSig


def contigs(n, genomes):
    meta = synthetic.make_metagenome(seed=2, n_scaffolds=n, min_len=1000, max_len=100000, n_genomes=genomes)
    rng = np.random.RandomState(0)
    cent = rng.dirichlet(np.ones(136) * 5, genomes)
    sig = cent[meta.genome] * rng.gamma(400, 1 / 400.0, (n, 136))
    sig /= sig.sum(1, keepdims=True)
    order = np.argsort(-meta.length, kind='stable')
    return [C(int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), sig[i].copy()) for i in order]


td = load_td_dist()
gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=3)
PER, MINBIN = 95, 500000
cs = contigs(n_gpu, max(4, n_gpu // 250))
g = Greedy()
g.greedy(cs[:2000], MINBIN, PER, gcDist, td, covDist)                    # warm-up (context, library load)
t0 = time.time()
cores = g.greedy(cs, MINBIN, PER, gcDist, td, covDist)
dt = time.time() - t0
binned = sum(c.binId >= 0 for c in cs)
print('GPU-assisted: %d contigs -> %d cores, %d contigs binned, %d growth steps in %.2f s = %.1f us per step, %.2e candidate '
      'evaluations/s' % (n_gpu, len(cores), binned, g.numSteps, dt, dt / max(g.numSteps, 1) * 1e6, g.numSteps * n_gpu / 2.0 / dt))
cc = contigs(n_cpu, max(4, n_cpu // 250))
t0 = time.time()
co = O.greedy_literal(cc, MINBIN // 10, O.Distributions(gcDist, PER, td, PER, covDist, PER))
dc = time.time() - t0
cg = contigs(n_cpu, max(4, n_cpu // 250))
t0 = time.time()
g.greedy(cg, MINBIN // 10, PER, gcDist, td, covDist)
dg = time.time() - t0
assert [c.binId for c in cc] == [c.binId for c in cg]
print('reference loop (1 core): %d contigs -> %d cores in %.1f s; same instance GPU-assisted %.3f s (%d steps): x%.0f' % (
    n_cpu, len(co), dc, dg, g.numSteps, dc / dg))
