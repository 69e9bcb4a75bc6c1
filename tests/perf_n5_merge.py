"""(Lives under tests/ because its CPU leg imports the oracle.)  N5 row: time of the GPU merge table (merge.py:96-156)
on a C2-sized binning, and of the reference's loop (oracle restatement, one core, as upstream) on a smaller one.
python tests/perf_n5_merge.py [n_gpu] [bins_gpu] [n_cpu] [bins_cpu]"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from dbb_b200 import synthetic
from dbb_b200.data import load_td_dist
from dbb_b200.merge import Merge
from oracle import dbb_oracle as O

n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
b_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 600
b_cpu = int(sys.argv[4]) if len(sys.argv) > 4 else 12


def contigs(n, nbins):
    genomes = max(2, nbins // 2)                                     # two bins per genome: merge candidates exist
    meta = synthetic.make_metagenome(seed=2, n_scaffolds=n, min_len=1000, max_len=100000, n_genomes=genomes)
    rng = np.random.RandomState(0)
    cent = rng.dirichlet(np.ones(136) * 5, genomes)
    sig = cent[meta.genome] * rng.gamma(400, 1 / 400.0, (n, 136))
    sig /= sig.sum(1, keepdims=True)
    out = []
    for i in range(n):
        c = O.Contig('c%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), sig[i].copy())
        c.binId = 'bin%d' % (int(meta.genome[i]) * 2 + int(rng.randint(0, 2)))
        out.append(c)
    return out


td = load_td_dist()
gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=3)
PERS = (99, 99, 99)
m = Merge()
args = (gcDist, PERS[0], td, PERS[1], covDist, PERS[2])
m.mergeTable(contigs(500, 6), *args)                                 # warm-up (context, library load)
cs = contigs(n_gpu, b_gpu)
t0 = time.time()
bins, table = m.mergeTable(cs, *args)
dt = time.time() - t0
B = len(bins)
print('GPU merge table: %d binned contigs x %d bins in %.3f s (host arrays in, table out; %.2e contig-bin pairs/s); the '
      "reference's loop visits contigs x (bins-1) x bins = %.2e inner iterations" % (n_gpu, B, dt, n_gpu * B / dt, n_gpu * (B - 1.0) * B))
cc = contigs(n_cpu, b_cpu)
t0 = time.time()
txt = O.merge_table_literal(cc, O.Distributions(*args))
dc = time.time() - t0
t0 = time.time()
bins2, table2 = m.mergeTable(cc, *args)
dg = time.time() - t0
assert m.formatTable(bins2, table2) == txt
inner = n_cpu * (len(bins2) - 1.0) * len(bins2)
print('reference loop (1 core): %d contigs x %d bins in %.1f s = %.1f us per inner iteration; same instance on the GPU %.4f s '
      '(x%.0f), files identical; extrapolated to the large instance: %.0f s' % (n_cpu, len(bins2), dc, dc / inner * 1e6, dg, dc / dg,
                                                                               dc / inner * n_gpu * (B - 1.0) * B))
