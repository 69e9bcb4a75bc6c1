"""(Lives under tests/ because its CPU leg imports the oracle.)  N6 row: time of the GPU table of unbinned.py:111-186 on a
C2-sized contig set, and of the reference's loop (oracle restatement, one core, as upstream) on a smaller one.
python tests/perf_n6_unbinned.py [n_gpu] [bins_gpu] [n_cpu] [bins_cpu]"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from dbb_b200 import synthetic
from dbb_b200.data import load_td_dist
from dbb_b200.unbinned import Unbinned, Scaffold
from oracle import dbb_oracle as O

n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
b_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 600
b_cpu = int(sys.argv[4]) if len(sys.argv) > 4 else 12


def contigs(n, nbins):
    genomes = max(2, nbins // 2)
    meta = synthetic.make_metagenome(seed=2, n_scaffolds=n, min_len=1000, max_len=100000, n_genomes=genomes)
    rng = np.random.RandomState(0)
    cent = rng.dirichlet(np.ones(136) * 5, genomes)
    sig = cent[meta.genome] * rng.gamma(400, 1 / 400.0, (n, 136))
    sig /= sig.sum(1, keepdims=True)
    out = []
    for i in range(n):
        c = O.Contig('s%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), sig[i].copy())
        c.binId = -1 if rng.rand() < 0.5 else 'bin%d' % (int(meta.genome[i]) * 2 + int(rng.randint(0, 2)))
        out.append(c)
    return out


def split(cs):
    scaffolds = []
    for c in cs:
        if c.binId == -1:
            s = Scaffold(c.contigId)
            s.add(c)
            scaffolds.append(s)
    return scaffolds, [c for c in cs if c.binId != -1]


td = load_td_dist()
gcDist, covDist = synthetic.synthetic_gc_dist(sorted(td.keys())), synthetic.synthetic_cov_dist(seed=3)
args = (gcDist, 99, td, 99, covDist, 99)
u = Unbinned()
u.unbinnedTable(*split(contigs(500, 6)), *args)                       # warm-up (context, library load)
sc, binned = split(contigs(n_gpu, b_gpu))
t0 = time.time()
bins, table = u.unbinnedTable(sc, binned, *args)
dt = time.time() - t0
B = len(bins)
print("GPU table: %d unbinned contigs x %d bins (%d binned contigs define them) in %.3f s (host arrays in, table out); the "
      "reference's loop visits up to contigs x bins x (1 + bins) = %.2e inner iterations" % (len(sc), B, len(binned), dt, len(sc) * B * (1.0 + B)))
cc = contigs(n_cpu, b_cpu)
t0 = time.time()
txt = O.unbinned_table_literal(cc, O.Distributions(*args), 0, 0, False)
dc = time.time() - t0
sc2, binned2 = split(cc)
t0 = time.time()
bins2, table2 = u.unbinnedTable(sc2, binned2, *args)
dg = time.time() - t0
assert u.formatTable(sc2, bins2, table2, 0) == txt
print('reference loop (1 core): %d unbinned contigs x %d bins in %.2f s; same instance on the GPU %.4f s (x%.0f), files identical'
      % (len(sc2), len(bins2), dc, dg, dc / dg))
