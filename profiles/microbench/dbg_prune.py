import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from dbb_b200 import device, synthetic
from dbb_b200.distributions import PairTables
from dbb_b200.genomicSignatures import GenomicSignatures
from dbb_b200.data import load_td_dist
meta = synthetic.make_metagenome(seed=51, n_scaffolds=2600, min_len=1000, max_len=60000, n_genomes=12)
seqs = synthetic.generate_sequences(meta)
_, freq = GenomicSignatures(4, 1).signatures(seqs)
freq[100] = freq[101]; freq[2000] = freq[101]
sig = device.pad_signatures(torch.from_numpy(freq).cuda())
td = load_td_dist(); gcd = synthetic.synthetic_gc_dist(sorted(td.keys())); cvd = synthetic.synthetic_cov_dist(seed=7)
n = meta.n_scaffolds
tables = PairTables(meta.length, GC=meta.gc_obs, coverage=meta.coverage, tdDist=td, tdDistPer=80, gcDist=gcd, gcDistPer=80, covDist=cvd, covDistPer=80)
dt = device.DevicePairTables(tables)
filt = device.FILTER_GC | device.FILTER_COV
full = device.pairs(sig, dt, k=20, topk_filter=filt, want_count=True)
pr = device.pairs_pruned(sig, dt, k=20, topk_filter=filt, want_count=True)
got = device.unsort_rows(pr, n)
print('tiles', pr['tiles_visited'], pr['tiles_total'])
for key in ('idx','dist','count','neighbour_bp'):
    neq = (got[key] != full[key])
    if neq.dim() > 1: neq = neq.any(1)
    print(key, int(neq.sum()))
bad = (got['idx'] != full['idx']).any(1).nonzero()[:,0]
for r in bad[:3].tolist():
    print('row', r, 'len', meta.length[r])
    print(' full idx', full['idx'][r].tolist()); print(' got  idx', got['idx'][r].tolist())
    print(' full d', full['dist'][r].tolist()); print(' got  d', got['dist'][r].tolist())
