cd $GRAFT_REPO_ROOT
cat > /tmp/san_small.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from dbb_b200 import device, synthetic, _pairs
from dbb_b200.data import load_td_dist
from dbb_b200.distributions import PairTables
from dbb_b200.genomicSignatures import GenomicSignatures
meta = synthetic.make_metagenome(seed=3, n_scaffolds=700, min_len=500, max_len=60000, n_genomes=4, frac_n=0.05, frac_lower=0.05, extra_lengths=(300000,))
seqs = synthetic.generate_sequences(meta)
counts, freq = GenomicSignatures(4, 1).signatures(seqs)
tables = PairTables(meta.length, tdDist=load_td_dist(), tdDistPer=80)
a = _pairs.run_knn_host(freq, tables, k=20, prune=1)
b = _pairs.run_knn_host(freq, tables, k=20, prune=0)
assert all(np.array_equal(a[k], b[k]) for k in ('idx', 'dist', 'count', 'neighbour_bp'))
print('sanitizer workload ok', int(counts.sum()))
PY
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_small.py ) > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo memcheck rc=$? >> gpurun_out/r02_sanitizer_memcheck.txt
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_small.py ) > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo racecheck rc=$? >> gpurun_out/r02_sanitizer_racecheck.txt
tail -n 6 gpurun_out/r02_sanitizer_memcheck.txt gpurun_out/r02_sanitizer_racecheck.txt
