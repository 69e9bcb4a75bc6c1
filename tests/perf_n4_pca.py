"""N4 row: wall time of PCA().pcaMatrix on a C2-sized signature matrix (50 000 x 136 float64, host arrays in and out)
against the reference's own computation (numpy: centre, scale, full SVD -- pca.py:56-75) on the box's host cores."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from dbb_b200.pca import PCA

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.RandomState(0)
cent = rng.dirichlet(np.ones(136) * 5, 200)
A = cent[rng.randint(0, 200, n)] * rng.gamma(50, 1 / 50.0, (n, 136))
A /= A.sum(1, keepdims=True)
p = PCA()
p.pcaMatrix(A[:2000].copy(), 1.0, True, False)              # warm-up
t0 = time.time()
pc, var = p.pcaMatrix(A, 1.0, True, False)
dt = time.time() - t0
t0 = time.time()
X = A - A.mean(0)
X /= X.std(0)
U, d, Vt = np.linalg.svd(X, full_matrices=False)
ref = U[:, :p.npc] * d[:p.npc]
dc = time.time() - t0
print('PCA of %d x 136: GPU-backed pcaMatrix %.3f s (Gram + eigh + scores, incl. 2 x %.0f MB H2D and %.0f MB D2H), numpy SVD path %.3f s; '
      'max |variance diff| %.1e' % (n, dt, A.nbytes / 1e6, pc.nbytes / 1e6, dc, np.abs(var - d ** 2 / (d ** 2).sum()).max()))
