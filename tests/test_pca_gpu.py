"""N4 row: PCA of a signature matrix against the reference's own PCA.pcaMatrix (golden) and numpy."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _same_up_to_sign(a, b, tol):
    s = np.sign((a * b).sum(axis=0))
    return np.abs(a - b * s).max() <= tol


def test_pca_matches_reference_golden():
    from dbb_b200.pca import PCA
    g = util.golden_npz('ref_pca.npz')
    A = g['A']
    for tag, frac, centre in (('c', 1.0, True), ('c90', 0.90, True), ('raw', 0.95, False)):
        p = PCA()
        pc, var = p.pcaMatrix(A.copy(), frac, centre, False)
        assert p.npc == int(g[tag + '_npc'][0])
        assert tuple(pc.shape) == tuple(g[tag + '_pc_shape'])
        assert np.abs(var - g[tag + '_variance']).max() <= 1e-12
        k = min(8, pc.shape[1])
        assert np.abs(p.d[:k] - g[tag + '_d'][:k]).max() <= 1e-10 * g[tag + '_d'][0]
        assert _same_up_to_sign(pc[:, :k], g[tag + '_pc'][:, :k], 1e-9 * np.abs(g[tag + '_pc']).max())
        assert _same_up_to_sign(p.Vt[:k].T, g[tag + '_Vt'][:k].T, 1e-9)
        assert np.abs(p.U[:, :k] * p.d[:k] - pc[:, :k]).max() <= 1e-12
    with pytest.raises(ValueError):
        PCA().pcaMatrix(A.copy(), 0.9, True, True)                   # upstream fails the same way (pca.py:61)


def test_pca_file_round_trip_and_numpy_cross_check(tmp_path):
    from dbb_b200.pca import PCA
    rng = np.random.RandomState(3)
    A = rng.dirichlet(np.ones(136), 2500)
    path = tmp_path / 'sig.tsv'
    with open(path, 'w') as f:
        f.write('Sequence Id' + ''.join('\tk%d' % i for i in range(136)) + '\n')
        for i, row in enumerate(A):
            f.write('s%d\t' % i + '\t'.join(repr(float(x)) for x in row) + '\n')
    p = PCA()
    names, pc, var = p.pcaFile(str(path), fraction=1.0, bCenter=True, bScale=False)
    assert list(names[:3]) == ['s0', 's1', 's2'] and pc.shape[0] == 2500
    X = A - A.mean(0)
    X /= X.std(0)
    U, d, Vt = np.linalg.svd(X, full_matrices=False)
    assert np.abs(p.d[:50] - d[:50]).max() <= 1e-9 * d[0]
    assert np.abs(var - d ** 2 / (d ** 2).sum()).max() <= 1e-12
    assert _same_up_to_sign(pc[:, :5], (U * d)[:, :5], 1e-8)
    p.save(str(tmp_path / 'pca.npz'), names, pc, var)
    n2, pc2, var2 = p.loadFile(str(tmp_path / 'pca.npz'))
    assert np.array_equal(pc2, pc) and np.array_equal(var2, var)
