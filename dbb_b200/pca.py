"""PCA of the signature matrix -- drop-in for dbb/pca.py (``PCA``, ``Center``), GPU backed.

``PCA().pcaMatrix(A, fraction, bCenter, bScale)`` returns ``(pc, variance)`` and leaves ``d, Vt, variance, sumvariance,
npc, dinv`` (and ``U``) on the object as pca.py:56-75 does; ``pcaFile``, ``pc``, ``save``, ``loadFile`` keep their
signatures.  The Gram matrix of the centred data and the scores are computed on the GPU in FP64 (csrc/pca.cu), the
136 x 136 symmetric eigenproblem on the host.  Differences from ``numpy.linalg.svd`` that any caller of an SVD must
already tolerate: the sign of each component is arbitrary (here: the largest-magnitude entry of each ``Vt`` row is
positive) and directions with singular values at rounding level (signatures sum to 1, so the centred matrix has rank
<= 135) are numerical noise.  Unlike ``Center`` the caller's array is not modified in place.  Upstream's argument
mix-up in pca.py:61 (``bScale`` passed as ``axis``) is reproduced, see ``pcaMatrix``.
"""
import ctypes

import numpy as np

from . import _lib


class PCA(object):
    def __init__(self):
        pass

    def pcaFile(self, dataMatrixFile, fraction=0.90, bCenter=True, bScale=True):
        """pca.py:30-54: TSV with row and column headers; rows are data points, columns are variables."""
        names, rows = [], []
        with open(dataMatrixFile) as f:
            next(f)
            for line in f:
                lineSplit = line.split('\t')
                names.append(lineSplit[0])
                rows.append([float(x) for x in lineSplit[1:]])
        data = np.array(rows, dtype=np.float64)
        self.pcaMatrix(data, fraction, bCenter, bScale)
        return np.array(names), self.pc(), self.variance

    def pcaMatrix(self, A, fraction=0.90, bCenter=True, bScale=True):
        assert 0 <= fraction <= 1
        A = np.ascontiguousarray(A, dtype=np.float64)
        n, w = A.shape
        L = _lib.lib()
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        mean, std, gram = np.zeros(w), np.ones(w), np.zeros((w, w))
        if bCenter:
            # pca.py:60-61 calls Center(A, bScale): bScale lands in Center's `axis` parameter and `scale` keeps its
            # default True.  So bScale=False (the only upstream call, main.py:327) centres the columns AND divides them
            # by their standard deviation; bScale=True asks for axis 1 and fails on the subtraction (kept: ValueError).
            if bScale:
                raise ValueError('operands could not be broadcast together (pca.py:61 passes bScale as Center\'s axis)')
            _lib.check(L.dbb_pca_gram_host(vp(A), n, w, 1, vp(mean), vp(std), vp(gram)))
        else:
            gram = self._gram_uncentred(A, L, vp)
        self.mean, self.std = mean, std
        evals, evecs = np.linalg.eigh((gram + gram.T) * 0.5)
        order = np.argsort(-evals, kind='stable')
        evals, evecs = np.maximum(evals[order], 0.0), evecs[:, order]
        k = min(n, w)
        self.d = np.sqrt(evals[:k])
        Vt = evecs[:, :k].T.copy()
        sign = np.sign(Vt[np.arange(k), np.abs(Vt).argmax(axis=1)])
        sign[sign == 0] = 1.0
        self.Vt = Vt * sign[:, None]
        assert np.all(self.d[:-1] >= self.d[1:])
        self.variance = self.d ** 2
        self.variance /= np.sum(self.variance)
        self.sumvariance = np.cumsum(self.variance)
        self.npc = int(np.searchsorted(self.sumvariance, fraction) + 1)
        self.dinv = np.array([1 / d if d > self.d[0] * 1e-6 else 0 for d in self.d])
        npc = min(self.npc, k)
        V = np.ascontiguousarray(self.Vt[:npc].T)
        self._scores = np.zeros((n, npc))
        _lib.check(L.dbb_pca_project_host(vp(A), n, w, vp(mean), vp(std), vp(V), npc, vp(self._scores)))
        return self.pc(), self.variance

    @staticmethod
    def _gram_uncentred(A, L, vp):
        # the uncentred case: Gram matrix of A itself = centred Gram + n * mean mean^T
        n, w = A.shape
        mean, std, gram = np.zeros(w), np.ones(w), np.zeros((w, w))
        _lib.check(L.dbb_pca_gram_host(vp(A), n, w, 0, vp(mean), vp(std), vp(gram)))
        return gram + n * np.outer(mean, mean)

    @property
    def U(self):
        """Left singular vectors of the retained components (pca.py:63): scores / d."""
        n = self._scores.shape[1]
        return self._scores * self.dinv[:n]

    def pc(self):
        """pca.py:77-80: U[:, :npc] * d[:npc]."""
        return self._scores

    def vars_pc(self, x):
        n = self._scores.shape[1]
        return self.d[:n] * np.dot(self.Vt[:n], x.T).T

    def pc_vars(self, p):
        n = self._scores.shape[1]
        return np.dot(self.Vt[:n].T, (self.dinv[:n] * p).T).T

    def save(self, outputFile, seqIds, pc, variance):
        np.savez(outputFile, seqIds=seqIds, pc=pc, variance=variance)

    def loadFile(self, inputFile):
        npzfile = np.load(inputFile)
        return npzfile['seqIds'], npzfile['pc'], npzfile['variance']
