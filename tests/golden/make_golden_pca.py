#!/usr/bin/env python
"""Golden vectors for the N4 row: the reference's own ``PCA.pcaMatrix`` (pca.py:56-75, imported unmodified -- the module
is plain numpy and runs under Python 3) on a seeded signature-like matrix.  pca.py:61 calls ``Center(A, bScale)``, i.e.
passes ``bScale`` as Center's ``axis`` argument: with ``bScale=False`` (the only call upstream makes, main.py:327) the
columns are centred AND scaled by their standard deviation (``scale`` keeps its default True); with ``bScale=True``
the call fails (axis 1).  ``bScale`` is given as the integer 0 here because current numpy rejects a bool axis.
Build container only; the output ``ref_pca.npz`` is committed."""
import os
import sys

import numpy as np

REF = '/root/reference/dbb'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
from pca import PCA  # noqa: E402


def main():
    rng = np.random.RandomState(77)
    cent = rng.dirichlet(np.ones(136) * 5, 6)
    member = rng.randint(0, 6, 300)
    A = cent[member] * rng.gamma(60, 1 / 60.0, (300, 136))
    A /= A.sum(1, keepdims=True)
    out = {'A': A}
    for tag, frac, c, s in (('c', 1.0, True, 0), ('c90', 0.90, True, 0), ('raw', 0.95, False, 0)):
        p = PCA()
        pc, var = p.pcaMatrix(A.copy(), frac, c, s)
        out[tag + '_pc'] = pc[:, :8]
        out[tag + '_pc_shape'] = np.array(pc.shape)
        out[tag + '_variance'] = var
        out[tag + '_d'] = p.d
        out[tag + '_Vt'] = p.Vt[:8]
        out[tag + '_npc'] = np.array([p.npc])
        print(tag, 'npc', p.npc, 'variance[:3]', var[:3])
    np.savez_compressed(os.path.join(HERE, 'ref_pca.npz'), **out)


if __name__ == '__main__':
    main()
