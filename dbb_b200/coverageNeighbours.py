"""Drop-in for the reference's ``dbb/coverageNeighbours.py`` (class ``CoverageNeighbours``).

Contract (coverageNeighbours.py:28-103, distributions.py:106-127): core = the LONGER of (i, j), ties -> j;
``row_i[j] = 1`` iff ``lo <= (unbinned.coverage - c) * 100.0 / c <= hi`` with c = core.coverage (0.1 when
<= 0) and (lo, hi) from the nearest length key of the unbinned one (or fixedSeqLen).  FP64 on device."""
import logging

import numpy as np

from . import _pairs
from .distributions import PairTables


class CoverageNeighbours(object):
    DENSE_LIMIT = 8192

    def __init__(self, fixedSeqLen=None):
        self.logger = logging.getLogger()
        self.fixedSeqLen = fixedSeqLen

    def run(self, seqs, covDist, covDistPer, threads=1):
        if len(seqs) == 0:
            return
        length = np.array([s.length for s in seqs], dtype=np.int64)
        sig = np.zeros((len(seqs), 136), dtype=np.float32)       # distances are not used by this predicate
        cov = np.array([s.coverage for s in seqs], dtype=np.float64)
        tables = PairTables(length, coverage=cov, covDist=covDist, covDistPer=covDistPer,
                            fixedSeqLen=self.fixedSeqLen)
        res = _pairs.run_pairs_host(sig, tables, want_cov_mask=True)
        self.rows = _pairs.attach_rows(seqs, 'covNeighbours', res['cov_mask'], self.DENSE_LIMIT)
