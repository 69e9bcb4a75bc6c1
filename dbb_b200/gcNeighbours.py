"""Drop-in for the reference's ``dbb/gcNeighbours.py`` (class ``GcNeighbours``).

Contract (gcNeighbours.py:29-105, distributions.py:75-93): core = the LONGER of (i, j), ties -> j;
``row_i[j] = 1`` iff ``lo <= unbinned.GC - core.GC <= hi`` with (lo, hi) looked up by (nearest mean-GC key
of the core, nearest length key of the unbinned one or of fixedSeqLen, percentile keys (100 -/+ per)/2).
FP64 arithmetic on device, so the mask is exact."""
import logging

import numpy as np

from . import _pairs
from .distributions import PairTables


class GcNeighbours(object):
    DENSE_LIMIT = 8192

    def __init__(self, fixedSeqLen):
        self.logger = logging.getLogger()
        self.fixedSeqLen = fixedSeqLen

    def run(self, seqs, gcDist, gcDistPer, threads=1):
        if len(seqs) == 0:
            return
        length = np.array([s.length for s in seqs], dtype=np.int64)
        sig = np.zeros((len(seqs), 136), dtype=np.float32)       # distances are not used by this predicate
        GC = np.array([s.GC for s in seqs], dtype=np.float64)
        tables = PairTables(length, GC=GC, gcDist=gcDist, gcDistPer=gcDistPer, fixedSeqLen=self.fixedSeqLen)
        res = _pairs.run_pairs_host(sig, tables, want_gc_mask=True)
        self.rows = _pairs.attach_rows(seqs, 'gcNeighbours', res['gc_mask'], self.DENSE_LIMIT)
