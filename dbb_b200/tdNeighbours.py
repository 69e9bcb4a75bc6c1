"""Drop-in for the reference's ``dbb/tdNeighbours.py`` (class ``TdNeighbours``), CUDA underneath.

``TdNeighbours(fixedSeqLen).run(seqs, tdDist, tdDistPer, threads)`` keeps the reference contract
(tdNeighbours.py:28-104): for every ordered pair (i, j) incl. i == j, ``row_i[j] = 1`` iff
``sum|sig_i - sig_j| < tdDist[nearest(len of the shorter, or fixedSeqLen)][nearest(tdDistPer)]`` and the
row is attached as ``seqs[i].tdNeighbours``.  ``threads`` is accepted and ignored.  The N x N work runs
in the fused distance kernel (csrc/pairs.cu); only the bit mask comes back.

Precision (stated deviation): the reference compares a float64 distance with a float64 bound; here signatures are
float32, the 136 terms are accumulated in FP32 and the bound is rounded to float32, so ``|d32 - d64| <= 2e-6`` and a pair
whose float64 distance lies within 2e-6 of its bound may fall on the other side (measured on the synthetic configs:
0-4 such pairs per 4 million; tests/ assert exact equality outside that band).  The GC and coverage predicates are
evaluated in FP64 on the device and are exact.

``knn`` is the array-native addition (not in the reference): fused per-row top-k plus the size of the
common neighbourhood that greedy.py:407-417 derives from the three masks.
"""
import logging

import numpy as np

from . import _pairs
from .distributions import PairTables


class TdNeighbours(object):
    DENSE_LIMIT = 8192          # above this many contigs rows are lazy views (dense rows cost 8*N^2 bytes)

    def __init__(self, fixedSeqLen):
        self.logger = logging.getLogger()
        self.fixedSeqLen = fixedSeqLen

    def run(self, seqs, tdDist, tdDistPer, threads=1):
        if len(seqs) == 0:
            return
        sig, length = _pairs.objects_to_arrays(seqs)
        tables = PairTables(length, tdDist=tdDist, tdDistPer=tdDistPer, fixedSeqLen=self.fixedSeqLen)
        res = _pairs.run_pairs_host(sig, tables, want_td_mask=True)
        self.rows = _pairs.attach_rows(seqs, 'tdNeighbours', res['td_mask'], self.DENSE_LIMIT)

    def knn(self, sig, length, GC=None, cov=None, k=20, tdDist=None, tdDistPer=None, gcDist=None,
            gcDistPer=None, covDist=None, covDistPer=None, topk_filter=None, metric='l1', prune=None):
        """Returns (idx int32[N,k] padded -1, dist float32[N,k] padded +inf, count int32[N],
        neighbour_bp int64[N]).  One ctypes call, ``dbb_knn_host`` (include/dbb_cuda.h); no torch.

        Top-k candidates of row i: j != i passing the predicates in ``topk_filter`` (default: the GC
        and coverage filters that are enabled), ordered by (distance asc, j asc).  ``count`` /
        ``neighbour_bp``: size / total length of {j : all enabled predicates hold}, j == i included.
        ``metric``: 'l1' is the reference's Manhattan distance (genomicSignatures.py:183-184); 'l2' (Euclidean) is an
        addition for k-NN callers and cannot be combined with the GC / coverage predicates.
        ``prune``: the exact tile-skipping sweep (same results, about twice as fast on metagenome-like inputs).  It needs a
        TD bound, the Manhattan metric and signature rows that sum to 1 -- frequencies do, raw counts do not; the library
        checks this on the device.  None (default): pruned whenever those conditions hold, otherwise the full sweep on the
        GPU; True: required (ValueError / DbbError when a condition fails); False: always the full sweep."""
        tables = PairTables(length, GC=GC, coverage=cov, tdDist=tdDist, tdDistPer=tdDistPer, gcDist=gcDist,
                            gcDistPer=gcDistPer, covDist=covDist, covDistPer=covDistPer,
                            fixedSeqLen=self.fixedSeqLen)
        if topk_filter is None:
            topk_filter = (_pairs.FILTER_GC if gcDist is not None else 0) | (_pairs.FILTER_COV if covDist is not None else 0)
        can_prune = tdDist is not None and metric == 'l1' and k > 0 and tables.n > 256
        if prune and not can_prune:
            raise ValueError('prune=True needs tdDist, the Manhattan metric, k > 0 and more than two tiles of scaffolds')
        if k == 0 or metric != 'l1':
            res = _pairs.run_pairs_host(sig, tables, k=k, topk_filter=topk_filter, want_count=True, metric=metric)
            return res['idx'], res['dist'], res['count'], res['neighbour_bp']
        res = _pairs.run_knn_host(sig, tables, k=k, topk_filter=topk_filter, prune=(-1 if prune is None else int(bool(prune))))
        return res['idx'], res['dist'], res['count'], res['neighbour_bp']
