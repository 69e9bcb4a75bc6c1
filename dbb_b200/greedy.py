"""Greedy core growth, core merging and the neighbourhood-size order of dbb/greedy.py (``Greedy.__greedy``,
greedy.py:167-251; ``__merge``, :253-334; ``__sortContigs``, :390-422), GPU assisted.

``Greedy().greedy(sortedContigs, minBinSize, buildDistPer, gcDist, tdDist, covDist)`` is the reference's private
``__greedy`` with the same arguments (minus ``genomicSig`` / ``numThreads``, which it only used for the L1 distance):
it sets ``contig.bProcessed`` / ``contig.binId`` and returns the list of ``Core`` objects with ``.contigs``.  Element
type of ``sortedContigs``: any object with ``.length, .GC, .coverage, .tetraSig`` (``Contig``, greedy.py:133-141).
The scan over the candidates of each growth step runs on the GPU (csrc/greedy.cu), the sequential outer loop in C++.
"""
import ctypes
import logging

import numpy as np

from . import _lib
from .distributions import Distributions, _sorted_keys, nearest_index
from .refineBins import Core, UNBINNED


class Greedy(object):
    UNBINNED = UNBINNED

    def __init__(self, bDebug=False):
        self.logger = logging.getLogger()
        self.numSteps = 0

    def greedy(self, sortedContigs, minBinSize, buildDistPer, gcDist, tdDist, covDist):
        bin_id, cores = self._grow(sortedContigs, minBinSize, buildDistPer, gcDist, tdDist, covDist)
        for core in cores:
            core.contigs = []
        for i, contig in enumerate(sortedContigs):
            contig.binId = int(bin_id[i])
            contig.bProcessed = bin_id[i] != UNBINNED
            if contig.bProcessed:
                cores[contig.binId].contigs.append(contig)
        # greedy.py:227 appends the members in the order they were added; recover it is not possible from bin ids
        # alone, so the lists here are in contig order (the seed first, as in the reference)
        return cores

    def merge(self, cores, mergeDistPer, gcDist, tdDist, covDist):
        """``Greedy.__merge`` (greedy.py:253-334) with the same arguments (minus ``genomicSig`` / ``numThreads``): walks the
        cores in order; the current merged core absorbs, step by step, the later unprocessed core that is closest in
        coverage among those within the GC / coverage / TD distributions at ``mergeDistPer`` (``withinDistGC(curCore,
        curMergedCore)``, ``withinDistCov``, ``witinDistTD(d, curCore)``), with the running length-weighted means of
        :315-320.  This is the growth loop of ``__greedy`` over cores instead of contigs with every result kept (no
        minimum bin size), so it runs through the same entry point (``dbb_greedy_host``: the scan of each step on the
        GPU).  Sets ``core.bProcessed`` / ``core.binId`` on the input cores and ``contig.binId`` on their contigs;
        returns the merged cores (``.contigs``: the members' contigs in input-core order; upstream appends them in merge
        order).  Not reproduced: upstream's merged core aliases the first member's ``tetraSig`` array and overwrites it."""
        bin_id, merged = self._grow(cores, -1, mergeDistPer, gcDist, tdDist, covDist)
        for m in merged:
            m.contigs = []
        seen = set()
        for i, core in enumerate(cores):
            core.bProcessed = True
            b = int(bin_id[i])
            if b in seen:
                core.binId = b                          # an absorbed core is renumbered (greedy.py:311) ...
            seen.add(b)                                 # ... the first core of a merged bin keeps its old id, as upstream
            merged[b].contigs += list(core.contigs or [])
        for m in merged:
            for contig in m.contigs:
                contig.binId = m.binId
        return merged

    def sortContigs(self, contigs, minSeqLen, buildDistPer, gcDist, tdDist, covDist, numThreads=1, bSort=False):
        """``Greedy.__sortContigs`` (greedy.py:390-422): the three neighbour rows of every contig under the fixed
        sequence length ``minSeqLen`` (attached as ``.gcNeighbours / .tdNeighbours / .covNeighbours`` exactly as the
        reference's three ``run`` calls do, :395-404), the base pairs of their intersection per contig (:407-417) and the
        order by decreasing neighbourhood size (:420).  ONE launch of the pair kernel produces the three masks and the
        neighbourhood sizes (``neighbour_bp``: the same intersection summed on the device).
        Upstream assigns the sorted array to a local and drops it, so ``__greedy`` sees the caller's order;
        that is the default here too (``bSort=False`` returns the contigs unchanged); ``bSort=True`` returns them in the
        order upstream computed.  Returns (contigs, neighbourhoodSizes int64[N], sortedIndices)."""
        from . import _pairs
        from .coverageNeighbours import CoverageNeighbours
        from .distributions import PairTables
        from .gcNeighbours import GcNeighbours
        from .tdNeighbours import TdNeighbours
        n = len(contigs)
        if n == 0:
            return contigs, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        sig, length = _pairs.objects_to_arrays(contigs)
        tables = PairTables(length, GC=np.array([c.GC for c in contigs], dtype=np.float64),
                            coverage=np.array([c.coverage for c in contigs], dtype=np.float64), tdDist=tdDist, tdDistPer=buildDistPer,
                            gcDist=gcDist, gcDistPer=buildDistPer, covDist=covDist, covDistPer=buildDistPer, fixedSeqLen=minSeqLen)
        res = _pairs.run_pairs_host(sig, tables, want_count=True, want_td_mask=True, want_gc_mask=True, want_cov_mask=True)
        _pairs.attach_rows(contigs, 'gcNeighbours', res['gc_mask'], GcNeighbours.DENSE_LIMIT)
        _pairs.attach_rows(contigs, 'tdNeighbours', res['td_mask'], TdNeighbours.DENSE_LIMIT)
        _pairs.attach_rows(contigs, 'covNeighbours', res['cov_mask'], CoverageNeighbours.DENSE_LIMIT)
        sizes = res['neighbour_bp'].astype(np.int64)
        sortedIndices = np.argsort(sizes)[::-1]                     # greedy.py:420
        if bSort:
            contigs = [contigs[i] for i in sortedIndices]
        return contigs, sizes, sortedIndices

    def _grow(self, sortedContigs, minBinSize, buildDistPer, gcDist, tdDist, covDist):
        n = len(sortedContigs)
        dist = Distributions(gcDist, buildDistPer, tdDist, buildDistPer, covDist, buildDistPer)
        arr = lambda it, dt=np.float64: np.ascontiguousarray(np.array(list(it), dtype=dt))
        length = arr((c.length for c in sortedContigs), np.int64)
        sig = np.ascontiguousarray(np.array([c.tetraSig for c in sortedContigs], dtype=np.float64).reshape(n, _lib.NUM_KMERS))
        gc, cov = arr(c.GC for c in sortedContigs), arr(c.coverage for c in sortedContigs)
        td_lens = _sorted_keys(tdDist)
        tdb = np.ascontiguousarray(np.array([tdDist[l][dist.tdKey] for l in td_lens], dtype=np.float64)[nearest_index(td_lens, length)])
        means = _sorted_keys(gcDist)
        gc_lens = _sorted_keys(gcDist[means[0]])
        cov_lens = _sorted_keys(covDist)
        gcl, covl = nearest_index(gc_lens, length), nearest_index(cov_lens, length)
        mean_keys = np.ascontiguousarray(np.array(means, dtype=np.float64))
        gc_lo = np.array([[gcDist[m][l][dist.gcLowerBoundKey] for l in gc_lens] for m in means], dtype=np.float64)
        gc_hi = np.array([[gcDist[m][l][dist.gcUpperBoundKey] for l in gc_lens] for m in means], dtype=np.float64)
        cov_lo = np.array([covDist[l][dist.covLowerBoundKey] for l in cov_lens], dtype=np.float64)
        cov_hi = np.array([covDist[l][dist.covUpperBoundKey] for l in cov_lens], dtype=np.float64)
        bin_id = np.full(max(n, 1), UNBINNED, dtype=np.int32)
        core_len = np.zeros(max(n, 1), dtype=np.int64)
        core_gc, core_cov = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        core_sig = np.zeros((max(n, 1), _lib.NUM_KMERS))
        n_cores, n_steps = ctypes.c_int64(0), ctypes.c_int64(0)
        if n:
            vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            inp = _lib.GreedyInput(n, vp(sig), vp(length), vp(gc), vp(cov), vp(tdb), vp(gcl), vp(covl), vp(mean_keys),
                                   vp(gc_lo), vp(gc_hi), gc_lo.shape[0], gc_lo.shape[1], vp(cov_lo), vp(cov_hi), cov_lo.shape[0])
            _lib.check(_lib.lib().dbb_greedy_host(ctypes.byref(inp), int(minBinSize), vp(bin_id), ctypes.byref(n_cores),
                                                  vp(core_len), vp(core_gc), vp(core_cov), vp(core_sig), ctypes.byref(n_steps)))
        self.numSteps = int(n_steps.value)
        cores = []
        for b in range(int(n_cores.value)):
            core = Core(b, int(core_len[b]), float(core_gc[b]), float(core_cov[b]), core_sig[b].copy())
            cores.append(core)
        return bin_id, cores
