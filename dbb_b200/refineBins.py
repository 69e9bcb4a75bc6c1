"""Rectangular filtered nearest-core search ("next" row N1): the numeric core of the reference's
``dbb/refineBins.py`` (``RefineBins.__workerThread`` :42-96, core statistics :288-304), also the inner loop of
``unbinned.py:137-176``, ``merge.py`` and ``compare.py``.  File handling, conflict resolution and reporting stay with
the reference's driver code (out of scope); this module provides the two pieces those drivers
compute: the per-bin statistics and, for every unbinned contig, the closest valid core in GC / TD / coverage space.
"""
import ctypes
import logging

import numpy as np

from . import _lib
from .distributions import Distributions, nearest_index, _sorted_keys

UNBINNED = -1                                    # greedy.py Greedy.UNBINNED


class Core(object):
    """greedy.py:147-157."""

    def __init__(self, binId, length, GC, coverage, tetraSig):
        self.binId, self.length, self.GC, self.coverage, self.tetraSig = binId, length, GC, coverage, tetraSig
        self.contigs = None


def coreStatistics(binnedContigs, numKmers=_lib.NUM_KMERS):
    """refineBins.py:288-304: length-weighted RUNNING means of GC, coverage and signature per bin, in the order the
    contigs are given (the update is order dependent, so it is kept as the reference writes it).  Host side: one
    pass over the binned contigs."""
    cores = {}
    for contig in binnedContigs:
        if contig.binId not in cores:
            cores[contig.binId] = Core(contig.binId, 0, 0, 0, np.zeros(numKmers))
        core = cores[contig.binId]
        core.length += contig.length
        weight = float(contig.length) / core.length
        core.GC = contig.GC * weight + core.GC * (1.0 - weight)
        core.coverage = contig.coverage * weight + core.coverage * (1.0 - weight)
        core.tetraSig = np.asarray(contig.tetraSig, dtype=np.float64) * weight + core.tetraSig * (1.0 - weight)
    return cores


def searchInput(queryContigs, core_list, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer):
    """The C ABI's ``dbb_refine_input`` for ``queryContigs`` against ``core_list`` (in that order): every nearest-key
    lookup of distributions.py depends on one contig or one core only, so it is resolved here once per object
    instead of once per pair.  Returns (struct, arrays that must stay alive while it is in use)."""
    U, B = len(queryContigs), len(core_list)
    dist = Distributions(gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
    u_len = np.array([c.length for c in queryContigs], dtype=np.float64)
    arr = lambda it, dt=np.float64: np.ascontiguousarray(np.array(list(it), dtype=dt))
    u_sig = np.ascontiguousarray(np.array([c.tetraSig for c in queryContigs], dtype=np.float64).reshape(U, _lib.NUM_KMERS))
    c_sig = np.ascontiguousarray(np.array([c.tetraSig for c in core_list], dtype=np.float64).reshape(B, _lib.NUM_KMERS))
    u_gc, u_cov = arr(c.GC for c in queryContigs), arr(c.coverage for c in queryContigs)
    c_gc, c_cov = arr(c.GC for c in core_list), arr(c.coverage for c in core_list)
    td_lens = _sorted_keys(tdDist)
    u_tdb = np.ascontiguousarray(np.array([tdDist[l][dist.tdKey] for l in td_lens])[nearest_index(td_lens, u_len)])
    means = _sorted_keys(gcDist)
    gc_lens = _sorted_keys(gcDist[means[0]])
    cov_lens = _sorted_keys(covDist)
    u_gcl, u_covl = nearest_index(gc_lens, u_len), nearest_index(cov_lens, u_len)
    c_gcm = nearest_index(means, c_gc)
    gc_lo = np.array([[gcDist[m][l][dist.gcLowerBoundKey] for l in gc_lens] for m in means], dtype=np.float64)
    gc_hi = np.array([[gcDist[m][l][dist.gcUpperBoundKey] for l in gc_lens] for m in means], dtype=np.float64)
    cov_lo = np.array([covDist[l][dist.covLowerBoundKey] for l in cov_lens], dtype=np.float64)
    cov_hi = np.array([covDist[l][dist.covUpperBoundKey] for l in cov_lens], dtype=np.float64)
    keep = (u_sig, c_sig, u_gc, u_cov, c_gc, c_cov, u_tdb, u_gcl, u_covl, c_gcm, gc_lo, gc_hi, cov_lo, cov_hi)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    inp = _lib.RefineInput(U, B, vp(u_sig), vp(c_sig), vp(u_gc), vp(u_cov), vp(c_gc), vp(c_cov), vp(u_tdb),
                           vp(u_gcl), vp(u_covl), vp(c_gcm), vp(gc_lo), vp(gc_hi), gc_lo.shape[0], gc_lo.shape[1],
                           vp(cov_lo), vp(cov_hi), cov_lo.shape[0])
    return inp, keep


class RefineBins(object):
    def __init__(self, bDebug=False):
        self.logger = logging.getLogger()

    def closestCores(self, unbinnedContigs, cores, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer,
                     want_within=False, want_dist=False):
        """For every unbinned contig: (closest int32[U,3] core POSITION in GC / TD / coverage space or -1,
        min_dist float64[U,3], assigned int32[U] position or -1[, within uint8[U,B]][, dist float64[U,B]]).  ``cores``
        is a dict binId -> Core (iterated in ascending binId order, which is CPython-2.7's order for small ints) or a
        list (taken in the given order)."""
        core_list = [cores[k] for k in sorted(cores)] if isinstance(cores, dict) else list(cores)
        U, B = len(unbinnedContigs), len(core_list)
        inp, keep = searchInput(unbinnedContigs, core_list, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        closest = np.full((U, 3), -1, dtype=np.int32)
        mind = np.full((U, 3), 1e9, dtype=np.float64)
        assigned = np.full(U, -1, dtype=np.int32)
        within = np.zeros((U, max(B, 1)), dtype=np.uint8) if want_within else None
        dmat = np.zeros((U, max(B, 1)), dtype=np.float64) if want_dist else None
        if U:
            out = _lib.RefineOutput(_lib.ptr(closest), _lib.ptr(mind), _lib.ptr(assigned), _lib.ptr(within), _lib.ptr(dmat))
            _lib.check(_lib.lib().dbb_refine_host(ctypes.byref(inp), ctypes.byref(out)))
        self.coreOrder = [c.binId for c in core_list]
        res = (closest, mind, assigned)
        if want_within:
            res = res + (within[:, :B],)
        if want_dist:
            res = res + (dmat[:, :B],)
        return res

    def refine(self, unbinnedContigs, cores, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer):
        """What refineBins.py:60-96 + :172-175 leave behind: ``contig.binId`` of every unbinned contig is set to the
        bin of the single core closest in all three spaces, else UNBINNED.  Returns the tuple list the reference's
        writer thread receives: (index, newBinNum, length, bAssigned, bFailedDistanceTest, bInvalidWithAllCores)."""
        closest, _, assigned = self.closestCores(unbinnedContigs, cores, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        out = []
        for i, contig in enumerate(unbinnedContigs):
            ok = assigned[i] >= 0
            newBin = self.coreOrder[assigned[i]] if ok else UNBINNED
            contig.binId = newBin
            out.append((i, newBin, contig.length, bool(ok), bool(not ok and closest[i, 0] >= 0), bool(closest[i, 0] < 0)))
        return out
