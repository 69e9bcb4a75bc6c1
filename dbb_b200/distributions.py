"""Threshold tables: host-side mirror of the reference's ``dbb/distributions.py``.

Same class and method names (``Distributions``, ``boundsDistGC``, ``withinDistGC``, ``boundDistTD``,
``witinDistTD`` [sic], ``boundsDistCov``, ``withinDistCov``, ``gcDistance``, ``covDistance``) so call
sites such as refineBins.py:60-78 keep working, plus ``PairTables``: the per-scaffold resolution of
every bound (SURVEY.md section 8a "algebraic restatement") that the CUDA pair kernel consumes.  The
reference resolves the nearest table key PER PAIR (distributions.py:76-110 -> __findNearest :135-138);
every such lookup depends on one scaffold only, so it is done once per scaffold here.

Frozen tie rule: nearest key over keys in ASCENDING order, first minimum wins (lower key at an exact
midpoint).  CPython-2.7 dict order, which the reference inherits, is not emulated.
"""
import ast
import logging
import os

import numpy as np


def readDistributions(preprocessDir, gcDistPer=None, tdDistPer=None, covDistPer=None, dataDir=None):
    """distributions.py:26-48.  ``dataDir`` replaces the reference's package ``data/`` directory (the
    upstream gc_dist.txt is not shipped with the reference checkout)."""
    dataDir = dataDir or os.path.join(os.path.dirname(os.path.realpath(__file__)), 'data')
    out = []
    for path in (os.path.join(dataDir, 'gc_dist.txt'), os.path.join(dataDir, 'td_dist.txt'),
                 os.path.join(preprocessDir, 'coverage_dist.txt')):
        if not os.path.exists(path):
            logging.getLogger().error('  [Error] Input file does not exists: ' + path + '\n')
            raise SystemExit(1)                            # common.py:24-28 (log + sys.exit)
        with open(path, 'r') as f:
            out.append(ast.literal_eval(f.read()))
    return tuple(out)


def _sorted_keys(d):
    return sorted(d.keys())


def _find_nearest(keys, value):
    """distributions.py:135-138 over ascending keys."""
    arr = np.asarray(keys, dtype=np.float64)
    return keys[int(np.abs(arr - value).argmin())]


def nearest_index(keys_sorted, values):
    """Vectorised ``__findNearest``: index of the nearest key for every value (first minimum)."""
    keys = np.asarray(keys_sorted, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    out = np.empty(v.shape[0], dtype=np.int32)
    step = max(1, (1 << 22) // max(1, keys.shape[0]))
    for s in range(0, v.shape[0], step):
        out[s:s + step] = np.abs(keys[None, :] - v[s:s + step, None]).argmin(axis=1)
    return out


class Distributions(object):
    """distributions.py:51-145."""

    def __init__(self, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer):
        self.logger = logging.getLogger()
        if gcDist is not None:
            self.gcDist = gcDist
            sampleMeanGC = _sorted_keys(gcDist)[0]
            sampleSeqLen = _sorted_keys(gcDist[sampleMeanGC])[0]
            d = gcDist[sampleMeanGC][sampleSeqLen]
            self.gcLowerBoundKey = _find_nearest(_sorted_keys(d), (100 - gcDistPer) / 2.0)
            self.gcUpperBoundKey = _find_nearest(_sorted_keys(d), (100 + gcDistPer) / 2.0)
        if tdDist is not None:
            self.tdDist = tdDist
            self.tdKey = _find_nearest(_sorted_keys(tdDist[_sorted_keys(tdDist)[0]]), tdDistPer)
        if covDist is not None:
            self.covDist = covDist
            d = covDist[_sorted_keys(covDist)[0]]
            self.covLowerBoundKey = _find_nearest(_sorted_keys(d), (100 - covDistPer) / 2.0)
            self.covUpperBoundKey = _find_nearest(_sorted_keys(d), (100 + covDistPer) / 2.0)

    def boundsDistGC(self, unbinnedContig, core, fixedSeqLen=None):
        closestMeanGC = _find_nearest(_sorted_keys(self.gcDist), core.GC)
        lens = _sorted_keys(self.gcDist[closestMeanGC])
        closestSeqLen = _find_nearest(lens, unbinnedContig.length if fixedSeqLen is None else fixedSeqLen)
        d = self.gcDist[closestMeanGC][closestSeqLen]
        return d[self.gcLowerBoundKey], d[self.gcUpperBoundKey]

    def withinDistGC(self, unbinnedContig, core, fixedSeqLen=None):
        lowerBound, upperBound = self.boundsDistGC(unbinnedContig, core, fixedSeqLen)
        gcDiff = unbinnedContig.GC - core.GC
        return bool(gcDiff >= lowerBound and gcDiff <= upperBound)

    def boundDistTD(self, unbinnedContig, fixedSeqLen=None):
        closestSeqLen = _find_nearest(_sorted_keys(self.tdDist),
                                      unbinnedContig.length if fixedSeqLen is None else fixedSeqLen)
        return self.tdDist[closestSeqLen][self.tdKey]

    def witinDistTD(self, tetraDist, unbinnedContig, fixedSeqLen=None):
        return bool(tetraDist < self.boundDistTD(unbinnedContig, fixedSeqLen))

    def boundsDistCov(self, unbinnedContig, core, fixedSeqLen=None):
        closestSeqLen = _find_nearest(_sorted_keys(self.covDist),
                                      unbinnedContig.length if fixedSeqLen is None else fixedSeqLen)
        d = self.covDist[closestSeqLen]
        return d[self.covLowerBoundKey], d[self.covUpperBoundKey]

    def withinDistCov(self, unbinnedContig, core, fixedSeqLen=None):
        lowerBound, upperBound = self.boundsDistCov(unbinnedContig, core, fixedSeqLen)
        if core.coverage > 0:
            covPerDiff = (unbinnedContig.coverage - core.coverage) * 100.0 / core.coverage
        else:
            covPerDiff = (unbinnedContig.coverage - 0.1) * 100.0 / 0.1
        return bool(covPerDiff >= lowerBound and covPerDiff <= upperBound)

    def gcDistance(self, unbinnedContig, core):
        return abs(unbinnedContig.GC - core.GC)

    def covDistance(self, unbinnedContig, core):
        return abs(unbinnedContig.coverage - core.coverage)


class PairTables(object):
    """Per-scaffold inputs of the pair kernel (fields of ``dbb_pairs_input``), all numpy, host side.

    length int32[N]; td_bound float32[N]; gc float64[N], gc_mbin/gc_lbin int32[N], gc_lo/gc_hi
    float64[nm, nl]; cov float64[N], cov_lbin int32[N], cov_lo/cov_hi float64[nl].  Unused groups are
    None.  With ``fixedSeqLen`` every length bin is the bin of that one length (greedy.py:395-403)."""

    def __init__(self, length, GC=None, coverage=None, tdDist=None, tdDistPer=None, gcDist=None,
                 gcDistPer=None, covDist=None, covDistPer=None, fixedSeqLen=None):
        length = np.asarray(length)
        if length.size and (length.max() > np.iinfo(np.int32).max or length.min() < 0):
            raise ValueError('sequence lengths must fit int32')
        self.n = int(length.shape[0])
        self.length = np.ascontiguousarray(length, dtype=np.int32)
        lookup_len = length if fixedSeqLen is None else np.full(length.shape, fixedSeqLen)
        dist = Distributions(gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        self.td_bound = self.td_bound64 = None
        self.gc = self.gc_mbin = self.gc_lbin = self.gc_lo = self.gc_hi = None
        self.cov = self.cov_lbin = self.cov_lo = self.cov_hi = None
        if tdDist is not None:
            lens = _sorted_keys(tdDist)
            tab = np.array([tdDist[l][dist.tdKey] for l in lens], dtype=np.float64)
            self.td_bound64 = tab[nearest_index(lens, lookup_len)]
            self.td_bound = np.ascontiguousarray(self.td_bound64, dtype=np.float32)
        if gcDist is not None:
            if GC is None:
                raise ValueError('gcDist given without GC values')
            means = _sorted_keys(gcDist)
            lens = _sorted_keys(gcDist[means[0]])
            for m in means:
                if _sorted_keys(gcDist[m]) != lens:
                    raise ValueError('gcDist: every mean-GC entry must carry the same length keys')
            self.gc = np.ascontiguousarray(GC, dtype=np.float64)
            self.gc_mbin = nearest_index(means, self.gc)
            self.gc_lbin = nearest_index(lens, lookup_len)
            self.gc_lo = np.array([[gcDist[m][l][dist.gcLowerBoundKey] for l in lens] for m in means], dtype=np.float64)
            self.gc_hi = np.array([[gcDist[m][l][dist.gcUpperBoundKey] for l in lens] for m in means], dtype=np.float64)
        if covDist is not None:
            if coverage is None:
                raise ValueError('covDist given without coverage values')
            lens = _sorted_keys(covDist)
            self.cov = np.ascontiguousarray(coverage, dtype=np.float64)
            self.cov_lbin = nearest_index(lens, lookup_len)
            self.cov_lo = np.array([covDist[l][dist.covLowerBoundKey] for l in lens], dtype=np.float64)
            self.cov_hi = np.array([covDist[l][dist.covUpperBoundKey] for l in lens], dtype=np.float64)
