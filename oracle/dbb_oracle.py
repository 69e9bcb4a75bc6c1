"""CPU oracle for DBB's hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  Nothing under ``dbb_b200/`` imports it: the product path is CUDA only
and fails loudly when ``libdbbcuda.so`` is missing.

This is a Python 3 restatement of the reference's arithmetic (the reference is Python 2 and cannot
be imported as-is).  Every function cites the ``/root/reference/dbb`` lines it follows.  The
"literal" functions keep the reference's loop shape (per-window dict lookup, per-pair Python loop)
and are the authority; the ``*_np`` functions are vectorised forms tested against the literal ones.

Pinning: upstream ships no tests or golden vectors.  The oracle is pinned against the reference's
OWN source files executed unmodified under Python 3 with interpreter shims
(``tests/golden/make_golden.py`` -> ``tests/golden/*.json|npz``; see that script's header).

Frozen decisions where the reference leaves behaviour to CPython-2.7 dict order
(``distributions.py:135-138`` ``__findNearest`` = first ``argmin`` in ``dict.keys()`` order):
keys are taken in ASCENDING order, so at an exact midpoint between two keys the LOWER key wins.
"""

import numpy as np

K = 4

# ----------------------------------------------------------------------------------------------
# genomicSignatures.py:35-84 -- canonical k-mer table
# ----------------------------------------------------------------------------------------------

_COMPL = str.maketrans('ACGT', 'TGCA')          # genomicSignatures.py:39 (py2: string.maketrans)


def rev_comp(seq):
    """genomicSignatures.py:81-84."""
    return seq.translate(_COMPL)[::-1]


def lexicographically_lowest(seq):
    """genomicSignatures.py:74-79."""
    rseq = rev_comp(seq)
    if seq < rseq:
        return seq
    return rseq


def make_kmer_col_names(k=K):
    """genomicSignatures.py:44-72.  Returns (kmerCols, kmerToCanonicalIndex)."""
    baseWords = ("A", "C", "G", "T")
    mers = ["A", "C", "G", "T"]
    for _ in range(1, k):
        workingList = []
        for mer in mers:
            for char in baseWords:
                workingList.append(mer + char)
        mers = workingList

    retList = []
    for mer in mers:
        kmer = lexicographically_lowest(mer)
        if kmer not in retList:
            retList.append(kmer)

    # genomicSignatures.py:64 calls sorted(retList) and discards the result (no-op)

    kmerToCanonicalIndex = {}
    for index, kmer in enumerate(retList):
        kmerToCanonicalIndex[kmer] = index
        kmerToCanonicalIndex[rev_comp(kmer)] = index
    return retList, kmerToCanonicalIndex


KMER_COLS, KMER_TO_INDEX = make_kmer_col_names(K)
NUM_KMERS = len(KMER_COLS)                       # 136


def code_lut256():
    """2-bit code (A=0,C=1,G=2,T=3, first base most significant) -> canonical column."""
    lut = np.zeros(256, dtype=np.uint8)
    for code in range(256):
        kmer = ''.join('ACGT'[(code >> (2 * (3 - p))) & 3] for p in range(4))
        lut[code] = KMER_TO_INDEX[kmer]
    return lut


# ----------------------------------------------------------------------------------------------
# genomicSignatures.py:134-150 -- counting and normalisation
# ----------------------------------------------------------------------------------------------

def seq_counts_literal(seq):
    """genomicSignatures.py:135-144: the integer count vector the reference computes but never
    returns.  ``seq`` is a str; any window holding a character outside uppercase ACGT raises
    KeyError in the reference and is skipped."""
    sig = [0] * NUM_KMERS
    numMers = len(seq) - K + 1
    for i in range(0, numMers):
        try:
            kmerIndex = KMER_TO_INDEX[seq[i:i + K]]
            sig[kmerIndex] += 1
        except KeyError:
            pass
    return sig


def seq_signature_literal(seq):
    """genomicSignatures.py:134-150 (0/0 -> NaN, as numpy does in the reference)."""
    sig = np.array(seq_counts_literal(seq), dtype=float)
    with np.errstate(invalid='ignore', divide='ignore'):
        sig /= np.sum(sig)
    return sig


_BASE_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in zip(b'ACGT', range(4)):
    _BASE_CODE[_c] = _v
_LUT256 = code_lut256()


def seq_counts_np(seq):
    """Vectorised form of seq_counts_literal.  ``seq`` is str, bytes or a uint8 array."""
    if isinstance(seq, str):
        seq = seq.encode('latin-1')
    a = np.frombuffer(seq, dtype=np.uint8) if not isinstance(seq, np.ndarray) else seq
    if a.size < K:
        return np.zeros(NUM_KMERS, dtype=np.int64)
    c = _BASE_CODE[a]
    bad = (c == 255)
    c = (c & 3).astype(np.uint16)
    code = (c[0:-3] << 6) | (c[1:-2] << 4) | (c[2:-1] << 2) | c[3:]
    ok = ~(bad[0:-3] | bad[1:-2] | bad[2:-1] | bad[3:])
    return np.bincount(_LUT256[code[ok]], minlength=NUM_KMERS).astype(np.int64)


def seq_signature_np(seq):
    cnt = seq_counts_np(seq).astype(float)
    with np.errstate(invalid='ignore', divide='ignore'):
        return cnt / np.sum(cnt)


def distance(sig1, sig2):
    """genomicSignatures.py:183-184."""
    return np.sum(np.abs(sig1 - sig2))


# ----------------------------------------------------------------------------------------------
# genomicSignatures.py:97-126, 186-194 -- TSV format
# ----------------------------------------------------------------------------------------------

def format_sig_value(x):
    """py2 ``str(np.float64)`` (genomicSignatures.py:119 ``map(str, sig)``) prints 12 significant
    digits; py3 prints the shortest round-trip repr.  The port keeps the py2 text."""
    s = '%.12g' % x
    if 'e' not in s and '.' not in s and 'n' not in s and 'i' not in s:
        s += '.0'                                   # py2 str(1.0) == '1.0'
    return s


def tsv_header():
    """genomicSignatures.py:101-105."""
    return 'Sequence Id' + ''.join('\t' + kmer for kmer in KMER_COLS) + '\n'


def tsv_row(seqId, sig):
    """genomicSignatures.py:118-120."""
    return seqId + '\t' + '\t'.join(map(format_sig_value, sig)) + '\n'


def read_tsv(path):
    """genomicSignatures.py:186-194."""
    sig = {}
    with open(path) as f:
        next(f)
        for line in f:
            lineSplit = line.split('\t')
            sig[lineSplit[0]] = np.array([float(x) for x in lineSplit[1:]])
    return sig


# ----------------------------------------------------------------------------------------------
# distributions.py:51-145
# ----------------------------------------------------------------------------------------------

def _keys(d):
    """py2 ``dict.keys()`` is a list in hash order; frozen here to ascending order."""
    return sorted(d.keys())


def find_nearest(array, value):
    """distributions.py:135-138 (first minimum in array order)."""
    idx = (np.abs(np.array(array) - value)).argmin()
    return array[idx]


def in_range(testValue, lowerBound, upperBound):
    """distributions.py:140-145 (inclusive on both sides)."""
    if testValue >= lowerBound and testValue <= upperBound:
        return True
    return False


class Distributions(object):
    """distributions.py:51-133."""

    def __init__(self, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer):
        if gcDist is not None:                                                  # :57-63
            self.gcDist = gcDist
            sampleMeanGC = _keys(gcDist)[0]
            sampleSeqLen = _keys(gcDist[sampleMeanGC])[0]
            d = gcDist[sampleMeanGC][sampleSeqLen]
            self.gcLowerBoundKey = find_nearest(_keys(d), (100 - gcDistPer) / 2.0)
            self.gcUpperBoundKey = find_nearest(_keys(d), (100 + gcDistPer) / 2.0)

        if tdDist is not None:                                                  # :65-67
            self.tdDist = tdDist
            self.tdKey = find_nearest(_keys(tdDist[_keys(tdDist)[0]]), tdDistPer)

        if covDist is not None:                                                 # :69-73
            self.covDist = covDist
            d = covDist[_keys(covDist)[0]]
            self.covLowerBoundKey = find_nearest(_keys(d), (100 - covDistPer) / 2.0)
            self.covUpperBoundKey = find_nearest(_keys(d), (100 + covDistPer) / 2.0)

    def boundsDistGC(self, unbinnedContig, core, fixedSeqLen=None):             # :75-87
        closestMeanGC = find_nearest(_keys(self.gcDist), core.GC)
        if fixedSeqLen is None:
            closestSeqLen = find_nearest(_keys(self.gcDist[closestMeanGC]), unbinnedContig.length)
        else:
            closestSeqLen = find_nearest(_keys(self.gcDist[closestMeanGC]), fixedSeqLen)
        d = self.gcDist[closestMeanGC][closestSeqLen]
        return d[self.gcLowerBoundKey], d[self.gcUpperBoundKey]

    def withinDistGC(self, unbinnedContig, core, fixedSeqLen=None):             # :89-93
        lowerBound, upperBound = self.boundsDistGC(unbinnedContig, core, fixedSeqLen)
        gcDiff = unbinnedContig.GC - core.GC
        return in_range(gcDiff, lowerBound, upperBound)

    def boundDistTD(self, unbinnedContig, fixedSeqLen=None):                    # :95-101
        if fixedSeqLen is None:
            closestSeqLen = find_nearest(_keys(self.tdDist), unbinnedContig.length)
        else:
            closestSeqLen = find_nearest(_keys(self.tdDist), fixedSeqLen)
        return self.tdDist[closestSeqLen][self.tdKey]

    def witinDistTD(self, tetraDist, unbinnedContig, fixedSeqLen=None):         # :103-104 (strict <)
        return tetraDist < self.boundDistTD(unbinnedContig, fixedSeqLen)

    def boundsDistCov(self, unbinnedContig, core, fixedSeqLen=None):            # :106-116
        if fixedSeqLen is None:
            closestSeqLen = find_nearest(_keys(self.covDist), unbinnedContig.length)
        else:
            closestSeqLen = find_nearest(_keys(self.covDist), fixedSeqLen)
        d = self.covDist[closestSeqLen]
        return d[self.covLowerBoundKey], d[self.covUpperBoundKey]

    def withinDistCov(self, unbinnedContig, core, fixedSeqLen=None):            # :118-127
        lowerBound, upperBound = self.boundsDistCov(unbinnedContig, core, fixedSeqLen)
        if core.coverage > 0:
            covPerDiff = (unbinnedContig.coverage - core.coverage) * 100.0 / core.coverage
        else:
            covPerDiff = (unbinnedContig.coverage - 0.1) * 100.0 / 0.1
        return in_range(covPerDiff, lowerBound, upperBound)


# ----------------------------------------------------------------------------------------------
# tdNeighbours.py:41-51, gcNeighbours.py:43-52, coverageNeighbours.py:41-50 -- one row each
# ----------------------------------------------------------------------------------------------

class Contig(object):
    """Shape of greedy.py:133-141 (the element type of ``seqs``)."""

    def __init__(self, contigId, length, GC, coverage, tetraSig):
        self.contigId = contigId
        self.length = length
        self.GC = GC
        self.coverage = coverage
        self.tetraSig = tetraSig


def td_row_literal(seqI, seqs, distributions, fixedSeqLen):
    """tdNeighbours.py:41-51."""
    row = np.zeros(len(seqs))
    for j, seqJ in enumerate(seqs):
        dist = np.sum(np.abs(seqI.tetraSig - seqJ.tetraSig))
        if seqI.length > seqJ.length:
            bWithin = distributions.witinDistTD(dist, seqJ, fixedSeqLen)
        else:
            bWithin = distributions.witinDistTD(dist, seqI, fixedSeqLen)
        if bWithin:
            row[j] = 1
    return row


def gc_row_literal(seqI, seqs, distributions, fixedSeqLen):
    """gcNeighbours.py:42-52."""
    row = np.zeros(len(seqs))
    for j, seqJ in enumerate(seqs):
        if seqI.length > seqJ.length:
            core = seqI
            unbinned = seqJ
        else:
            core = seqJ
            unbinned = seqI
        if distributions.withinDistGC(unbinned, core, fixedSeqLen):
            row[j] = 1
    return row


def cov_row_literal(seqI, seqs, distributions, fixedSeqLen):
    """coverageNeighbours.py:40-50."""
    row = np.zeros(len(seqs))
    for j, seqJ in enumerate(seqs):
        if seqI.length > seqJ.length:
            core = seqI
            unbinned = seqJ
        else:
            core = seqJ
            unbinned = seqI
        if distributions.withinDistCov(unbinned, core, fixedSeqLen):
            row[j] = 1
    return row


# ----------------------------------------------------------------------------------------------
# Vectorised (numpy, float64) forms of the three rows, over arrays instead of objects.
# They follow the algebraic restatement: every bound depends only on per-scaffold scalars.
# ----------------------------------------------------------------------------------------------

def nearest_index_sorted(keys_sorted, values):
    """find_nearest over ascending keys for an array of values (lowest key wins ties)."""
    keys = np.asarray(keys_sorted, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64)
    return np.abs(keys[None, :] - v.reshape(-1, 1)).argmin(axis=1).reshape(v.shape)


class DistTables(object):
    """Flattened view of the nested dicts used by the vectorised oracle (and mirrored by the
    host code in dbb_b200/distributions.py, which is tested against this)."""

    def __init__(self, gcDist=None, gcDistPer=None, tdDist=None, tdDistPer=None,
                 covDist=None, covDistPer=None):
        d = Distributions(gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        if tdDist is not None:
            self.tdLens = np.array(_keys(tdDist), dtype=np.float64)
            self.tdBound = np.array([tdDist[l][d.tdKey] for l in _keys(tdDist)], dtype=np.float64)
        if covDist is not None:
            self.covLens = np.array(_keys(covDist), dtype=np.float64)
            self.covLo = np.array([covDist[l][d.covLowerBoundKey] for l in _keys(covDist)], dtype=np.float64)
            self.covHi = np.array([covDist[l][d.covUpperBoundKey] for l in _keys(covDist)], dtype=np.float64)
        if gcDist is not None:
            self.gcMeans = np.array(_keys(gcDist), dtype=np.float64)
            # the reference looks lengths up inside gcDist[closestMeanGC]; every mean-GC entry is
            # assumed to carry the same length keys (true for the shipped td table's sibling)
            self.gcLens = np.array(_keys(gcDist[_keys(gcDist)[0]]), dtype=np.float64)
            self.gcLo = np.array([[gcDist[m][l][d.gcLowerBoundKey] for l in _keys(gcDist[m])]
                                  for m in _keys(gcDist)], dtype=np.float64)
            self.gcHi = np.array([[gcDist[m][l][d.gcUpperBoundKey] for l in _keys(gcDist[m])]
                                  for m in _keys(gcDist)], dtype=np.float64)


def l1_rows_np(sig, rows):
    """float64 L1 distances of ``rows`` against all N (tdNeighbours.py:43)."""
    sig = np.asarray(sig, dtype=np.float64)
    out = np.empty((len(rows), sig.shape[0]), dtype=np.float64)
    for n, i in enumerate(rows):
        out[n] = np.abs(sig - sig[i]).sum(axis=1)
    return out


def td_bound_rows_np(length, rows, tables, fixedSeqLen):
    """Per-pair TD bound: table entry of the SHORTER scaffold (ties -> i), or of fixedSeqLen."""
    length = np.asarray(length)
    if fixedSeqLen is not None:
        b = tables.tdBound[nearest_index_sorted(tables.tdLens, np.array([fixedSeqLen]))[0]]
        return np.full((len(rows), length.shape[0]), b)
    per = tables.tdBound[nearest_index_sorted(tables.tdLens, length)]
    out = np.empty((len(rows), length.shape[0]), dtype=np.float64)
    for n, i in enumerate(rows):
        out[n] = np.where(length[i] > length, per, per[i])
    return out


def td_rows_np(sig, length, rows, tables, fixedSeqLen):
    with np.errstate(invalid='ignore'):
        return (l1_rows_np(sig, rows) < td_bound_rows_np(length, rows, tables, fixedSeqLen)).astype(np.float64)


def gc_rows_np(length, GC, rows, tables, fixedSeqLen):
    length = np.asarray(length)
    GC = np.asarray(GC, dtype=np.float64)
    mbin = nearest_index_sorted(tables.gcMeans, GC)
    if fixedSeqLen is not None:
        lbin = np.full(length.shape, nearest_index_sorted(tables.gcLens, np.array([fixedSeqLen]))[0])
    else:
        lbin = nearest_index_sorted(tables.gcLens, length)
    out = np.empty((len(rows), length.shape[0]), dtype=np.float64)
    for n, i in enumerate(rows):
        i_is_core = length[i] > length                         # else core = j
        core_m = np.where(i_is_core, mbin[i], mbin)
        unb_l = np.where(i_is_core, lbin, lbin[i])
        lo = tables.gcLo[core_m, unb_l]
        hi = tables.gcHi[core_m, unb_l]
        diff = np.where(i_is_core, GC - GC[i], GC[i] - GC)     # unbinned.GC - core.GC
        out[n] = ((diff >= lo) & (diff <= hi)).astype(np.float64)
    return out


def cov_rows_np(length, cov, rows, tables, fixedSeqLen):
    length = np.asarray(length)
    cov = np.asarray(cov, dtype=np.float64)
    if fixedSeqLen is not None:
        lbin = np.full(length.shape, nearest_index_sorted(tables.covLens, np.array([fixedSeqLen]))[0])
    else:
        lbin = nearest_index_sorted(tables.covLens, length)
    out = np.empty((len(rows), length.shape[0]), dtype=np.float64)
    for n, i in enumerate(rows):
        i_is_core = length[i] > length
        core_c = np.where(i_is_core, cov[i], cov)
        unb_c = np.where(i_is_core, cov, cov[i])
        unb_l = np.where(i_is_core, lbin, lbin[i])
        eff = np.where(core_c > 0, core_c, 0.1)
        diff = (unb_c - eff) * 100.0 / eff
        out[n] = ((diff >= tables.covLo[unb_l]) & (diff <= tables.covHi[unb_l])).astype(np.float64)
    return out


def knn_rows_np(sig, rows, k, allowed=None):
    """Top-k definition of the new kNN output: for row i the candidates are
    j != i (restricted to ``allowed[n]`` if given), ordered by (float64 L1 asc, j asc).
    Returns (idx int64[len(rows), k] padded -1, dist float64 padded +inf, full distance rows)."""
    d = l1_rows_np(sig, rows)
    n_all = d.shape[1]
    idx = np.full((len(rows), k), -1, dtype=np.int64)
    dist = np.full((len(rows), k), np.inf, dtype=np.float64)
    for n, i in enumerate(rows):
        ok = ~np.isnan(d[n])
        ok[i] = False
        if allowed is not None:
            ok &= allowed[n].astype(bool)
        cand = np.nonzero(ok)[0]
        order = cand[np.lexsort((cand, d[n][cand]))][:k]
        idx[n, :order.size] = order
        dist[n, :order.size] = d[n][order]
    return idx, dist, d


# ----------------------------------------------------------------------------------------------
# "next" row N1 -- refineBins.py:42-96 (one unbinned contig against all cores) and :288-304
# ----------------------------------------------------------------------------------------------

class Core(object):
    """greedy.py:147-157."""

    def __init__(self, binId, length, GC, coverage, tetraSig):
        self.binId, self.length, self.GC, self.coverage, self.tetraSig = binId, length, GC, coverage, tetraSig


def core_statistics_literal(binnedContigs, numKmers=NUM_KMERS):
    """refineBins.py:288-304 (running length-weighted means, list-of-floats signature as the reference keeps it)."""
    cores = {}
    for contig in binnedContigs:
        if contig.binId not in cores:
            cores[contig.binId] = Core(contig.binId, 0, 0, 0, [0] * numKmers)
        core = cores[contig.binId]
        core.length += contig.length
        weight = float(contig.length) / core.length
        core.GC = contig.GC * weight + core.GC * (1.0 - weight)
        core.coverage = contig.coverage * weight + core.coverage * (1.0 - weight)
        for i in range(0, len(core.tetraSig)):
            core.tetraSig[i] = contig.tetraSig[i] * weight + core.tetraSig[i] * (1.0 - weight)
    return cores


def refine_worker_literal(unbinnedContig, cores, distributions):
    """refineBins.py:52-94 for one contig; ``cores`` is iterated in ascending binId order.  Returns
    (closestBin[3], minDistances[3], newBinNum, bAssigned, bFailedDistanceTest, bInvalidWithAllCores)."""
    closestBin = [None] * 3
    minDistances = [1e9] * 3
    for binId in sorted(cores):
        core = cores[binId]
        if not distributions.withinDistGC(unbinnedContig, core):
            continue
        if not distributions.withinDistCov(unbinnedContig, core):
            continue
        tdDistance = distance(unbinnedContig.tetraSig, np.asarray(core.tetraSig))
        if not distributions.witinDistTD(tdDistance, unbinnedContig):
            continue
        distances = [abs(unbinnedContig.GC - core.GC), tdDistance, abs(unbinnedContig.coverage - core.coverage)]
        for i, d in enumerate(distances):
            if d < minDistances[i]:
                minDistances[i] = d
                closestBin[i] = binId
    bAssigned = bFailed = bInvalid = False
    if closestBin[0] is not None:
        if closestBin[0] == closestBin[1] and closestBin[0] == closestBin[2]:
            bAssigned, newBinNum = True, cores[closestBin[0]].binId
        else:
            bFailed, newBinNum = True, -1
    else:
        bInvalid, newBinNum = True, -1
    return closestBin, minDistances, newBinNum, bAssigned, bFailed, bInvalid


# ----------------------------------------------------------------------------------------------
# N5: bin-to-bin merge table -- merge.py:66-156 (statistics of the bins, then every contig of bin 1 against bin 2
# and against all bins but its own)
# ----------------------------------------------------------------------------------------------

MERGE_HEADER = ('Src Bin Id\tGC\tCoverage\tDest Bin Id\tGC\tCoverage\tContigs in Src Bin\tWithin GC\tWithin coverage\t'
                'Within tetra\tClosest GC\tClosest coverage\tClosest tetra\tMerging score\n')          # merge.py:95


def merge_table_literal(contigs, distributions, numKmers=NUM_KMERS):
    """merge.py:66-156 for ``contigs`` in the order merge.py meets them (objects with contigId, length, GC, coverage,
    tetraSig and binId; binId -1 = unbinned, skipped as :70-71 does).  Bins are visited in order of first appearance
    (the frozen stand-in for CPython 2.7's dict order).  Returns the text merge.py writes.  Kept as upstream has it:
    ``distances = [gcDistance, tdDistance, covDistance]`` (:140), so the column headed 'Closest coverage' counts
    the bins closest in TD space and 'Closest tetra' those closest in coverage."""
    bins = {}
    binnedContigs = {}
    for contig in contigs:
        binId = contig.binId
        if binId != -1:
            binnedContigs.setdefault(binId, []).append(contig)
            if binId not in bins:
                bins[binId] = Core(binId, 0, 0, 0, [0] * numKmers)
            b = bins[binId]
            b.length += contig.length
            weight = float(contig.length) / b.length
            b.GC = contig.GC * weight + b.GC * (1.0 - weight)
            b.coverage = contig.coverage * weight + b.coverage * (1.0 - weight)
            for i in range(0, len(b.tetraSig)):
                b.tetraSig[i] = contig.tetraSig[i] * weight + b.tetraSig[i] * (1.0 - weight)
    sig_of = dict((k, np.asarray(v.tetraSig, dtype=np.float64)) for k, v in bins.items())
    out = [MERGE_HEADER]
    for binId1, bin1 in bins.items():
        cs = binnedContigs[binId1]
        for binId2, bin2 in bins.items():
            if binId1 == binId2:
                continue
            out.append('%s\t%.1f\t%.1f\t%s\t%.1f\t%.1f\t%d' % (binId1, bin1.GC * 100, bin1.coverage, binId2, bin2.GC * 100,
                                                              bin2.coverage, len(cs)))
            numWithinGC = numWithinCov = numWithinTetra = closestGC = closestCov = closestTetra = 0
            for contig in cs:
                tdDistanceQuery = distance(contig.tetraSig, sig_of[binId2])
                if distributions.withinDistGC(contig, bin2):
                    numWithinGC += 1
                if distributions.withinDistCov(contig, bin2):
                    numWithinCov += 1
                if distributions.witinDistTD(tdDistanceQuery, contig):
                    numWithinTetra += 1
                closestBin = [None] * 3
                minDistances = [1e9] * 3
                for binId, b in bins.items():
                    if binId == binId1:
                        continue
                    if not distributions.withinDistGC(contig, b):
                        continue
                    if not distributions.withinDistCov(contig, b):
                        continue
                    tdDistance = distance(contig.tetraSig, sig_of[binId])
                    if not distributions.witinDistTD(tdDistance, contig):
                        continue
                    distances = [abs(contig.GC - b.GC), tdDistance, abs(contig.coverage - b.coverage)]
                    for i, d in enumerate(distances):
                        if d < minDistances[i]:
                            minDistances[i] = d
                            closestBin[i] = binId
                if closestBin[0] == binId2:
                    closestGC += 1
                if closestBin[1] == binId2:
                    closestCov += 1
                if closestBin[2] == binId2:
                    closestTetra += 1
            out.append('\t%d\t%d\t%d\t%d\t%d\t%d' % (numWithinGC, numWithinCov, numWithinTetra, closestGC, closestCov, closestTetra))
            out.append('\t%.2f\n' % (float(closestGC + closestCov + closestTetra) / (3 * len(cs))))
    return ''.join(out)


# ----------------------------------------------------------------------------------------------
# N6: statistics of unbinned scaffolds -- unbinned.py:70-186
# ----------------------------------------------------------------------------------------------

UNBINNED_HEADER = ('Scaffold Id\tLength\tGC\tCoverage\tBin Id\tGC\tCoverage\tContigs in Scaffold\tWithin GC\tWithin coverage\t'
                   'Within tetra\tClosest GC\tClosest coverage\tClosest tetra\tBinning score\n')       # unbinned.py:110


def unbinned_table_literal(contigs, distributions, minScaffoldLen, minContigLen, allScaffolds, numKmers=NUM_KMERS):
    """unbinned.py:70-186 for ``contigs`` in the order unbinned.py meets them (contigId, length, GC, coverage, tetraSig,
    binId; -1 = unbinned).  Scaffolds and bins are visited in order of first appearance.  Returns the text it writes."""
    import re
    bins = {}
    unbinnedContigs = {}
    for contig in contigs:
        scaffoldId = contig.contigId
        match = re.search('_c[0-9]*$', contig.contigId)
        if match:
            scaffoldId = contig.contigId[0:match.start()]
        if contig.length < minContigLen:
            continue
        if contig.binId == -1:
            unbinnedContigs.setdefault(scaffoldId, []).append(contig)
        else:
            if contig.binId not in bins:
                bins[contig.binId] = Core(contig.binId, 0, 0, 0, [0] * numKmers)
            b = bins[contig.binId]
            b.length += contig.length
            weight = float(contig.length) / b.length
            b.GC = contig.GC * weight + b.GC * (1.0 - weight)
            b.coverage = contig.coverage * weight + b.coverage * (1.0 - weight)
            for i in range(0, len(b.tetraSig)):
                b.tetraSig[i] = contig.tetraSig[i] * weight + b.tetraSig[i] * (1.0 - weight)
            if allScaffolds:
                unbinnedContigs.setdefault(scaffoldId, []).append(contig)
    sig_of = dict((k, np.asarray(v.tetraSig, dtype=np.float64)) for k, v in bins.items())
    out = [UNBINNED_HEADER]
    for scaffoldId, cs in unbinnedContigs.items():
        for queryBinId, queryBin in bins.items():
            numWithinGC = numWithinCov = numWithinTetra = closestGC = closestCov = closestTetra = binningScore = 0
            scaffoldGC = scaffoldCov = scaffoldLen = 0
            for contig in cs:
                scaffoldLen += contig.length
                weight = float(contig.length) / scaffoldLen
                scaffoldGC = contig.GC * weight + scaffoldGC * (1.0 - weight)
                scaffoldCov = contig.coverage * weight + scaffoldCov * (1.0 - weight)
                tdDistanceQuery = distance(contig.tetraSig, sig_of[queryBinId])
                bInDists = True
                if distributions.withinDistGC(contig, queryBin):
                    numWithinGC += 1
                else:
                    bInDists = False
                if distributions.withinDistCov(contig, queryBin):
                    numWithinCov += 1
                else:
                    bInDists = False
                if distributions.witinDistTD(tdDistanceQuery, contig):
                    numWithinTetra += 1
                else:
                    bInDists = False
                closestBin = [None] * 3
                minDistances = [1e9] * 3
                if bInDists:
                    for binId, b in bins.items():
                        if not distributions.withinDistGC(contig, b):
                            continue
                        if not distributions.withinDistCov(contig, b):
                            continue
                        tdDistance = distance(contig.tetraSig, sig_of[binId])
                        if not distributions.witinDistTD(tdDistance, contig):
                            continue
                        distances = [abs(contig.GC - b.GC), tdDistance, abs(contig.coverage - b.coverage)]
                        for i, d in enumerate(distances):
                            if d < minDistances[i]:
                                minDistances[i] = d
                                closestBin[i] = binId
                if closestBin[0] == queryBinId:
                    closestGC += 1
                if closestBin[1] == queryBinId:
                    closestCov += 1
                if closestBin[2] == queryBinId:
                    closestTetra += 1
                if closestBin[1] == queryBinId and closestBin[2] == queryBinId:
                    binningScore += 1
            if scaffoldLen >= minScaffoldLen:
                out.append('%s\t%d\t%.1f\t%.1f\t%s\t%.1f\t%.1f\t%d' % (scaffoldId, scaffoldLen, scaffoldGC * 100, scaffoldCov, queryBinId,
                                                                       queryBin.GC * 100, queryBin.coverage, len(cs)))
                out.append('\t%d\t%d\t%d\t%d\t%d\t%d' % (numWithinGC, numWithinCov, numWithinTetra, closestGC, closestCov, closestTetra))
                out.append('\t%.2f\n' % (float(binningScore) / len(cs)))
    return ''.join(out)


# ----------------------------------------------------------------------------------------------
# compare.py:70-152 (one scaffold against one bin; what `dbb compare` prints)
# ----------------------------------------------------------------------------------------------

def compare_text_literal(contigs, queryBinId, queryScaffoldId, distributions, numKmers=NUM_KMERS):
    """compare.py:70-152 for ``contigs`` in the order compare.py meets them (binId -1 = unbinned: compare.py makes a bin
    of those too, :88-90).  Bins are visited in order of first appearance.  Returns the text it prints."""
    import re
    queryContigs, queryBinContigs, bins = [], [], {}
    for contig in contigs:
        scaffoldId = contig.contigId
        match = re.search('_c[0-9]*$', contig.contigId)
        if match:
            scaffoldId = contig.contigId[0:match.start()]
        if scaffoldId == queryScaffoldId:
            queryContigs.append(contig)
        if contig.binId not in bins:
            bins[contig.binId] = Core(contig.binId, 0, 0, 0, [0] * numKmers)
        b = bins[contig.binId]
        b.length += contig.length
        weight = float(contig.length) / b.length
        b.GC = contig.GC * weight + b.GC * (1.0 - weight)
        b.coverage = contig.coverage * weight + b.coverage * (1.0 - weight)
        for i in range(0, len(b.tetraSig)):
            b.tetraSig[i] = contig.tetraSig[i] * weight + b.tetraSig[i] * (1.0 - weight)
        if contig.binId == queryBinId:
            queryBinContigs.append(contig)
    sig_of = dict((k, np.asarray(v.tetraSig, dtype=np.float64)) for k, v in bins.items())
    queryBin = bins[queryBinId]
    meanDeltaTetra = np.mean([distance(c.tetraSig, sig_of[queryBinId]) for c in queryBinContigs])
    out = ['\n', 'Contig Id\tGC\tCoverage\tdelta GC\tdelta Coverage\tdelta tetra\tWithin GC\tWithin coverage\tWithin tetra\t'
           'Closest GC\tClosest coverage\tClosest tetra\n']
    for contig in queryContigs:
        tdDistanceQuery = distance(contig.tetraSig, sig_of[queryBinId])
        bWithinGC = distributions.withinDistGC(contig, queryBin)
        bWithinCov = distributions.withinDistCov(contig, queryBin)
        bWithinTetra = distributions.witinDistTD(tdDistanceQuery, contig)
        closestBin = [None] * 3
        minDistances = [1e9] * 3
        for binId, b in bins.items():
            if not distributions.withinDistGC(contig, b):
                continue
            if not distributions.withinDistCov(contig, b):
                continue
            tdDistance = distance(contig.tetraSig, sig_of[binId])
            if not distributions.witinDistTD(tdDistance, contig):
                continue
            distances = [abs(contig.GC - b.GC), abs(contig.coverage - b.coverage), tdDistance]      # compare.py:138
            for i, d in enumerate(distances):
                if d < minDistances[i]:
                    minDistances[i] = d
                    closestBin[i] = binId
        out.append('%s\t%.1f\t%.1f\t%.1f\t%.1f\t%.2f\t%s\t%s\t%s\t%s\t%s\t%s\n' % (
            contig.contigId, contig.GC * 100, contig.coverage, (contig.GC - queryBin.GC) * 100, contig.coverage - queryBin.coverage,
            tdDistanceQuery, str(bool(bWithinGC)), str(bool(bWithinCov)), str(bool(bWithinTetra)), str(closestBin[0]), str(closestBin[1]),
            str(closestBin[2])))
    out.append('\n')
    out.append('Bin Id = %s, mean GC = %.1f, mean coverage = %.1f, mean delta tetra = %.2f\n' % (queryBinId, queryBin.GC * 100,
                                                                                            queryBin.coverage, meanDeltaTetra))
    return ''.join(out)


# ----------------------------------------------------------------------------------------------
# N2: sequence side of preprocess -- splitScaffolds.py:27-53, seqUtils.py:258-265,
# preprocess.py:94-100, :118-161 (what a worker computes from the sequence alone)
# ----------------------------------------------------------------------------------------------

def splits_literal(scaffoldId, scaffold, minN):
    """splitScaffolds.py:27-53, statement for statement.  Returns an insertion-ordered dict
    contigId -> [start, end] (the reference's py2 dict has no defined order; callers sort)."""
    contigs = {}
    contigCount = 0
    nCount = 0
    startIndex = 0
    nStart = 0
    for curIndex, ch in enumerate(scaffold):
        if ch == 'N' or ch == 'n':
            if nCount == 0:
                nStart = curIndex
            nCount += 1
        else:
            if nCount > minN:
                contigs[scaffoldId + '_c' + str(contigCount)] = [startIndex, nStart]
                contigCount += 1
                startIndex = curIndex
            nCount = 0
    if startIndex != len(scaffold) - 1:
        if nCount == 0:
            contigs[scaffoldId + '_c' + str(contigCount)] = [startIndex, len(scaffold)]
        else:
            contigs[scaffoldId + '_c' + str(contigCount)] = [startIndex, nStart]
    return contigs


def base_count_literal(seq):
    """seqUtils.py:258-265: case-insensitive, U counted as T."""
    testSeq = seq.upper()
    a = testSeq.count('A')
    c = testSeq.count('C')
    g = testSeq.count('G')
    t = testSeq.count('T') + testSeq.count('U')
    return a, c, g, t


def calculate_gc_literal(seq):
    """preprocess.py:94-100 (ZeroDivisionError when the sequence has no A/C/G/T/U, as upstream)."""
    a, c, g, t = base_count_literal(seq)
    gc = g + c
    at = a + t
    return float(gc) / (gc + at)


def sequence_stats_literal(scaffoldId, scaffoldSeq, minN):
    """The sequence-derived part of preprocess.py:118-161 for one scaffold: returns
    (scaffoldStats [len, GC, sig], contigStats {contigId: [len, GC, sig, start, end]}); the single-contig case
    stores the scaffold's own statistics under the scaffold id (:160-161; its trailing [0, scaffoldSeq] pair is
    reported as [0, len])."""
    contigs = splits_literal(scaffoldId, scaffoldSeq, minN)
    scaffoldStats = [len(scaffoldSeq), calculate_gc_literal(scaffoldSeq), seq_signature_literal(scaffoldSeq)]
    contigStats = {}
    if len(contigs) > 1:
        for contigId, split in contigs.items():
            sub = scaffoldSeq[split[0]:split[1]]
            contigStats[contigId] = [split[1] - split[0], calculate_gc_literal(sub), seq_signature_literal(sub),
                                     split[0], split[1]]
    else:
        contigStats[scaffoldId] = scaffoldStats + [0, len(scaffoldSeq)]
    return scaffoldStats, contigStats


# ----------------------------------------------------------------------------------------------
# N3: greedy core growth -- greedy.py:167-251 (Greedy.__greedy)
# ----------------------------------------------------------------------------------------------

def greedy_literal(sortedContigs, minBinSize, distributions):
    """greedy.py:172-251, statement for statement (progress output dropped; ``genomicSig.distance`` is
    ``distance`` above).  Sets ``bProcessed`` / ``binId`` on the contigs, returns the list of cores."""
    for contig in sortedContigs:
        contig.bProcessed = False
    numCores = 0
    cores = []
    for contigIndex, contig in enumerate(sortedContigs):
        if contig.bProcessed:
            continue
        contig.bProcessed = True
        contig.binId = numCores
        contigsInCurCore = [contig]
        curCore = Core(numCores, contig.length, contig.GC, contig.coverage, np.array(contig.tetraSig, dtype=float))
        bCoreExpanded = True
        while bCoreExpanded:
            closestContig = None
            closestAbsCovDiff = 1e12
            for curContig in sortedContigs[contigIndex + 1:]:
                if curContig.bProcessed:
                    continue
                absCovDiff = abs(curContig.coverage - curCore.coverage)
                if absCovDiff > closestAbsCovDiff:
                    continue
                if not distributions.withinDistGC(curContig, curCore):
                    continue
                if not distributions.withinDistCov(curContig, curCore):
                    continue
                tdDistance = distance(curContig.tetraSig, curCore.tetraSig)
                if not distributions.witinDistTD(tdDistance, curContig):
                    continue
                closestAbsCovDiff = absCovDiff
                closestContig = curContig
            if closestContig is not None:
                bCoreExpanded = True
                closestContig.bProcessed = True
                closestContig.binId = numCores
                contigsInCurCore.append(closestContig)
                curCore.length += closestContig.length
                weight = float(closestContig.length) / curCore.length
                curCore.GC = closestContig.GC * weight + curCore.GC * (1.0 - weight)
                curCore.coverage = closestContig.coverage * weight + curCore.coverage * (1.0 - weight)
                for i in range(0, len(curCore.tetraSig)):
                    curCore.tetraSig[i] = closestContig.tetraSig[i] * weight + curCore.tetraSig[i] * (1.0 - weight)
            else:
                bCoreExpanded = False
                if curCore.length > minBinSize:
                    numCores += 1
                    curCore.contigs = contigsInCurCore
                    cores.append(curCore)
                else:
                    for contig in contigsInCurCore:
                        contig.bProcessed = False
                        contig.binId = -1
    return cores


def merge_literal(cores, distributions):
    """greedy.py:258-333 (``Greedy.__merge``), statement for statement (progress output dropped; ``genomicSig.distance``
    is ``distance`` above; ``distributions`` is built by the caller with mergeDistPer for all three tables as :256 does).
    Sets ``bProcessed`` / ``binId`` on the input cores and ``binId`` on their contigs, returns the merged cores."""
    for core in cores:
        core.bProcessed = False
    numMergedCores = 0
    mergedCores = []
    for coreIndex, core in enumerate(cores):
        if core.bProcessed:
            continue
        core.bProcessed = True
        curMergedCore = Core(numMergedCores, core.length, core.GC, core.coverage, core.tetraSig)
        curMergedCore.contigs = core.contigs
        bCoreMerged = True
        while bCoreMerged:
            closestCore = None
            closestAbsCovDiff = 1e12
            for curCore in cores[coreIndex + 1:]:
                if curCore.bProcessed:
                    continue
                absCovDiff = abs(curCore.coverage - curMergedCore.coverage)
                if absCovDiff > closestAbsCovDiff:
                    continue
                if not distributions.withinDistGC(curCore, curMergedCore):
                    continue
                if not distributions.withinDistCov(curCore, curMergedCore):
                    continue
                tdDistance = distance(curCore.tetraSig, curMergedCore.tetraSig)
                if not distributions.witinDistTD(tdDistance, curCore):
                    continue
                closestAbsCovDiff = absCovDiff
                closestCore = curCore
            if closestCore is not None:
                bCoreMerged = True
                closestCore.bProcessed = True
                closestCore.binId = numMergedCores
                curMergedCore.contigs += closestCore.contigs
                curMergedCore.length += closestCore.length
                weight = float(closestCore.length) / curMergedCore.length
                curMergedCore.GC = closestCore.GC * weight + curMergedCore.GC * (1.0 - weight)
                curMergedCore.coverage = closestCore.coverage * weight + curMergedCore.coverage * (1.0 - weight)
                for i in range(0, len(curCore.tetraSig)):
                    curMergedCore.tetraSig[i] = closestCore.tetraSig[i] * weight + curMergedCore.tetraSig[i] * (1.0 - weight)
            else:
                bCoreMerged = False
                numMergedCores += 1
                mergedCores.append(curMergedCore)
    for core in mergedCores:
        for contig in core.contigs:
            contig.binId = core.binId
    return mergedCores


def neighbourhood_sizes_literal(contigs):
    """greedy.py:406-420 (``Greedy.__sortContigs``): base pairs of the common GC / TD / coverage neighbourhood of every
    contig (attributes set by the three neighbour classes) and the order upstream computes and then drops (F8)."""
    neighbourhoodSizes = []
    for contig in contigs:
        covNeighbours = np.where(contig.covNeighbours == 1)[0]
        gcNeighbours = np.where(contig.gcNeighbours == 1)[0]
        tdNeighbours = np.where(contig.tdNeighbours == 1)[0]
        intersectCovGC = np.intersect1d(covNeighbours, gcNeighbours, assume_unique=True)
        neighbourIndices = np.intersect1d(tdNeighbours, intersectCovGC, assume_unique=True)
        totalNeighbourBPs = sum(contigs[i].length for i in neighbourIndices)
        neighbourhoodSizes.append(totalNeighbourBPs)
    sortedIndices = np.argsort(neighbourhoodSizes)[::-1]
    return neighbourhoodSizes, sortedIndices
