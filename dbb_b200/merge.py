"""Bin-to-bin merge table of dbb/merge.py (``Merge.run``, merge.py:35-158), GPU assisted ("next" row N5).

``Merge().run(preprocessDir, binningFile, gcDistPer, tdDistPer, covDistPer, outputFile)`` keeps the reference's
signature and writes the reference's file.  For every ordered pair of different bins merge.py walks the contigs of the
first bin twice -- against the second bin (three predicates) and against all bins but their own (closest valid bin in
three spaces, recomputed for every second bin) -- with ``genomicSig.distance`` and two ``__findNearest`` per pair.
Here that is ONE launch of the filtered nearest-core search (csrc/refine.cu, ``dbb_merge_table_host``): queries = the
binned contigs, cores = the bins, the contig's own bin skipped, flags and closest bins tallied on the device into an
int32 [B][B][6] table.  Running bin statistics (:66-92) are the order-dependent host loop of
``refineBins.coreStatistics``.

Frozen where the reference leaves it to CPython 2.7's dict order for string keys: bins are visited in order of first
appearance in ``contigs.seq_stats.tsv`` (rows of the output and winners of exact distance ties follow that order).
Kept as upstream has it: ``distances = [gcDistance, tdDistance, covDistance]`` (:140), so the column headed 'Closest
coverage' counts bins closest in TD space and 'Closest tetra' those closest in coverage.
"""
import ctypes
import logging
import os

import numpy as np

from . import _lib
from .distributions import readDistributions
from .genomicSignatures import GenomicSignatures
from .refineBins import UNBINNED, coreStatistics, searchInput

HEADER = ('Src Bin Id\tGC\tCoverage\tDest Bin Id\tGC\tCoverage\tContigs in Src Bin\tWithin GC\tWithin coverage\t'
          'Within tetra\tClosest GC\tClosest coverage\tClosest tetra\tMerging score\n')                   # merge.py:95


class Contig(object):
    """greedy.py:133-141."""

    def __init__(self, contigId, length, GC, coverage, tetraSig):
        self.binId = UNBINNED
        self.contigId, self.length, self.GC, self.coverage, self.tetraSig = contigId, length, GC, coverage, tetraSig


def checkFileExists(inputFile):
    """common.py:24-28."""
    if not os.path.exists(inputFile):
        logging.getLogger().error('  [Error] Input file does not exists: ' + inputFile + '\n')
        raise SystemExit(1)


def readSeqStats(seqStatsFile):
    """seqUtils.py:34-43: id -> [length, GC, coverage], in file order."""
    seqStats = {}
    with open(seqStatsFile) as f:
        f.readline()
        for line in f:
            lineSplit = line.split('\t')
            seqStats[lineSplit[0]] = [int(lineSplit[1]), float(lineSplit[2]), float(lineSplit[3])]
    return seqStats


def readContigIdtoBinId(binningFile):
    """greedy.py:55-66 (a parameter line, a header line, then contig id and bin id per line)."""
    contigIdToBinId = {}
    with open(binningFile) as f:
        f.readline()
        f.readline()
        for line in f:
            lineSplit = line.split('\t')
            contigIdToBinId[lineSplit[0]] = UNBINNED if lineSplit[1] == 'unbinned' else lineSplit[1]
    return contigIdToBinId


class Merge(object):
    def __init__(self, bDebug=False):
        self.logger = logging.getLogger()

    @staticmethod
    def binStatistics(binnedContigs):
        """merge.py:66-92: the bins in visiting order (first appearance) with their running length-weighted statistics and
        ``.contigs``.  Host side."""
        cores = coreStatistics(binnedContigs)                 # insertion order = first appearance
        bins = list(cores.values())
        for b in bins:
            b.contigs = []
        for c in binnedContigs:
            cores[c.binId].contigs.append(c)
        return bins

    def mergeTable(self, binnedContigs, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer, want_closest=False):
        """``binnedContigs``: objects with length, GC, coverage, tetraSig and a binId other than UNBINNED, in the order
        merge.py meets them.  Returns (bins, table[, closest]): ``bins`` the list of ``Core`` statistics in visiting
        order (``.contigs`` = the bin's contigs), ``table`` int32 [B][B][6] with, for source bin b1 and destination b2,
        the number of b1's contigs within the GC / coverage / TD distribution of b2 and the number whose closest valid
        bin in GC / TD / coverage space is b2 (merge.py:112-152); ``closest`` int32 [U][3] bin positions or -1."""
        bins = self.binStatistics(binnedContigs)
        pos = dict((b.binId, n) for n, b in enumerate(bins))
        U, B = len(binnedContigs), len(bins)
        table = np.zeros((B, B, 6), dtype=np.int32)
        closest = np.full((U, 3), -1, dtype=np.int32) if want_closest else None
        if U:
            search, keep = searchInput(binnedContigs, bins, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
            own = np.ascontiguousarray(np.array([pos[c.binId] for c in binnedContigs], dtype=np.int32))
            inp = _lib.MergeInput(search, _lib.ptr(own))
            _lib.check(_lib.lib().dbb_merge_table_host(ctypes.byref(inp), _lib.ptr(table), _lib.ptr(closest)))
            del keep
        return (bins, table, closest) if want_closest else (bins, table)

    def formatTable(self, bins, table):
        """The text merge.py:94-156 writes."""
        out = [HEADER]
        for n1, bin1 in enumerate(bins):
            n = len(bin1.contigs)
            for n2, bin2 in enumerate(bins):
                if n1 == n2:
                    continue
                t = table[n1, n2]
                out.append('%s\t%.1f\t%.1f\t%s\t%.1f\t%.1f\t%d' % (bin1.binId, bin1.GC * 100, bin1.coverage, bin2.binId,
                                                              bin2.GC * 100, bin2.coverage, n))
                out.append('\t%d\t%d\t%d\t%d\t%d\t%d' % (t[0], t[1], t[2], t[3], t[4], t[5]))
                out.append('\t%.2f\n' % (float(int(t[3]) + int(t[4]) + int(t[5])) / (3 * n)))
        return ''.join(out)

    def run(self, preprocessDir, binningFile, gcDistPer, tdDistPer, covDistPer, outputFile, dataDir=None):
        """merge.py:35-158.  ``dataDir``: where gc_dist.txt / td_dist.txt live (upstream: the package's data/)."""
        contigStatsFile = os.path.join(preprocessDir, 'contigs.seq_stats.tsv')
        checkFileExists(contigStatsFile)
        contigTetraFile = os.path.join(preprocessDir, 'contigs.tetra.tsv')
        checkFileExists(contigTetraFile)
        self.logger.info('  Reading contig statistics.')
        contigStats = readSeqStats(contigStatsFile)
        self.logger.info('  Reading contig tetranucleotide signatures.')
        tetraSigs = GenomicSignatures(4, 1).read(contigTetraFile)
        self.logger.info('  Reading core bin assignments.')
        contigIdToBinId = readContigIdtoBinId(binningFile)
        self.logger.info('  Reading GC, TD, and coverage distributions.')
        gcDist, tdDist, covDist = readDistributions(preprocessDir, gcDistPer, tdDistPer, covDistPer, dataDir=dataDir)
        self.logger.info('')
        self.logger.info('  Calculating statistics of bins.')
        binned = []
        for contigId, (contigLen, contigGC, contigCov) in contigStats.items():
            binId = contigIdToBinId.get(contigId, UNBINNED)
            if binId != UNBINNED:
                contig = Contig(contigId, contigLen, contigGC, contigCov, tetraSigs[contigId])
                contig.binId = binId
                binned.append(contig)
        bins, table = self.mergeTable(binned, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        with open(outputFile, 'w') as fout:
            fout.write(self.formatTable(bins, table))
