"""Statistics of unbinned scaffolds of dbb/unbinned.py (``Unbinned.run``, unbinned.py:35-190), GPU assisted (row N6).

``Unbinned().run(preprocessDir, binningFile, gcDistPer, tdDistPer, covDistPer, minScaffoldLen, minContigLen,
allScaffolds, outputFile)`` keeps the reference's signature and writes the reference's file.  For every scaffold under
study and every bin, unbinned.py walks the scaffold's contigs: three predicates against the bin and -- for the contigs
within all three -- the closest valid bin in three spaces over ALL bins, recomputed for every query bin (contigs x
bins^2 distance evaluations).  Here it is one launch of the filtered nearest-core search (csrc/refine.cu,
``dbb_unbinned_table_host``): queries = the contigs, cores = the bins, flags and closest bins tallied on the device
into an int32 [scaffolds][bins][7] table.  "Closest bin == query bin" already implies that the contig is within all
three distributions of the query bin (closest bins are chosen among the valid ones), so the reference's ``bInDists``
guard needs no separate pass.  Running bin and scaffold statistics (:92-101, :126-132) are order-dependent host loops.

Frozen where the reference leaves it to CPython 2.7's dict order for string keys: scaffolds and bins are visited in
order of first appearance in ``contigs.seq_stats.tsv``.  Kept as upstream has it: ``distances = [gcDistance,
tdDistance, covDistance]`` (:169), so 'Closest coverage' counts TD space and 'Closest tetra' coverage space; the
binning score counts contigs whose closest bin in both of those is the query bin (:181-182).
"""
import ctypes
import logging
import os
import re

import numpy as np

from . import _lib
from .distributions import readDistributions
from .genomicSignatures import GenomicSignatures
from .merge import Contig, checkFileExists, readContigIdtoBinId, readSeqStats
from .refineBins import UNBINNED, coreStatistics, searchInput

HEADER = ('Scaffold Id\tLength\tGC\tCoverage\tBin Id\tGC\tCoverage\tContigs in Scaffold\tWithin GC\tWithin coverage\t'
          'Within tetra\tClosest GC\tClosest coverage\tClosest tetra\tBinning score\n')                   # unbinned.py:110


class Scaffold(object):
    """One entry of unbinned.py's ``unbinnedContigs`` with the running statistics of :126-132."""

    def __init__(self, scaffoldId):
        self.scaffoldId, self.contigs, self.length, self.GC, self.coverage = scaffoldId, [], 0, 0, 0

    def add(self, contig):
        self.contigs.append(contig)
        self.length += contig.length
        weight = float(contig.length) / self.length
        self.GC = contig.GC * weight + self.GC * (1.0 - weight)
        self.coverage = contig.coverage * weight + self.coverage * (1.0 - weight)


class Unbinned(object):
    def __init__(self, bDebug=False):
        self.logger = logging.getLogger()

    def unbinnedTable(self, scaffolds, binnedContigs, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer):
        """``scaffolds``: list of ``Scaffold`` (their contigs are the queries); ``binnedContigs``: the contigs that
        define the bins, in the order unbinned.py meets them.  Returns (bins, table): ``bins`` the ``Core`` statistics
        in visiting order, ``table`` int32 [scaffolds][bins][7] (include/dbb_cuda.h, ``dbb_unbinned_input``)."""
        cores = coreStatistics(binnedContigs)
        bins = list(cores.values())
        queries = [c for s in scaffolds for c in s.contigs]
        group = np.ascontiguousarray(np.array([g for g, s in enumerate(scaffolds) for _ in s.contigs], dtype=np.int32))
        table = np.zeros((len(scaffolds), len(bins), 7), dtype=np.int32)
        if queries and bins:
            search, keep = searchInput(queries, bins, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
            inp = _lib.UnbinnedInput(search, _lib.ptr(group), len(scaffolds))
            _lib.check(_lib.lib().dbb_unbinned_table_host(ctypes.byref(inp), _lib.ptr(table)))
            del keep
        return bins, table

    def formatTable(self, scaffolds, bins, table, minScaffoldLen):
        """The text unbinned.py:109-186 writes."""
        out = [HEADER]
        for g, s in enumerate(scaffolds):
            if s.length < minScaffoldLen:
                continue
            for b, queryBin in enumerate(bins):
                t = table[g, b]
                out.append('%s\t%d\t%.1f\t%.1f\t%s\t%.1f\t%.1f\t%d' % (s.scaffoldId, s.length, s.GC * 100, s.coverage, queryBin.binId,
                                                                  queryBin.GC * 100, queryBin.coverage, len(s.contigs)))
                out.append('\t%d\t%d\t%d\t%d\t%d\t%d' % (t[0], t[1], t[2], t[3], t[4], t[5]))
                out.append('\t%.2f\n' % (float(t[6]) / len(s.contigs)))
        return ''.join(out)

    def run(self, preprocessDir, binningFile, gcDistPer, tdDistPer, covDistPer, minScaffoldLen, minContigLen, allScaffolds,
            outputFile, dataDir=None):
        """unbinned.py:35-190.  ``dataDir``: where gc_dist.txt / td_dist.txt live (upstream: the package's data/)."""
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
        scaffolds, binned = {}, []
        for contigId, (contigLen, contigGC, contigCov) in contigStats.items():
            binId = contigIdToBinId.get(contigId, UNBINNED)
            scaffoldId = contigId
            match = re.search('_c[0-9]*$', contigId)
            if match:
                scaffoldId = contigId[0:match.start()]
            contig = Contig(contigId, contigLen, contigGC, contigCov, tetraSigs[contigId])
            contig.binId = binId
            if contigLen < minContigLen:
                continue
            if binId != UNBINNED:
                binned.append(contig)
            if binId == UNBINNED or allScaffolds:
                if scaffoldId not in scaffolds:
                    scaffolds[scaffoldId] = Scaffold(scaffoldId)
                scaffolds[scaffoldId].add(contig)
        scaffolds = list(scaffolds.values())
        bins, table = self.unbinnedTable(scaffolds, binned, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        with open(outputFile, 'w') as fout:
            fout.write(self.formatTable(scaffolds, bins, table, minScaffoldLen))
