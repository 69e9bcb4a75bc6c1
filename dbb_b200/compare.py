"""`dbb compare` of dbb/compare.py (``Compare.run``, compare.py:39-152): does a scaffold belong in a bin?

``Compare().run(preprocessDir, binningFile, queryBinId, queryScaffoldId, gcDistPer, tdDistPer, covDistPer)`` keeps the
reference's signature and prints the reference's report.  The numbers come from the filtered nearest-core search on
the GPU (csrc/refine.cu, ``dbb_refine_host``): one call for the contigs of the query scaffold against all bins
(predicates, closest valid bins, the distance to the query bin), one for the contigs of the query bin against the bin
itself (the mean distance of its members, :107-112).  As in the reference, the unbinned contigs form a bin of their own
(key ``Greedy.UNBINNED``, :88-90) and take part in the closest-bin search; bins are visited in order of first
appearance in ``contigs.seq_stats.tsv``.  Here ``distances = [gcDistance, covDistance, tdDistance]`` (:138), so the
three 'Closest' columns are what their headers say (merge.py and unbinned.py swap the last two).
"""
import logging
import os
import re

import numpy as np

from .distributions import readDistributions
from .genomicSignatures import GenomicSignatures
from .merge import Contig, checkFileExists, readContigIdtoBinId, readSeqStats
from .refineBins import UNBINNED, RefineBins, coreStatistics


class Compare(object):
    def __init__(self, bDebug=False):
        self.logger = logging.getLogger()

    def compareText(self, contigs, queryBinId, queryScaffoldId, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer):
        """``contigs``: all contigs in the order compare.py meets them (binId UNBINNED for unbinned ones).  Returns the
        report compare.py:113-152 prints."""
        queryContigs, queryBinContigs = [], []
        for contig in contigs:
            scaffoldId = contig.contigId
            match = re.search('_c[0-9]*$', contig.contigId)
            if match:
                scaffoldId = contig.contigId[0:match.start()]
            if scaffoldId == queryScaffoldId:
                queryContigs.append(contig)
            if contig.binId == queryBinId:
                queryBinContigs.append(contig)
        cores = coreStatistics(contigs)                                  # :88-98, unbinned contigs included
        bins = list(cores.values())
        q = [b.binId for b in bins].index(queryBinId)
        queryBin = bins[q]
        args = (gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer)
        rb = RefineBins()
        _, _, _, member_dist = rb.closestCores(queryBinContigs, [queryBin], *args, want_dist=True)
        meanDeltaTetra = np.mean(member_dist[:, 0])
        closest, _, _, within, dist = rb.closestCores(queryContigs, bins, *args, want_within=True, want_dist=True)
        out = ['\n', 'Contig Id\tGC\tCoverage\tdelta GC\tdelta Coverage\tdelta tetra\tWithin GC\tWithin coverage\tWithin tetra\t'
               'Closest GC\tClosest coverage\tClosest tetra\n']
        name = lambda p: str(None if p < 0 else bins[p].binId)
        for u, contig in enumerate(queryContigs):
            w = int(within[u, q])
            out.append('%s\t%.1f\t%.1f\t%.1f\t%.1f\t%.2f\t%s\t%s\t%s\t%s\t%s\t%s\n' % (
                contig.contigId, contig.GC * 100, contig.coverage, (contig.GC - queryBin.GC) * 100, contig.coverage - queryBin.coverage,
                dist[u, q], str(bool(w & 1)), str(bool(w & 2)), str(bool(w & 4)),
                name(closest[u, 0]), name(closest[u, 2]), name(closest[u, 1])))      # kernel order: GC, TD, coverage
        out.append('\n')
        out.append('Bin Id = %s, mean GC = %.1f, mean coverage = %.1f, mean delta tetra = %.2f\n' % (queryBinId, queryBin.GC * 100,
                                                                                                queryBin.coverage, meanDeltaTetra))
        return ''.join(out)

    def run(self, preprocessDir, binningFile, queryBinId, queryScaffoldId, gcDistPer, tdDistPer, covDistPer, dataDir=None):
        """compare.py:39-152.  ``dataDir``: where gc_dist.txt / td_dist.txt live (upstream: the package's data/)."""
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
        contigs = []
        for contigId, (contigLen, contigGC, contigCov) in contigStats.items():
            contig = Contig(contigId, contigLen, contigGC, contigCov, tetraSigs[contigId])
            contig.binId = contigIdToBinId.get(contigId, UNBINNED)
            contigs.append(contig)
        print(self.compareText(contigs, queryBinId, queryScaffoldId, gcDist, gcDistPer, tdDist, tdDistPer, covDist, covDistPer), end='')
