"""Sequence side of dbb/preprocess.py, batched on the GPU.

The reference's worker (preprocess.py:102-205) handles one scaffold at a time: split it at N runs
(``SplitScaffolds.splits``), GC of the scaffold and of every contig (``__calculateGC`` -> ``baseCount``), tetranucleotide
signature of the scaffold and of every contig (``genomicSig.seqSignature``, :141,158).  ``Preprocess.sequenceStatistics``
computes exactly those values for ALL scaffolds with one upload: one pack launch (codes, validity, N plane), two
splitter launches, one pack launch for the contigs (slices of the same buffer), one counting launch set each, one
base-count launch.  Coverage comes from BAM files through pysam and stays on the CPU (out of scope): the
``writeStatistics`` helper takes it as an argument and writes the four TSV files in the reference's format
(:207-226, :266-271 and the contig loop after it).
"""
import logging

import numpy as np

from .genomicSignatures import _format_value


class Preprocess(object):
    def __init__(self):
        self.logger = logging.getLogger()

    @staticmethod
    def _gc(acgt, seq_id):
        a, c, g, t = (int(x) for x in acgt)
        gc, at = g + c, a + t
        if gc + at == 0:
            # preprocess.py:100 divides by zero here (an empty contig, or one without A/C/G/T/U)
            raise ZeroDivisionError('float division by zero (no A/C/G/T/U in %s)' % seq_id)
        return float(gc) / (gc + at)

    def sequenceStatistics(self, scaffolds, minN):
        """scaffolds: dict id -> sequence (or list of (id, sequence)).  Returns (scaffoldSeqStats, contigSeqStats):
        ``scaffoldSeqStats[id] = [length, GC, tetraSig]`` and ``contigSeqStats[id][contigId] = [length, GC, tetraSig,
        start, end]`` -- the sequence-derived fields of preprocess.py:142,159,161 (coverage is the caller's)."""
        import torch
        from . import device
        items = list(scaffolds.items()) if isinstance(scaffolds, dict) else list(scaffolds)
        ids = [i for i, _ in items]
        raw = [s.encode('latin-1') if isinstance(s, str) else bytes(s) for _, s in items]
        n = len(raw)
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(r) for r in raw])
        flat = np.frombuffer(b''.join(raw) + b'\0\0\0\0', dtype=np.uint8)
        device.require_cuda()
        d_ascii = torch.from_numpy(flat.copy()).cuda()

        packed = device.pack_ranges(d_ascii, off[:-1], off[1:], want_nmask=True)
        c_off, c_begin, c_end = device.split_scaffolds(packed, minN)
        n_contigs = np.diff(c_off)
        multi = np.nonzero(n_contigs > 1)[0]                       # only these get per-contig statistics (:146)
        sel = np.concatenate([np.arange(c_off[s], c_off[s + 1]) for s in multi]) if multi.size else np.zeros(0, np.int64)
        owner = np.repeat(multi, n_contigs[multi]) if multi.size else np.zeros(0, np.int64)
        abs_b = off[owner] + c_begin[sel]
        abs_e = off[owner] + c_end[sel]

        def count(pk):
            c = torch.zeros((max(pk.n, 1), 136), dtype=torch.int32, device=d_ascii.device)
            plan = device.CountPlan(pk)
            plan.run(c, None)
            plan.close()
            return c[:pk.n].cpu().numpy().astype(np.int64)

        s_counts = count(packed)
        c_counts = count(device.pack_ranges(d_ascii, abs_b, abs_e)) if sel.size else np.zeros((0, 136), np.int64)
        acgt = device.base_counts(d_ascii, np.concatenate([off[:-1], abs_b]), np.concatenate([off[1:], abs_e]))

        def sig(c):
            v = c.astype(float)
            with np.errstate(invalid='ignore', divide='ignore'):
                v /= np.sum(v)                                     # genomicSignatures.py:147-148
            return v

        scaffoldSeqStats, contigSeqStats = {}, {}
        for i, sid in enumerate(ids):
            scaffoldSeqStats[sid] = [int(off[i + 1] - off[i]), self._gc(acgt[i], sid), sig(s_counts[i])]
            contigSeqStats[sid] = {}
        for j in range(sel.size):
            s = int(owner[j])
            k = int(sel[j] - c_off[s])
            cid = ids[s] + '_c' + str(k)
            b, e = int(c_begin[sel[j]]), int(c_end[sel[j]])
            contigSeqStats[ids[s]][cid] = [e - b, self._gc(acgt[n + j], cid), sig(c_counts[j]), b, e]
        for i, sid in enumerate(ids):
            if n_contigs[i] <= 1:                                  # :160-161
                contigSeqStats[sid][sid] = scaffoldSeqStats[sid] + [0, int(off[i + 1] - off[i])]
        return scaffoldSeqStats, contigSeqStats

    def writeStatistics(self, genomicSig, scaffoldSeqStats, contigSeqStats, coverage, mappedReads, scaffoldSeqStatsFile,
                        scaffoldTetraSigFile, contigSeqStatsFile, contigTetraSigFile):
        """The four TSV files of preprocess.py:207-226,266-271 (+ the contig loop).  ``coverage``: dict id -> float for
        scaffolds and contigs, ``mappedReads``: dict scaffold id -> int (both BAM-derived, supplied by the caller)."""
        fmt = lambda v: _format_value(v) if isinstance(v, float) else str(v)
        with open(scaffoldSeqStatsFile, 'w') as sOut, open(scaffoldTetraSigFile, 'w') as sSig, \
                open(contigSeqStatsFile, 'w') as cOut, open(contigTetraSigFile, 'w') as cSig:
            sOut.write('Scaffold Id\tLength\tGC\tCoverage (mean depth)\tMapped reads\n')
            sSig.write('Scaffold Id' + ''.join('\t' + k for k in genomicSig.canonicalKmerOrder()) + '\n')
            cOut.write('Contig Id\tLength\tGC\tCoverage (mean depth)\n')
            cSig.write('Contig Id' + ''.join('\t' + k for k in genomicSig.canonicalKmerOrder()) + '\n')
            for sid, st in scaffoldSeqStats.items():
                sOut.write(sid + '\t' + str(st[0]) + '\t' + fmt(st[1]) + '\t' + fmt(float(coverage[sid])) + '\t' +
                           str(mappedReads.get(sid, 0)) + '\n')
                sSig.write(sid + '\t' + '\t'.join(_format_value(v) for v in st[2]) + '\n')
                for cid, cs in contigSeqStats[sid].items():
                    cOut.write(cid + '\t' + str(cs[0]) + '\t' + fmt(cs[1]) + '\t' + fmt(float(coverage[cid])) + '\n')
                    cSig.write(cid + '\t' + '\t'.join(_format_value(v) for v in cs[2]) + '\n')
