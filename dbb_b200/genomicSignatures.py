"""Drop-in for the reference's ``dbb/genomicSignatures.py`` (class ``GenomicSignatures``), CUDA underneath.

Surface kept verbatim (genomicSignatures.py:35-194): ``GenomicSignatures(K, threads)`` with attributes
``K, compl, kmerCols, kmerToCanonicalIndex, totalThreads, logger``; ``canonicalKmerOrder()``,
``numKmers()``, ``seqSignature(seq)``, ``calculate(seqFile, outputFile)``, ``distance(sig1, sig2)``,
``read(path)``.  Additions for array-native callers: ``seqCounts`` and ``signatures``.

Counting always runs in csrc/count.cu through the C ABI; there is no CPU counting path.  Nothing CUDA
happens in ``__init__`` and the object holds no handles, so it can be pickled / forked the way
preprocess.py:395-413 does; the library is loaded lazily in whichever process first computes.
"""
import gzip
import logging
import sys

import numpy as np

from . import _lib


def _read_fasta(path):
    """FASTA -> list of (id, str), parsed exactly as ``seqUtils.readFasta`` (seqUtils.py:46-69) parses it:
    the id is the header up to the first SPACE (tabs stay in it), right-stripped; every sequence line loses its LAST
    character whatever it is (``line[0:-1]``: the newline normally -- a '\r' of a CRLF file stays in the sequence and
    invalidates the windows it touches, an unterminated last line loses a base); a repeated id REPLACES the earlier
    record (``seqs[seqId] = []``); '.gz' is opened with gzip (:49-50).  Any failure -- unreadable file, sequence text
    before the first header -- is the reference's log + ``sys.exit()`` (:64-67).  Order: first appearance of each id
    (the reference iterates a py2 dict and writes rows in completion order, i.e. in no defined order)."""
    opener = gzip.open if path.endswith('.gz') else open
    seqs = {}
    try:
        with opener(path, 'rt', newline='\n') as f:          # py2 text mode on POSIX: lines end at '\n' only, nothing is translated
            seqId = None
            for line in f:
                if line[0] == '>':
                    seqId = line[1:].partition(' ')[0].rstrip()
                    seqs[seqId] = []
                else:
                    seqs[seqId].append(line[0:-1])
    except Exception:
        logging.getLogger().error('  [Error] Failed to process sequence file: ' + path)
        sys.exit()
    return [(seqId, ''.join(parts)) for seqId, parts in seqs.items()]


def _format_value(x):
    """py2 ``str(numpy.float64)`` (what genomicSignatures.py:119 ``map(str, sig)`` wrote): 12 significant
    digits, always with a decimal point or exponent."""
    s = '%.12g' % x
    if 'e' not in s and '.' not in s and 'n' not in s and 'i' not in s:
        s += '.0'
    return s


class GenomicSignatures(object):
    def __init__(self, K, threads):
        self.logger = logging.getLogger()
        if K != 4:
            raise NotImplementedError('dbb_b200 implements tetranucleotide (K = 4) signatures only; '
                                      'the reference only ever passes DefaultValues.DEFAULT_KMER_SIZE = 4')
        self.K = K
        self.compl = str.maketrans('ACGT', 'TGCA')
        self.kmerCols, self.kmerToCanonicalIndex = self.__makeKmerColNames()
        self.totalThreads = threads

    def __makeKmerColNames(self):
        """Column list and k-mer -> column dict (genomicSignatures.py:44-72) from the library's table
        (csrc/table.cu), so host and device can never disagree on the column order."""
        names, lut = _lib.kmer_table()
        index = {}
        for code in range(256):
            kmer = ''.join('ACGT'[(code >> (2 * (3 - p))) & 3] for p in range(4))
            index[kmer] = int(lut[code])
        return list(names), index

    def canonicalKmerOrder(self):
        return self.kmerCols

    def numKmers(self):
        return len(self.kmerCols)

    # ---- array-native API ----------------------------------------------------------------------
    def signatures(self, seqs, want_counts=True, want_freq=True):
        """seqs: a list of str/bytes/uint8 arrays, or a tuple (uint8 concat, int64 offsets[N+1]).
        Returns (counts uint32[N,136] or None, freq float32[N,136] or None); freq rows of sequences
        without any valid window are NaN (genomicSignatures.py:147-148)."""
        if isinstance(seqs, tuple):
            flat, off = seqs
            flat = np.ascontiguousarray(flat, dtype=np.uint8)
            off = np.ascontiguousarray(off, dtype=np.int64)
        else:
            # a list of separate buffers (how the reference holds its sequences): one pointer + length per sequence, no
            # concatenation -- the library's packer threads read every sequence where it lies
            n = len(seqs)
            keep = []
            ptrs = np.zeros(max(n, 1), dtype=np.uint64)
            lens = np.zeros(max(n, 1), dtype=np.int64)
            for i, s in enumerate(seqs):
                if isinstance(s, str):
                    s = s.encode('latin-1', 'replace')       # non-latin-1 text cannot be ACGT anyway
                if not isinstance(s, np.ndarray):
                    a = np.frombuffer(s, dtype=np.uint8)
                elif s.dtype == np.uint8 and s.ndim == 1 and s.flags.c_contiguous:
                    a = s
                else:
                    a = np.ascontiguousarray(s, dtype=np.uint8).reshape(-1)
                keep.append(a)
                ptrs[i] = a.__array_interface__['data'][0]
                lens[i] = a.shape[0]
            counts = np.zeros((n, _lib.NUM_KMERS), dtype=np.uint32) if want_counts else None
            freq = np.zeros((n, _lib.NUM_KMERS), dtype=np.float32) if want_freq else None
            if n:
                _lib.check(_lib.lib().dbb_tetra_signatures_host_ptrs(_lib.ptr(ptrs), _lib.ptr(lens), n, _lib.ptr(counts), _lib.ptr(freq)))
            del keep
            return counts, freq
        n = off.shape[0] - 1
        counts = np.zeros((n, _lib.NUM_KMERS), dtype=np.uint32) if want_counts else None
        freq = np.zeros((n, _lib.NUM_KMERS), dtype=np.float32) if want_freq else None
        if n:
            if flat.shape[0] == 0:
                flat = np.zeros(1, dtype=np.uint8)
            _lib.check(_lib.lib().dbb_tetra_signatures_host(_lib.ptr(flat), _lib.ptr(off), n, _lib.ptr(counts), _lib.ptr(freq)))
        return counts, freq

    def seqCounts(self, seq):
        """The integer count vector the reference computes at genomicSignatures.py:135-144 but never
        returns (uint32[136], bit-exact)."""
        return self.signatures([seq], want_freq=False)[0][0]

    # ---- reference API ---------------------------------------------------------------------------
    def seqSignature(self, seq):
        """genomicSignatures.py:134-150.  Counts come from the GPU; the final ``counts / counts.sum()`` of
        these 136 integers is done in float64 exactly as the reference does, so the returned float64
        vector is bit-identical to the reference's (NaN when no window is valid)."""
        sig = self.seqCounts(seq).astype(float)
        with np.errstate(invalid='ignore', divide='ignore'):
            sig /= np.sum(sig)
        return sig

    def calculate(self, seqFile, outputFile):
        """genomicSignatures.py:152-181: signatures of every sequence of a FASTA file -> TSV (header
        ``Sequence Id`` + 136 k-mers, :101-105; one row per sequence, :118-120).  The reference writes rows
        in worker completion order; here they are in file order."""
        self.logger.info('  Determining tetranucleotide signatures of each sequence.')
        seqs = _read_fasta(seqFile)
        counts, _ = self.signatures([s for _, s in seqs], want_freq=False)
        with open(outputFile, 'w') as fout:
            fout.write('Sequence Id')
            for kmer in self.canonicalKmerOrder():
                fout.write('\t' + kmer)
            fout.write('\n')
            for (seqId, _), c in zip(seqs, counts):
                sig = c.astype(float)
                with np.errstate(invalid='ignore', divide='ignore'):
                    sig /= np.sum(sig)
                fout.write(seqId)
                fout.write('\t' + '\t'.join(map(_format_value, sig)))
                fout.write('\n')

    def distance(self, sig1, sig2):
        """genomicSignatures.py:183-184 -- scalar convenience on two host vectors, unchanged.  All-pairs
        distances never go through here: they are computed on the GPU by tdNeighbours.TdNeighbours."""
        return np.sum(np.abs(sig1 - sig2))

    def read(self, tetraProfileFile):
        """genomicSignatures.py:186-194."""
        sig = {}
        with open(tetraProfileFile) as f:
            next(f)
            for line in f:
                lineSplit = line.split('\t')
                sig[lineSplit[0]] = np.array([float(x) for x in lineSplit[1:]])
        return sig
