#!/usr/bin/env python
"""Golden vectors for the N3 row: the reference's own ``Greedy.__greedy`` (greedy.py:167-251).  greedy.py as a whole is
Python-2-only syntax (print statements), so the SOURCE TEXT of ``class Core`` (:147-157) and of the ``__greedy`` method
is cut out of the file and executed unmodified inside a two-line ``class Greedy`` shell, with ``xrange = range`` in its
namespace; ``Distributions`` and ``GenomicSignatures`` are the reference's own modules (same interpreter shims as
make_golden.py).  Build container only; the output ``ref_greedy.npz`` is committed."""
import ast
import logging
import os
import string
import sys
import types

import numpy as np

REF = '/root/reference/dbb'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

string.maketrans = str.maketrans
_stub = types.ModuleType('seqUtils')
_stub.readFasta = lambda path: {}
sys.modules['seqUtils'] = _stub
sys.path.insert(0, REF)
from genomicSignatures import GenomicSignatures      # noqa: E402
from distributions import Distributions              # noqa: E402


class KeyListDict(dict):
    def keys(self):
        return sorted(dict.keys(self))


def wrap(d):
    if isinstance(d, dict):
        return KeyListDict((k, wrap(v)) for k, v in d.items())
    return d


def cut(lines, start_pat, indent):
    """source lines of the block that starts at the first line containing start_pat, up to the next line with the
    same or smaller indentation that starts a new def/class"""
    i = next(k for k, l in enumerate(lines) if start_pat in l)
    j = i + 1
    while j < len(lines):
        l = lines[j]
        if l.strip() and (len(l) - len(l.lstrip())) <= indent and (l.lstrip().startswith('def ') or l.lstrip().startswith('class ')):
            break
        j += 1
    return lines[i:j]


src = open(os.path.join(REF, 'greedy.py')).read().replace('\r\n', '\n').replace('\r', '\n').split('\n')
ns = {'xrange': range, 'logging': logging, 'sys': sys, 'Distributions': Distributions}
exec('\n'.join(cut(src, 'class Core(object)', 0)), ns)
exec('class Greedy(object):\n    UNBINNED = -1\n    def __init__(self):\n        self.logger = logging.getLogger()\n\n' +
     '\n'.join(cut(src, 'def __greedy(', 4)), ns)
Greedy, Core = ns['Greedy'], ns['Core']


class Contig(object):                                  # greedy.py:133-141
    def __init__(self, contigId, length, GC, coverage, tetraSig):
        self.binId = -1
        self.contigId, self.length, self.GC, self.coverage, self.tetraSig = contigId, length, GC, coverage, tetraSig


def main():
    from dbb_b200 import synthetic
    logging.getLogger().setLevel(logging.WARNING)
    gs = GenomicSignatures(4, 1)
    tdDist = ast.literal_eval(open(os.path.join(REF, 'data', 'td_dist.txt')).read())
    gcDist = synthetic.synthetic_gc_dist(sorted(tdDist.keys()))
    covDist = synthetic.synthetic_cov_dist(seed=7)
    out = {}
    for tag, seed, n, genomes, per, minBin in (('a', 31, 90, 3, 95, 40000), ('b', 32, 120, 5, 99, 15000)):
        meta = synthetic.make_metagenome(seed=seed, n_scaffolds=n, min_len=1500, max_len=25000, n_genomes=genomes)
        seqs = synthetic.generate_sequences(meta)
        order = np.argsort(-meta.length, kind='stable')            # greedy.py:420 intends longest first
        contigs = []
        for i in order:
            s = bytes(seqs[i]).decode('latin-1')
            contigs.append(Contig('c%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]), gs.seqSignature(s)))
        sig_in = np.array([c.tetraSig.copy() for c in contigs])    # __greedy overwrites the seed contigs' arrays in place
        cores = getattr(Greedy(), '_Greedy__greedy')(contigs, minBin, per, wrap(gcDist), wrap(tdDist), wrap(covDist), gs, 1)
        out[tag + '_length'] = np.array([c.length for c in contigs])
        out[tag + '_gc'] = np.array([c.GC for c in contigs])
        out[tag + '_cov'] = np.array([c.coverage for c in contigs])
        out[tag + '_sig'] = sig_in
        out[tag + '_params'] = np.array([per, minBin])
        out[tag + '_binId'] = np.array([c.binId for c in contigs])
        out[tag + '_core_len'] = np.array([c.length for c in cores])
        out[tag + '_core_gc'] = np.array([c.GC for c in cores])
        out[tag + '_core_cov'] = np.array([c.coverage for c in cores])
        out[tag + '_core_sig'] = np.array([np.array(c.tetraSig) for c in cores]).reshape(len(cores), 136)
        out[tag + '_core_sizes'] = np.array([len(c.contigs) for c in cores])
        print(tag, 'contigs', n, 'cores', len(cores), 'sizes', [len(c.contigs) for c in cores], 'unbinned', int((out[tag + '_binId'] < 0).sum()))
    np.savez_compressed(os.path.join(HERE, 'ref_greedy.npz'), **out)


if __name__ == '__main__':
    main()
