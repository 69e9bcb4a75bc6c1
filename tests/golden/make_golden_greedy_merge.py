#!/usr/bin/env python
"""Golden vectors for the rest of the N3 row: the reference's own ``Greedy.__merge`` (greedy.py:253-334) and the
neighbourhood sizes of ``Greedy.__sortContigs`` (greedy.py:390-422).  As in make_golden_greedy.py the SOURCE TEXT of
``class Core``, ``__greedy`` and ``__merge`` is cut out of greedy.py (Python-2-only as a whole) and executed unmodified
inside a ``class Greedy`` shell with ``xrange = range``.  ``__sortContigs`` constructs the three neighbour classes, whose
``run`` forks workers over py2 queues; its lines 406-417 (the neighbourhood sizes) are executed verbatim on rows produced
by the reference's own ``__workerThread`` loops, exactly as make_golden.py drives them.  Build container only; the
output ``ref_greedy_merge.npz`` is committed."""
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
import tdNeighbours, gcNeighbours, coverageNeighbours   # noqa: E402,E401

from make_golden_greedy import KeyListDict, wrap, cut, Contig   # noqa: E402,F401

src = open(os.path.join(REF, 'greedy.py')).read().replace('\r\n', '\n').replace('\r', '\n').split('\n')
ns = {'xrange': range, 'logging': logging, 'sys': sys, 'Distributions': Distributions}
exec('\n'.join(cut(src, 'class Core(object)', 0)), ns)
exec('class Greedy(object):\n    UNBINNED = -1\n    def __init__(self):\n        self.logger = logging.getLogger()\n\n' +
     '\n'.join(cut(src, 'def __greedy(', 4)) + '\n' + '\n'.join(cut(src, 'def __merge(', 4)), ns)
Greedy, Core = ns['Greedy'], ns['Core']


class _Q(object):
    """queue stand-in for the reference's worker loops (same as make_golden.py)"""
    def __init__(self, items=()):
        self.items = list(items)
    def get(self, *a, **k):
        return self.items.pop(0)
    def put(self, x, *a, **k):
        self.items.append(x)


def rows_of(cls, name, contigs, distributions, fixed):
    obj = cls(fixed)
    qin = _Q([(i, c) for i, c in enumerate(contigs)] + [(None, None)])
    qout = _Q()
    getattr(obj, '_%s__workerThread' % name)(contigs, distributions, qin, qout)
    rows = [None] * len(contigs)
    for i, row in qout.items:
        rows[i] = np.asarray(row, dtype=np.float64)
    return rows


def main():
    from dbb_b200 import synthetic
    logging.getLogger().setLevel(logging.WARNING)
    gs = GenomicSignatures(4, 1)
    tdDist = ast.literal_eval(open(os.path.join(REF, 'data', 'td_dist.txt')).read())
    gcDist = synthetic.synthetic_gc_dist(sorted(tdDist.keys()))
    covDist = synthetic.synthetic_cov_dist(seed=7)
    out = {}
    # ---- __merge: cores grown with a tight percentile, merged with a loose one
    for tag, seed, n, genomes, build_per, merge_per, minBin in (('a', 33, 140, 3, 60, 99, 8000), ('b', 34, 160, 4, 50, 98, 5000)):
        meta = synthetic.make_metagenome(seed=seed, n_scaffolds=n, min_len=1500, max_len=25000, n_genomes=genomes)
        seqs = synthetic.generate_sequences(meta)
        order = np.argsort(-meta.length, kind='stable')
        contigs = [Contig('c%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]),
                          gs.seqSignature(bytes(seqs[i]).decode('latin-1'))) for i in order]
        g = Greedy()
        cores = getattr(g, '_Greedy__greedy')(contigs, minBin, build_per, wrap(gcDist), wrap(tdDist), wrap(covDist), gs, 1)
        out[tag + '_in_len'] = np.array([c.length for c in cores])
        out[tag + '_in_gc'] = np.array([c.GC for c in cores])
        out[tag + '_in_cov'] = np.array([c.coverage for c in cores])
        out[tag + '_in_sig'] = np.array([np.array(c.tetraSig, dtype=float) for c in cores]).reshape(len(cores), 136)
        out[tag + '_in_sizes'] = np.array([len(c.contigs) for c in cores])
        first = np.array([c.binId for c in contigs])
        merged = getattr(g, '_Greedy__merge')(cores, merge_per, wrap(gcDist), wrap(tdDist), wrap(covDist), gs, 1)
        out[tag + '_params'] = np.array([build_per, merge_per, minBin])
        out[tag + '_contig_core'] = first                               # core of every contig before the merge (-1: none)
        out[tag + '_contig_bin'] = np.array([c.binId for c in contigs])  # after
        out[tag + '_core_bin'] = np.array([c.binId for c in cores])       # merged bin of every input core
        out[tag + '_out_len'] = np.array([c.length for c in merged])
        out[tag + '_out_gc'] = np.array([c.GC for c in merged])
        out[tag + '_out_cov'] = np.array([c.coverage for c in merged])
        out[tag + '_out_sig'] = np.array([np.array(c.tetraSig, dtype=float) for c in merged]).reshape(len(merged), 136)
        out[tag + '_out_sizes'] = np.array([len(c.contigs) for c in merged])
        print(tag, 'cores', len(cores), '-> merged', len(merged), 'sizes', [len(c.contigs) for c in merged])
    # ---- __sortContigs: neighbourhood sizes (greedy.py:406-417) on rows from the reference's own worker loops
    meta = synthetic.make_metagenome(seed=35, n_scaffolds=60, min_len=10001, max_len=60000, n_genomes=3)
    seqs = synthetic.generate_sequences(meta)
    contigs = [Contig('c%d' % i, int(meta.length[i]), float(meta.gc_obs[i]), float(meta.coverage[i]),
                      gs.seqSignature(bytes(seqs[i]).decode('latin-1'))) for i in range(meta.n_scaffolds)]
    minSeqLen, per = 10000, 80
    dG = Distributions(wrap(gcDist), per, None, None, None, None)           # as gcNeighbours.py:89 etc. build them
    dT = Distributions(None, None, wrap(tdDist), per, None, None)
    dC = Distributions(None, None, None, None, wrap(covDist), per)
    for c, r in zip(contigs, rows_of(gcNeighbours.GcNeighbours, 'GcNeighbours', contigs, dG, minSeqLen)):
        c.gcNeighbours = r
    for c, r in zip(contigs, rows_of(tdNeighbours.TdNeighbours, 'TdNeighbours', contigs, dT, minSeqLen)):
        c.tdNeighbours = r
    for c, r in zip(contigs, rows_of(coverageNeighbours.CoverageNeighbours, 'CoverageNeighbours', contigs, dC, minSeqLen)):
        c.covNeighbours = r
    body = cut(src, 'def __sortContigs(', 4)
    i0 = next(k for k, l in enumerate(body) if 'neighbourhoodSizes = []' in l)
    i1 = next(k for k, l in enumerate(body) if 'sortedIndices = np.argsort' in l)
    code = '\n'.join(l[8:] for l in body[i0:i1 + 1])                   # lines 406-420, dedented
    env = {'np': np, 'contigs': contigs}
    exec(code, env)
    out['s_length'] = np.array([c.length for c in contigs])
    out['s_gc'] = np.array([c.GC for c in contigs])
    out['s_cov'] = np.array([c.coverage for c in contigs])
    out['s_sig'] = np.array([c.tetraSig for c in contigs])
    out['s_params'] = np.array([minSeqLen, per])
    out['s_sizes'] = np.array(env['neighbourhoodSizes'], dtype=np.int64)
    out['s_sorted'] = np.array(env['sortedIndices'], dtype=np.int64)
    print('sortContigs sizes', out['s_sizes'][:8], 'order', out['s_sorted'][:8])
    np.savez_compressed(os.path.join(HERE, 'ref_greedy_merge.npz'), **out)


if __name__ == '__main__':
    main()
