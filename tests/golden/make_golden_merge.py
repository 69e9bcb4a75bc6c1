#!/usr/bin/env python
"""Golden vectors for the N5 row: the reference's own ``Merge.run`` (merge.py:35-158), driven through real files.

merge.py imports greedy.py and seqUtils.py, which are Python-2-only syntax, so -- as make_golden_greedy.py does -- the
SOURCE TEXT of ``class Merge`` is cut out of merge.py and executed inside a namespace that holds the reference's own
``GenomicSignatures`` / ``Distributions`` / ``checkFileExists`` / ``DefaultValues`` modules and the source text of
``Contig``, ``Core``, ``readContigIdtoBinId`` (greedy.py) and ``readSeqStats`` (seqUtils.py).  Two mechanical Python 3
edits are applied to the cut text and nothing else: ``.iteritems()`` -> ``.items()`` and ``xrange`` -> ``range`` (the
dicts involved are plain ``{}`` literals inside the method, so no subclass can supply ``iteritems``).  Python 3 dicts
iterate in insertion order: bins are visited in order of first appearance in contigs.seq_stats.tsv, which is the
order the mirror freezes (CPython 2.7's order for string keys is a hash accident).  ``readDistributions`` is replaced
by a function returning the tables, because upstream's ``dbb/data/gc_dist.txt`` does not exist.
Build container only; the output ``ref_merge.npz`` is committed."""
import ast
import logging
import os
import string
import sys
import tempfile
import types
from collections import defaultdict

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
from common import checkFileExists                   # noqa: E402
from defaultValues import DefaultValues              # noqa: E402
from make_golden_greedy import KeyListDict, wrap, cut  # noqa: E402,F401


def read(name):
    return open(os.path.join(REF, name)).read().replace('\r\n', '\n').replace('\r', '\n').split('\n')


def py3(lines):
    return '\n'.join(lines).replace('.iteritems()', '.items()').replace('xrange(', 'range(')


def build_namespace(tables):
    ns = {'os': os, 'logging': logging, 'defaultdict': defaultdict, 'checkFileExists': checkFileExists,
          'DefaultValues': DefaultValues, 'GenomicSignatures': GenomicSignatures, 'Distributions': Distributions,
          'readDistributions': lambda preprocessDir, a, b, c: tables}
    exec('class Greedy(object):\n    UNBINNED = -1\n', ns)
    g = read('greedy.py')
    exec(py3(cut(g, 'class Contig(object)', 0)), ns)
    exec(py3(cut(g, 'class Core(object)', 0)), ns)
    exec(py3(cut(g, 'def readContigIdtoBinId(', 0)), ns)
    exec(py3(cut(read('seqUtils.py'), 'def readSeqStats(', 0)), ns)
    exec(py3(cut(read('merge.py'), 'class Merge(object)', 0)), ns)
    return ns


def main():
    from dbb_b200 import synthetic
    logging.getLogger().setLevel(logging.WARNING)
    gs = GenomicSignatures(4, 1)
    tdDist = ast.literal_eval(open(os.path.join(REF, 'data', 'td_dist.txt')).read())
    gcDist = synthetic.synthetic_gc_dist(sorted(tdDist.keys()))
    covDist = synthetic.synthetic_cov_dist(seed=7)
    ns = build_namespace((wrap(gcDist), wrap(tdDist), wrap(covDist)))
    out = {}
    for tag, seed, n, genomes, pers, split in (('a', 41, 70, 4, (99, 99, 99), 1), ('b', 42, 110, 6, (95, 90, 99), 1),
                                                ('c', 43, 120, 3, (99, 99, 99), 3)):
        meta = synthetic.make_metagenome(seed=seed, n_scaffolds=n, min_len=1500, max_len=30000, n_genomes=genomes)
        seqs = synthetic.generate_sequences(meta)
        rng = np.random.RandomState(seed)
        # bins: the genome of origin, with 15 % of the contigs moved to a random other bin (so that "closest bin" is
        # not trivially empty) and 10 % unbinned; case c cuts every genome into `split` bins, so that most contigs
        # have valid bins other than their own and the three "closest" spaces disagree
        bin_of = meta.genome.astype(np.int64) * split + rng.randint(0, split, size=n)
        moved = rng.rand(n) < 0.15
        bin_of[moved] = rng.randint(0, genomes * split, size=int(moved.sum()))
        bin_of[rng.rand(n) < 0.10] = -1
        sigs = [gs.seqSignature(bytes(s).decode('latin-1')) for s in seqs]
        d = tempfile.mkdtemp()
        with open(os.path.join(d, 'contigs.seq_stats.tsv'), 'w') as f:
            f.write('Contig Id\tLength\tGC\tCoverage\n')
            for i in range(n):
                f.write('c%d\t%d\t%r\t%r\n' % (i, meta.length[i], float(meta.gc_obs[i]), float(meta.coverage[i])))
        with open(os.path.join(d, 'contigs.tetra.tsv'), 'w') as f:
            f.write('Contig Id\t' + '\t'.join(gs.canonicalKmerOrder()) + '\n')
            for i in range(n):
                f.write('c%d\t' % i + '\t'.join(repr(float(x)) for x in sigs[i]) + '\n')
        binning = os.path.join(d, 'bins.tsv')
        with open(binning, 'w') as f:
            f.write('# parameters\nContig Id\tBin Id\tLength\n')
            for i in range(n):
                f.write('c%d\t%s\t%d\n' % (i, 'unbinned' if bin_of[i] < 0 else 'bin%d' % bin_of[i], meta.length[i]))
        result = os.path.join(d, 'merge.tsv')
        ns['Merge']().run(d, binning, pers[0], pers[1], pers[2], result)
        text = open(result).read()
        out[tag + '_length'] = np.array(meta.length, dtype=np.int64)
        out[tag + '_gc'] = np.array(meta.gc_obs, dtype=np.float64)
        out[tag + '_cov'] = np.array(meta.coverage, dtype=np.float64)
        out[tag + '_sig'] = np.array(sigs, dtype=np.float64)
        out[tag + '_bin'] = np.array(['unbinned' if b < 0 else 'bin%d' % b for b in bin_of])
        out[tag + '_pers'] = np.array(pers)
        out[tag + '_merge_tsv'] = np.array(text)
        rows = text.strip().split('\n')[1:]
        nz = sum(1 for r in rows if float(r.split('\t')[-1]) > 0)
        print(tag, 'contigs', n, 'rows', len(rows), 'rows with a positive merging score', nz)
    np.savez_compressed(os.path.join(HERE, 'ref_merge.npz'), **out)


if __name__ == '__main__':
    main()
