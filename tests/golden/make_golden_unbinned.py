#!/usr/bin/env python
"""Golden vectors for the N6 row: the reference's own ``Unbinned.run`` (unbinned.py:35-190), driven through real files,
exactly as make_golden_merge.py drives ``Merge.run``: the source text of ``class Unbinned`` is executed with the two
mechanical Python 3 edits (``.iteritems()`` -> ``.items()``, ``xrange`` -> ``range``) in a namespace holding the
reference's own modules and the source text of ``Contig`` / ``Core`` / ``readContigIdtoBinId`` / ``readSeqStats``.
Scaffolds and bins are visited in order of first appearance in contigs.seq_stats.tsv (insertion order), which the
mirror freezes.  Contig ids carry the ``_c<k>`` suffix preprocess gives the pieces of a split scaffold, so scaffolds
with several contigs exist.  Build container only; the output ``ref_unbinned.npz`` is committed."""
import ast
import logging
import os
import re
import sys
import tempfile
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_merge as M                        # noqa: E402  (sets up the shims and sys.path)
from make_golden_merge import REF, read, py3, cut, wrap, GenomicSignatures, Distributions, checkFileExists, DefaultValues  # noqa: E402


def build_namespace(tables):
    ns = {'os': os, 'sys': sys, 're': re, 'logging': logging, 'defaultdict': defaultdict, 'checkFileExists': checkFileExists,
          'DefaultValues': DefaultValues, 'GenomicSignatures': GenomicSignatures, 'Distributions': Distributions,
          'readDistributions': lambda preprocessDir, a, b, c: tables}
    exec('class Greedy(object):\n    UNBINNED = -1\n', ns)
    g = read('greedy.py')
    exec(py3(cut(g, 'class Contig(object)', 0)), ns)
    exec(py3(cut(g, 'class Core(object)', 0)), ns)
    exec(py3(cut(g, 'def readContigIdtoBinId(', 0)), ns)
    exec(py3(cut(read('seqUtils.py'), 'def readSeqStats(', 0)), ns)
    exec(py3(cut(read('unbinned.py'), 'class Unbinned(object)', 0)), ns)
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
    cases = (('a', 61, 90, 3, (99, 99, 99), 2, 0.35, 3000, 1200, False),
             ('b', 62, 130, 4, (95, 99, 99), 3, 0.30, 0, 0, True))
    for tag, seed, n, genomes, pers, split, frac_unbinned, minScaffoldLen, minContigLen, allScaffolds in cases:
        meta = synthetic.make_metagenome(seed=seed, n_scaffolds=n, min_len=1000, max_len=30000, n_genomes=genomes)
        seqs = synthetic.generate_sequences(meta)
        rng = np.random.RandomState(seed)
        # contigs come in runs of 1-4 per scaffold (same genome: a scaffold is one piece of one genome)
        order = np.argsort(meta.genome, kind='stable')
        ids, bin_of = [None] * n, np.full(n, -1, dtype=np.int64)
        pos, sc = 0, 0
        while pos < n:
            run = int(rng.randint(1, 5))
            members = [i for i in order[pos:pos + run] if meta.genome[i] == meta.genome[order[pos]]]
            unb = rng.rand() < frac_unbinned
            b = int(meta.genome[order[pos]]) * split + int(rng.randint(0, split))
            for k, i in enumerate(members):
                ids[i] = 'scaffold%d_c%d' % (sc, k) if len(members) > 1 else 'scaffold%d' % sc
                bin_of[i] = -1 if unb else b
            pos += len(members)
            sc += 1
        sigs = [gs.seqSignature(bytes(s).decode('latin-1')) for s in seqs]
        d = tempfile.mkdtemp()
        with open(os.path.join(d, 'contigs.seq_stats.tsv'), 'w') as f:
            f.write('Contig Id\tLength\tGC\tCoverage\n')
            for i in range(n):
                f.write('%s\t%d\t%r\t%r\n' % (ids[i], meta.length[i], float(meta.gc_obs[i]), float(meta.coverage[i])))
        with open(os.path.join(d, 'contigs.tetra.tsv'), 'w') as f:
            f.write('Contig Id\t' + '\t'.join(gs.canonicalKmerOrder()) + '\n')
            for i in range(n):
                f.write(ids[i] + '\t' + '\t'.join(repr(float(x)) for x in sigs[i]) + '\n')
        binning = os.path.join(d, 'bins.tsv')
        with open(binning, 'w') as f:
            f.write('# parameters\nContig Id\tBin Id\tLength\n')
            for i in range(n):
                f.write('%s\t%s\t%d\n' % (ids[i], 'unbinned' if bin_of[i] < 0 else 'bin%d' % bin_of[i], meta.length[i]))
        result = os.path.join(d, 'unbinned.tsv')
        ns['Unbinned']().run(d, binning, pers[0], pers[1], pers[2], minScaffoldLen, minContigLen, allScaffolds, result)
        text = open(result).read()
        out[tag + '_id'] = np.array(ids)
        out[tag + '_length'] = np.array(meta.length, dtype=np.int64)
        out[tag + '_gc'] = np.array(meta.gc_obs, dtype=np.float64)
        out[tag + '_cov'] = np.array(meta.coverage, dtype=np.float64)
        out[tag + '_sig'] = np.array(sigs, dtype=np.float64)
        out[tag + '_bin'] = np.array(['unbinned' if b < 0 else 'bin%d' % b for b in bin_of])
        out[tag + '_params'] = np.array(list(pers) + [minScaffoldLen, minContigLen, int(allScaffolds)])
        out[tag + '_unbinned_tsv'] = np.array(text)
        rows = [r.split('\t') for r in text.strip().split('\n')[1:]]
        print(tag, 'contigs', n, 'rows', len(rows), 'rows with a positive score', sum(float(r[-1]) > 0 for r in rows),
              'multi-contig scaffolds', len(set(r[0] for r in rows if int(r[7]) > 1)))
    np.savez_compressed(os.path.join(HERE, 'ref_unbinned.npz'), **out)


if __name__ == '__main__':
    main()
