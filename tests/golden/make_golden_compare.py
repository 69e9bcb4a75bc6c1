#!/usr/bin/env python
"""Golden vectors for compare.py: the reference's own ``Compare.run`` (compare.py:39-152), driven through the input files
of the N6 golden cases (tests/golden/ref_unbinned.npz), its standard output captured.  Same method as
make_golden_merge.py / make_golden_unbinned.py; one more mechanical Python 3 edit is needed because compare.py reports
with print STATEMENTS: every logical line ``print X`` becomes ``print(X)`` (X up to the closing parenthesis of the
statement), nothing else.  Build container only; the output ``ref_compare.json`` is committed."""
import contextlib
import io
import json
import logging
import os
import re
import sys
import tempfile
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_merge as M                        # noqa: E402,F401  (sets up the shims and sys.path)
from make_golden_merge import REF, read, py3, cut, wrap, GenomicSignatures, Distributions, checkFileExists, DefaultValues  # noqa: E402


def print_statements_to_calls(text):
    out, lines, i = [], text.split('\n'), 0
    while i < len(lines):
        m = re.match(r'^(\s*)print (.*)$', lines[i])
        if not m:
            out.append(lines[i])
            i += 1
            continue
        stmt = m.group(2)
        while stmt.count('(') > stmt.count(')'):
            i += 1
            stmt += '\n' + lines[i]
        out.append(m.group(1) + 'print(' + stmt + ')')
        i += 1
    return '\n'.join(out)


def build_namespace(tables):
    ns = {'os': os, 'sys': sys, 're': re, 'logging': logging, 'defaultdict': defaultdict, 'checkFileExists': checkFileExists,
          'DefaultValues': DefaultValues, 'GenomicSignatures': GenomicSignatures, 'Distributions': Distributions,
          'readDistributions': lambda preprocessDir, a, b, c: tables, 'mean': np.mean}
    exec('class Greedy(object):\n    UNBINNED = -1\n', ns)
    g = read('greedy.py')
    exec(py3(cut(g, 'class Contig(object)', 0)), ns)
    exec(py3(cut(g, 'class Core(object)', 0)), ns)
    exec(py3(cut(g, 'def readContigIdtoBinId(', 0)), ns)
    exec(py3(cut(read('seqUtils.py'), 'def readSeqStats(', 0)), ns)
    exec(print_statements_to_calls(py3(cut(read('compare.py'), 'class Compare(object)', 0))), ns)
    return ns


def main():
    import ast
    from dbb_b200 import synthetic
    logging.getLogger().setLevel(logging.WARNING)
    gs = GenomicSignatures(4, 1)
    tdDist = ast.literal_eval(open(os.path.join(REF, 'data', 'td_dist.txt')).read())
    gcDist = synthetic.synthetic_gc_dist(sorted(tdDist.keys()))
    covDist = synthetic.synthetic_cov_dist(seed=7)
    ns = build_namespace((wrap(gcDist), wrap(tdDist), wrap(covDist)))
    g = np.load(os.path.join(HERE, 'ref_unbinned.npz'))
    out = []
    for tag in 'ab':
        ids, bins = [str(x) for x in g[tag + '_id']], [str(x) for x in g[tag + '_bin']]
        n = len(ids)
        d = tempfile.mkdtemp()
        with open(os.path.join(d, 'contigs.seq_stats.tsv'), 'w') as f:
            f.write('Contig Id\tLength\tGC\tCoverage\n')
            for i in range(n):
                f.write('%s\t%d\t%r\t%r\n' % (ids[i], g[tag + '_length'][i], float(g[tag + '_gc'][i]), float(g[tag + '_cov'][i])))
        with open(os.path.join(d, 'contigs.tetra.tsv'), 'w') as f:
            f.write('Contig Id\t' + '\t'.join(gs.canonicalKmerOrder()) + '\n')
            for i in range(n):
                f.write(ids[i] + '\t' + '\t'.join(repr(float(x)) for x in g[tag + '_sig'][i]) + '\n')
        binning = os.path.join(d, 'bins.tsv')
        with open(binning, 'w') as f:
            f.write('# parameters\nContig Id\tBin Id\tLength\n')
            for i in range(n):
                f.write('%s\t%s\t%d\n' % (ids[i], bins[i], g[tag + '_length'][i]))
        # query scaffolds: multi-contig ones, unbinned and binned; query bins: the first bins met
        scaffolds = []
        for x in ids:
            m = re.search('_c[0-9]*$', x)
            s = x[:m.start()] if m else x
            if m and s not in scaffolds:
                scaffolds.append(s)
        binIds = []
        for b in bins:
            if b != 'unbinned' and b not in binIds:
                binIds.append(b)
        pers = [int(x) for x in g[tag + '_params'][:3]]
        for q, scaffoldId in enumerate(scaffolds[:6]):
            binId = binIds[q % len(binIds)]
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                ns['Compare']().run(d, binning, binId, scaffoldId, pers[0], pers[1], pers[2])
            out.append({'case': tag, 'bin': binId, 'scaffold': scaffoldId, 'pers': pers, 'stdout': buf.getvalue()})
        print(tag, 'queries', len(scaffolds[:6]), 'lines', sum(o['stdout'].count('\n') for o in out if o['case'] == tag))
    with open(os.path.join(HERE, 'ref_compare.json'), 'w') as f:
        json.dump(out, f, indent=0)
    print(out[0]['stdout'])


if __name__ == '__main__':
    main()
