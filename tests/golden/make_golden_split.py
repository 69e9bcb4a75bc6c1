#!/usr/bin/env python
"""Golden vectors for the N2 row (sequence side of preprocess): the reference's own
``SplitScaffolds.splits`` (splitScaffolds.py:27-53, imported unmodified; its ``from seqUtils import readFasta``
is satisfied by a stub because seqUtils.py is Python-2-only syntax) and ``baseCount`` (seqUtils.py:258-265, the
function's source text is cut out of the file and executed unmodified).  Build container only; the output
``ref_split.json`` is committed."""
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference/dbb'
HERE = os.path.dirname(os.path.abspath(__file__))

_stub = types.ModuleType('seqUtils')
_stub.readFasta = lambda path: {}
sys.modules['seqUtils'] = _stub
sys.path.insert(0, REF)
from splitScaffolds import SplitScaffolds  # noqa: E402

src = open(os.path.join(REF, 'seqUtils.py')).read().replace('\r\n', '\n').replace('\r', '\n').split('\n')
start = next(i for i, l in enumerate(src) if l.startswith('def baseCount'))
end = next(i for i in range(start + 1, len(src)) if src[i].startswith('def '))
ns = {}
exec('\n'.join(src[start:end]), ns)
baseCount = ns['baseCount']


def main():
    rng = np.random.RandomState(20240611)
    seqs = ['', 'A', 'N', 'ACGT', 'NNNN', 'ACGTNNNNNACGTNNACGTNNN', 'NNNNNACGT', 'ACGNNNNNT', 'ACGTnnnnNNACGTAC',
            'ACGTNNNNACGTNNNNNACGNNNNNNAC', 'acgtuUNNNNNNNNNNNNRYKMacgt', 'ACGTNNNNNNNNNNA', 'ACNNNNNGTNNNNNNNNNNNNNNNNNNNNN',
            'NNNNNNNNNNNNACGTACGTNNNNNNNNNNNNN', 'ACGTACGTAC' * 7]
    for L in (50, 333, 1000, 4097):
        s = rng.choice(list('ACGT'), L)
        for _ in range(rng.randint(1, 6)):
            p, r = rng.randint(0, L), rng.randint(1, 40)
            s[p:p + r] = rng.choice(['N', 'n'], min(r, L - p))
        for _ in range(rng.randint(0, 4)):
            p, r = rng.randint(0, L), rng.randint(1, 30)
            s[p:p + r] = [c.lower() for c in s[p:p + r]]
        seqs.append(''.join(s))
    ss = SplitScaffolds()
    cases = []
    for i, s in enumerate(seqs):
        for minN in (0, 1, 3, 10, 25):
            contigs = ss.splits('s%d' % i, s, minN)
            cases.append({'seq': i, 'minN': minN, 'contigs': {k: list(map(int, v)) for k, v in contigs.items()}})
    counts = [list(map(int, baseCount(s))) for s in seqs]
    with open(os.path.join(HERE, 'ref_split.json'), 'w') as f:
        json.dump({'seqs': seqs, 'splits': cases, 'baseCount': counts}, f, indent=0, sort_keys=True)
    print('wrote %d sequences, %d split cases' % (len(seqs), len(cases)))


if __name__ == '__main__':
    main()
