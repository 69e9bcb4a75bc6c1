#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE'S OWN SOURCE FILES (unmodified, read from
/root/reference/dbb) under Python 3.  Run in the build container only; the outputs
(tests/golden/ref_*.json / .npz) are committed and are what the tests read.

The reference is Python 2.  Nothing in its hot-path arithmetic is py2-specific, only three API
names are, so the interpreter environment is shimmed -- the reference code itself is not edited:

  1. ``string.maketrans``            -> ``str.maketrans``  (genomicSignatures.py:26,39)
  2. module ``seqUtils``             -> stub (py2-only syntax at seqUtils.py:158; only ``readFasta``
                                        is imported and only ``calculate`` uses it, not exercised)
  3. ``dict.keys()[i]`` / list-typed ``keys()`` (distributions.py:59-73,76-110,137-138)
                                     -> the nested distribution dicts are passed as ``KeyListDict``
                                        (a dict subclass whose ``keys()`` returns an ASCENDING list).
     CPython-2.7 would return hash order; that order only matters at exact midpoints between two
     keys (``__findNearest`` takes the first arg-min).  Ascending order is the frozen rule
     (DESIGN.md), and the generator avoids midpoint lengths except in one labelled case.

The neighbour ``__workerThread`` methods (tdNeighbours.py:33-54, gcNeighbours.py:34-55,
coverageNeighbours.py:33-53) are driven in-process through list-backed queue objects, so the
reference's inner loops run verbatim without multiprocessing.
"""
import json
import os
import string
import sys
import types

import numpy as np

REF = '/root/reference/dbb'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

# ---- shims ----------------------------------------------------------------------------------
string.maketrans = str.maketrans
_stub = types.ModuleType('seqUtils')
_stub.readFasta = lambda path: {}
_stub.readSeqStats = lambda path: {}
sys.modules['seqUtils'] = _stub
_common = types.ModuleType('common')
_common.checkFileExists = lambda path: None
sys.modules.setdefault('common', _common)
# greedy.py is py2-only syntax (print statements); refineBins.py only needs these names from it
_greedy = types.ModuleType('greedy')


class _Greedy(object):
    UNBINNED = -1


class _Core(object):                                   # greedy.py:147-157
    def __init__(self, binId, length, GC, coverage, tetraSig):
        self.binId, self.length, self.GC, self.coverage, self.tetraSig, self.contigs = binId, length, GC, coverage, tetraSig, None


_greedy.Greedy, _greedy.Core = _Greedy, _Core
_greedy.Contig = _greedy.reportBins = _greedy.readContigIdtoBinId = None
sys.modules['greedy'] = _greedy
sys.path.insert(0, REF)

from genomicSignatures import GenomicSignatures      # noqa: E402  (reference source)
from tdNeighbours import TdNeighbours                # noqa: E402
from gcNeighbours import GcNeighbours                # noqa: E402
from coverageNeighbours import CoverageNeighbours    # noqa: E402
from distributions import Distributions              # noqa: E402
from refineBins import RefineBins                    # noqa: E402


class KeyListDict(dict):
    def keys(self):
        return sorted(dict.keys(self))


def wrap(d):
    if isinstance(d, dict):
        return KeyListDict((k, wrap(v)) for k, v in d.items())
    return d


class ListQueue(object):
    def __init__(self, items=()):
        self.items = list(items)

    def get(self, block=True, timeout=None):
        return self.items.pop(0)

    def put(self, x):
        self.items.append(x)


class Contig(object):                                  # greedy.py:133-141
    def __init__(self, contigId, length, GC, coverage, tetraSig):
        self.contigId, self.length, self.GC, self.coverage, self.tetraSig = contigId, length, GC, coverage, tetraSig


def run_rows(obj, mangled, seqs, distributions):
    qin = ListQueue([(i, s) for i, s in enumerate(seqs)] + [(None, None)])
    qout = ListQueue()
    getattr(obj, mangled)(seqs, distributions, qin, qout)
    rows = [None] * len(seqs)
    for i, row in qout.items:
        rows[i] = row
    return np.array(rows)


def main():
    from dbb_b200 import synthetic                     # shared seeded generator (host side)

    gs = GenomicSignatures(4, 1)

    # ---- 1. column table -------------------------------------------------------------------
    table = {'kmerCols': gs.canonicalKmerOrder(), 'numKmers': gs.numKmers(),
             'kmerToCanonicalIndex': gs.kmerToCanonicalIndex}

    # ---- 2. signatures: hand-written KATs + seeded random sequences ---------------------------
    rng = np.random.RandomState(12345)
    kat = ['', 'A', 'ACG', 'ACGT', 'AAAA', 'TTTT', 'AAAAA', 'ACGNACGT', 'acgtACGT', 'NNNN',
           'ACGTACGTACGT', 'GATTACA', 'ACGTN', 'NACGT', 'ACGUACGT', 'ACGTRYACGTT']
    for L in (4, 5, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65, 66, 67, 68, 127, 128, 129, 130, 131,
              255, 256, 257, 511, 512, 513, 1000, 1023, 1024, 1025, 4093, 4099):
        kat.append(''.join(rng.choice(list('ACGT'), L)))
    for L in (40, 200, 777, 2049):                    # with N, IUPAC and lowercase
        kat.append(''.join(rng.choice(list('ACGTNacgtRYn'), L, p=[.22, .22, .22, .22, .03, .01, .01, .01, .01, .02, .02, .01])))
    sigs = []
    with np.errstate(invalid='ignore', divide='ignore'):
        for s in kat:
            sigs.append(gs.seqSignature(s))
    sigs = np.array(sigs)

    # ---- 3. neighbour rows on a small seeded contig set --------------------------------------
    import ast
    tdDist = ast.literal_eval(open(os.path.join(REF, 'data', 'td_dist.txt')).read())
    meta = synthetic.make_metagenome(seed=7, n_scaffolds=48, min_len=600, max_len=30000, n_genomes=4)
    seqs_ascii = synthetic.generate_sequences(meta)
    gcDist = synthetic.synthetic_gc_dist(sorted(tdDist.keys()))
    covDist = synthetic.synthetic_cov_dist(seed=7)
    contigs = []
    with np.errstate(invalid='ignore', divide='ignore'):
        for n, s in enumerate(seqs_ascii):
            s = s.tobytes().decode('ascii')
            contigs.append(Contig('c%d' % n, int(meta.length[n]), float(meta.gc_obs[n]),
                                  float(meta.coverage[n]), gs.seqSignature(s)))
    # equal lengths (asymmetry case), a zero-coverage core, a NaN signature, a midpoint length
    contigs[5].length = contigs[6].length
    contigs[7].coverage = 0.0
    contigs[8].tetraSig = np.full(136, np.nan)
    contigs[9].length = 12500                           # exact midpoint of 10000/15000 (labelled)

    neigh = {}
    for tag, fixed, per in (('fixedNone_p80', None, 80), ('fixed10000_p80', 10000, 80), ('fixedNone_p99', None, 99)):
        dT = Distributions(None, None, wrap(tdDist), per, None, None)
        dG = Distributions(wrap(gcDist), per, None, None, None, None)
        dC = Distributions(None, None, None, None, wrap(covDist), per)
        neigh['td_' + tag] = run_rows(TdNeighbours(fixed), '_TdNeighbours__workerThread', contigs, dT)
        neigh['gc_' + tag] = run_rows(GcNeighbours(fixed), '_GcNeighbours__workerThread', contigs, dG)
        neigh['cov_' + tag] = run_rows(CoverageNeighbours(fixed), '_CoverageNeighbours__workerThread', contigs, dC)
    dist = np.array([[gs.distance(a.tetraSig, b.tetraSig) for b in contigs] for a in contigs])
    tetra_snapshot = np.array([c.tetraSig for c in contigs])       # (row 8 is NaN here; part 4 restores it)

    # ---- 4. refineBins.py:42-96 (__workerThread) on the same contigs: 30 binned into 4 cores, 18 unbinned -------
    class IterDict(KeyListDict):
        def iteritems(self):
            return iter(sorted(self.items()))

    for c in contigs:
        c.binId = -1
    contigs[8].tetraSig = gs.seqSignature(seqs_ascii[8].tobytes().decode('ascii'))     # undo the NaN row for this part
    binned = contigs[:30]
    for n, c in enumerate(binned):
        c.binId = int(meta.genome[n])
    cores = IterDict()
    for c in binned:                                    # refineBins.py:288-304, verbatim arithmetic
        if c.binId not in cores:
            cores[c.binId] = _Core(c.binId, 0, 0, 0, [0] * gs.numKmers())
        core = cores[c.binId]
        core.length += c.length
        weight = float(c.length) / core.length
        core.GC = c.GC * weight + core.GC * (1.0 - weight)
        core.coverage = c.coverage * weight + core.coverage * (1.0 - weight)
        for i in range(0, len(core.tetraSig)):
            core.tetraSig[i] = c.tetraSig[i] * weight + core.tetraSig[i] * (1.0 - weight)
    unbinned = contigs[30:]
    dR = Distributions(wrap(gcDist), 99, wrap(tdDist), 99, wrap(covDist), 99)
    qin = ListQueue(list(range(len(unbinned))) + [None])
    qout = ListQueue()
    getattr(RefineBins(), '_RefineBins__workerThread')(unbinned, cores, dR, gs, qin, qout)
    refine_rows = np.array([[r[0], r[1], r[2], int(r[3]), int(r[4]), int(r[5])] for r in qout.items])
    neigh['refine_rows'] = refine_rows
    neigh['refine_core_ids'] = np.array(sorted(cores.keys()))
    neigh['refine_core_len'] = np.array([cores[k].length for k in sorted(cores.keys())])
    neigh['refine_core_gc'] = np.array([cores[k].GC for k in sorted(cores.keys())])
    neigh['refine_core_cov'] = np.array([cores[k].coverage for k in sorted(cores.keys())])
    neigh['refine_core_sig'] = np.array([cores[k].tetraSig for k in sorted(cores.keys())])
    neigh['refine_bin_of_contig'] = np.array([c.binId for c in contigs])
    neigh['refine_sig8'] = contigs[8].tetraSig

    # the reference's td_dist.txt re-encoded as arrays (exact float64 values) for dbb_b200/data
    lens = sorted(tdDist.keys())
    pcts = sorted(tdDist[lens[0]].keys())
    np.savez_compressed(os.path.join(HERE, '..', '..', 'dbb_b200', 'data', 'td_dist.npz'),
                        lengths=np.array(lens, dtype=np.int64), percentiles=np.array(pcts, dtype=np.float64),
                        values=np.array([[tdDist[l][p] for p in pcts] for l in lens], dtype=np.float64))

    with open(os.path.join(HERE, 'ref_table.json'), 'w') as f:
        json.dump(table, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, 'ref_kat_seqs.json'), 'w') as f:
        json.dump(kat, f, indent=0)
    np.savez_compressed(os.path.join(HERE, 'ref_signatures.npz'), sigs=sigs)
    np.savez_compressed(
        os.path.join(HERE, 'ref_neighbours.npz'),
        length=np.array([c.length for c in contigs]), GC=np.array([c.GC for c in contigs]),
        coverage=np.array([c.coverage for c in contigs]), tetraSig=tetra_snapshot,
        distance=dist, **neigh)
    print('wrote golden vectors: %d KAT sequences, %d contigs' % (len(kat), len(contigs)))


if __name__ == '__main__':
    main()
