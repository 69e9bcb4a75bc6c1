"""N2 row: N-run splitter, base composition and batched sequence statistics against the reference's own
outputs (golden) and the oracle's literal loops."""
import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _random_scaffolds(seed, n, max_len=6000):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        L = int(rng.randint(0, max_len)) if i % 17 else int(rng.choice([0, 1, 2, 3, 4, 63, 64, 65, 2047, 2048, 2049]))
        s = rng.choice(list('ACGT'), L)
        for _ in range(rng.randint(0, 5)):
            if L == 0:
                break
            p, r = rng.randint(0, L), int(rng.choice([1, 2, 3, 4, 5, 10, 11, 60, 64, 70, 130, 3000]))
            s[p:p + r] = rng.choice(['N', 'n'], len(s[p:p + r]))
        for _ in range(rng.randint(0, 3)):
            if L == 0:
                break
            p, r = rng.randint(0, L), rng.randint(1, 50)
            s[p:p + r] = rng.choice(list('acgtuURYK'), len(s[p:p + r]))
        out.append(''.join(s))
    return out


def test_splits_and_base_counts_match_reference_golden():
    import torch
    from dbb_b200 import device
    from dbb_b200.splitScaffolds import SplitScaffolds
    g = util.golden_json('ref_split.json')
    seqs = g['seqs']
    ss = SplitScaffolds()
    by_min = {}
    for case in g['splits']:
        by_min.setdefault(case['minN'], []).append(case)
    for minN, cases in by_min.items():
        got = ss.splitsBatch(['s%d' % c['seq'] for c in cases], [seqs[c['seq']] for c in cases], minN)
        for c, d in zip(cases, got):
            assert d == c['contigs'], (c['seq'], minN)
    assert ss.splits('s5', seqs[5], 3) == O.splits_literal('s5', seqs[5], 3)
    raw = [s.encode() for s in seqs]
    off = np.zeros(len(raw) + 1, np.int64); off[1:] = np.cumsum([len(r) for r in raw])
    d_ascii = torch.from_numpy(np.frombuffer(b''.join(raw) + b'\0\0\0\0', dtype=np.uint8).copy()).cuda()
    acgt = device.base_counts(d_ascii, off[:-1], off[1:])
    assert acgt.tolist() == g['baseCount']


def test_splitter_random_scaffolds_vs_literal_loop():
    from dbb_b200.splitScaffolds import SplitScaffolds
    seqs = _random_scaffolds(5, 400)
    seqs.append('ACGT' * 300 + 'N' * 70000 + 'ACGT' * 10 + 'n' * 5 + 'A')      # a run spanning many 2 048-base steps
    seqs.append('N' * 4096)
    ids = ['x%d' % i for i in range(len(seqs))]
    ss = SplitScaffolds()
    for minN in (0, 4, 10, 64):
        got = ss.splitsBatch(ids, seqs, minN)
        for sid, s, d in zip(ids, seqs, got):
            assert d == O.splits_literal(sid, s, minN), (sid, minN)


def test_sequence_statistics_equal_the_per_scaffold_reference_loop(tmp_path):
    from dbb_b200.preprocess import Preprocess
    from dbb_b200.genomicSignatures import GenomicSignatures
    rng = np.random.RandomState(11)
    seqs = [s for s in _random_scaffolds(6, 150, max_len=9000) if sum(ch in 'ACGTacgtUu' for ch in s) > 0]
    # keep scaffolds whose every contig has at least one base the GC can be computed from (upstream divides by zero)
    good = {}
    for i, s in enumerate(seqs):
        c = O.splits_literal('k%d' % i, s, 10)
        if all(sum(ch in 'ACGTacgtUu' for ch in s[b:e]) > 0 for b, e in c.values()):
            good['k%d' % i] = s
    assert len(good) > 100
    sStats, cStats = Preprocess().sequenceStatistics(good, 10)
    n_multi = 0
    for sid, s in good.items():
        o_s, o_c = O.sequence_stats_literal(sid, s, 10)
        assert sStats[sid][0] == o_s[0] and sStats[sid][1] == o_s[1]
        assert np.array_equal(sStats[sid][2], o_s[2], equal_nan=True)            # float64 from exact counts: bit-identical
        assert set(cStats[sid]) == set(o_c)
        n_multi += len(o_c) > 1
        for cid, want in o_c.items():
            got = cStats[sid][cid]
            assert got[0] == want[0] and got[1] == want[1] and got[3:] == want[3:], cid
            assert np.array_equal(got[2], want[2], equal_nan=True), cid
    assert n_multi > 20
    bad = dict(good); bad['empty'] = 'NNNNNNNNNNNNNNNNNNNNNNNNACGT'                 # leading run -> empty contig -> upstream ZeroDivisionError
    with pytest.raises(ZeroDivisionError):
        Preprocess().sequenceStatistics(bad, 10)
    # TSV files in the reference's format
    gs = GenomicSignatures(4, 1)
    cov = {k: 10.0 + i for i, k in enumerate(list(sStats) + [c for v in cStats.values() for c in v])}
    files = [str(tmp_path / f) for f in ('s.tsv', 's.tetra.tsv', 'c.tsv', 'c.tetra.tsv')]
    Preprocess().writeStatistics(gs, sStats, cStats, cov, {}, *files)
    lines = open(files[1]).read().split('\n')
    assert lines[0] == O.tsv_header().replace('Sequence Id', 'Scaffold Id').rstrip('\n')
    first = next(iter(sStats))
    assert lines[1] == O.tsv_row(first, sStats[first][2]).rstrip('\n')
    assert open(files[3]).readline().startswith('Contig Id\t')
    assert sum(1 for _ in open(files[2])) == 1 + sum(len(v) for v in cStats.values())
