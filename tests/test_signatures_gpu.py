"""Parity of stage 1 (tetramer counting + normalisation) on the GPU against the oracle and the golden
vectors produced by the reference's own code.  Integer counts: bit-exact.  float32 frequencies:
|f32 - f64| <= 1.2e-7 * f64.  seqSignature (float64): bit-identical to the reference."""
import numpy as np
import pytest

from oracle import dbb_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gs():
    from dbb_b200.genomicSignatures import GenomicSignatures
    return GenomicSignatures(4, 1)


def test_golden_kat_signatures_bit_identical(gs):
    kat = util.golden_json('ref_kat_seqs.json')
    ref = util.golden_npz('ref_signatures.npz')['sigs']
    for s, g in zip(kat, ref):
        got = gs.seqSignature(s)
        assert got.dtype == np.float64 and got.shape == (136,)
        assert np.array_equal(got, g, equal_nan=True), repr(s[:40])


def test_kat_by_hand(gs):
    cols = gs.canonicalKmerOrder()
    c = gs.seqCounts('ACGT'); assert c.sum() == 1 and c[cols.index('ACGT')] == 1
    assert np.array_equal(gs.seqCounts('AAAA'), gs.seqCounts('TTTT')) and gs.seqCounts('AAAA')[cols.index('AAAA')] == 1
    assert gs.seqCounts('AAAAA')[cols.index('AAAA')] == 2
    assert gs.seqCounts('ACGNACGT').sum() == 1 and gs.seqCounts('acgtACGT').sum() == 1
    for s in ('', 'ACG', 'NNNN'):
        assert gs.seqCounts(s).sum() == 0 and np.isnan(gs.seqSignature(s)).all()


def _check_batch(gs, seqs):
    counts, freq = gs.signatures(seqs)
    for i, s in enumerate(seqs):
        exp = O.seq_counts_np(s)
        assert np.array_equal(counts[i].astype(np.int64), exp), 'counts differ for sequence %d (len %d)' % (i, len(s))
        tot = exp.sum()
        if tot == 0:
            assert np.isnan(freq[i]).all()
        else:
            f64 = exp / tot
            assert (np.abs(freq[i].astype(np.float64) - f64) <= util.FREQ_RTOL * f64).all()


def test_ragged_lengths_exact(gs):
    rng = np.random.RandomState(1)
    lens = list(range(0, 12)) + [15, 16, 17, 31, 32, 33, 60, 61, 62, 63, 64, 65, 66, 67, 68, 95, 96, 97, 127, 128, 129,
                                 130, 191, 192, 193, 255, 256, 257, 1023, 1024, 1025, 2047, 2048, 2049, 2559, 2560, 2561, 2600]
    seqs = [util.random_dna(rng, L) for L in lens] + [util.random_dna(rng, L, 0.02) for L in lens]
    _check_batch(gs, seqs)


def test_every_kernel_class_and_boundaries_exact(gs):
    """Lengths straddling the live S=1/2/3/4 class limits (read from the library: dbb_count_class_limits, in blocks of
    64 bases) by one base either side, and the chunking of very long scaffolds (one chunk, chunk + 1 block, two chunks
    + 1 base, a chunk count that is not a multiple of 3), clean and with invalid bases scattered / in runs / at the ends.
    A poly-A scaffold as long as one S = 4 chunk puts every super-k-mer of the chunk into ONE 16-bit counter."""
    import ctypes
    from dbb_b200 import _lib
    lim = (ctypes.c_int64 * 4)()
    _lib.check(_lib.lib().dbb_count_class_limits(lim))
    t1, t2, t3, chunk = [int(x) for x in lim]
    assert 0 <= t1 <= t2 <= t3 < chunk <= 4095
    rng = np.random.RandomState(2)
    lens = []
    for t in (t1, t2, t3, chunk):
        lens += [t * 64 - 1, t * 64, t * 64 + 1, t * 64 + 64, t * 64 + 65]
    lens += [3000, 5001, 33333, 250001, 2 * chunk * 64 + 1, 4 * chunk * 64 + 12345]
    seqs = []
    for L in lens:
        a = util.random_dna(rng, L)
        seqs.append(a)
        b = util.random_dna(rng, L, 0.001)
        s = rng.randint(0, max(1, L - 60)); b[s:s + rng.randint(1, 50)] = ord('N')
        s = rng.randint(0, max(1, L - 600)); b[s:s + rng.randint(10, 500)] |= 0x20
        b[:2] = ord('N'); b[-3:] = ord('n')
        seqs.append(b)
    seqs.append(np.full(chunk * 64, ord('A'), dtype=np.uint8))
    seqs.append(np.full(t3 * 64, ord('G'), dtype=np.uint8))
    seqs.append(np.full(t2 * 64, ord('T'), dtype=np.uint8))
    _check_batch(gs, seqs)


def test_low_complexity_and_all_invalid(gs):
    seqs = [np.full(100000, ord('A'), dtype=np.uint8), np.frombuffer(b'ACGT' * 30000, dtype=np.uint8),
            np.full(5000, ord('N'), dtype=np.uint8), np.frombuffer(b'AC' * 700, dtype=np.uint8),
            np.frombuffer(b'acgt' * 1000, dtype=np.uint8)]
    _check_batch(gs, seqs)


def test_reverse_complement_invariance_and_window_count(gs):
    rng = np.random.RandomState(3)
    comp = np.zeros(256, dtype=np.uint8); comp[list(b'ACGT')] = list(b'TGCA')
    seqs = [util.random_dna(rng, L) for L in (100, 5000, 70000, 300000)]
    rc = [comp[s][::-1].copy() for s in seqs]
    c1, _ = gs.signatures(seqs); c2, _ = gs.signatures(rc)
    assert np.array_equal(c1, c2)
    assert [int(x) for x in c1.sum(axis=1)] == [len(s) - 3 for s in seqs]


def test_device_pipeline_matches_host_generator_and_oracle():
    """synthetic.generate_sequences (host) and dbb_synth_ascii_dev (device) emit identical bytes; the
    device pack -> count pipeline gives exact counts for every scaffold of a small metagenome."""
    import torch
    from dbb_b200 import device, synthetic
    meta = synthetic.make_metagenome(seed=11, n_scaffolds=300, min_len=1000, max_len=50000, n_genomes=5,
                                     frac_n=0.05, frac_lower=0.05, extra_lengths=(1200000,))
    host = synthetic.generate_sequences(meta)
    d_ascii, seq_off = device.synth_ascii(meta)
    flat = d_ascii.cpu().numpy()
    for i, s in enumerate(host):
        assert np.array_equal(flat[seq_off[i]:seq_off[i + 1]], s), 'generator mismatch at scaffold %d' % i
    packed = device.pack_ascii(d_ascii, seq_off)
    plan = device.CountPlan(packed)
    counts = torch.zeros((meta.n_scaffolds, 136), dtype=torch.int32, device='cuda')
    freq = torch.zeros((meta.n_scaffolds, 140), dtype=torch.float32, device='cuda')
    plan.run(counts, freq)
    torch.cuda.synchronize()
    c = counts.cpu().numpy().astype(np.int64); f = freq.cpu().numpy()
    for i, s in enumerate(host):
        exp = O.seq_counts_np(s)
        assert np.array_equal(c[i], exp), i
        f64 = exp / exp.sum()
        assert (np.abs(f[i, :136] - f64) <= util.FREQ_RTOL * f64).all() and (f[i, 136:] == 0).all()
    plan.run(counts, freq)           # the plan is re-runnable
    torch.cuda.synchronize()
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), c)


def test_calculate_writes_reference_tsv(gs, tmp_path):
    rng = np.random.RandomState(5)
    seqs = [('s%d some description' % i, util.random_dna(rng, L).tobytes().decode()) for i, L in enumerate((500, 64, 3, 7777))]
    fa = tmp_path / 'in.fna'
    with open(fa, 'w') as f:
        for sid, s in seqs:
            f.write('>%s\n' % sid)
            for o in range(0, len(s), 70):
                f.write(s[o:o + 70] + '\n')
    out = tmp_path / 'out.tsv'
    gs.calculate(str(fa), str(out))
    lines = open(out).read().split('\n')
    assert lines[0] + '\n' == O.tsv_header()
    for (sid, s), line in zip(seqs, lines[1:]):
        with np.errstate(invalid='ignore', divide='ignore'):
            assert line + '\n' == O.tsv_row(sid.split()[0], O.seq_signature_literal(s))
    back = gs.read(str(out))
    assert set(back) == {'s0', 's1', 's2', 's3'} and back['s0'].shape == (136,)
