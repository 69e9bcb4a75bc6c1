"""Seeded synthetic metagenomes of the BASELINE.json shapes.

The base at (scaffold s, position p) is a pure function of (seed, s, p, previous three bases), built
on a counter-based 64-bit mixer, so the host generator here (numpy; small configs and CPU tests) and
the device generator ``dbb_synth_ascii`` (csrc/synth.cu; large configs) emit IDENTICAL bytes.

Model: G genomes, genome g has GC_g ~ U(0.25, 0.75) and an order-3 Markov chain whose 64x4
transition rows are Dirichlet(20 * 4 * p_base) draws around the GC-implied base frequencies, so
scaffolds of one genome are tetranucleotide-distance neighbours and different genomes mostly are
not.  The Markov context is re-seeded from the hash every CHUNK bases so that generation is
chunk-parallel.  1 % of scaffolds carry one run of N, 1 % one lowercase stretch (reference fact F6:
lowercase is invalid for genomicSignatures.py:134-144).
"""
import numpy as np

CHUNK = 64
_K_S = np.uint64(0x9E3779B97F4A7C15)
_K_P = np.uint64(0xD1B54A32D192ED03)
_K_C = np.uint64(0x8CB92BA72F3D8DD7)

# dbb/data/td_dist.txt length keys (reference); lengths at exact midpoints are re-drawn
TD_LEN_KEYS = [500, 600, 700, 800, 900, 1000, 1200, 1400, 1600, 1800, 2000, 2500, 3000, 3500, 4000,
               4500, 5000, 6000, 7000, 8000, 9000, 10000, 15000, 20000, 25000, 30000, 35000, 40000,
               45000, 50000, 60000, 70000, 80000, 90000, 100000, 200000, 300000, 400000, 600000,
               800000, 1000000]
_MIDPOINTS = set((a + b) // 2 for a, b in zip(TD_LEN_KEYS[:-1], TD_LEN_KEYS[1:]) if (a + b) % 2 == 0)


def mix64(x):
    """splitmix64 finaliser over uint64 arrays (wraps mod 2^64)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over='ignore'):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


class Metagenome(object):
    """Per-scaffold scalars + per-genome Markov tables."""

    def __init__(self):
        self.seed = 0
        self.genome = None      # int32[N]
        self.length = None      # int64[N]
        self.coverage = None    # float64[N]
        self.gc_obs = None      # float64[N]  (fraction, 0..1)
        self.genome_gc = None   # float64[G]
        self.cum = None         # uint32[G, 64, 3]  24-bit cumulative thresholds
        self.n_start = None     # int64[N] (-1: none)
        self.n_len = None       # int64[N]
        self.lc_start = None    # int64[N]
        self.lc_len = None      # int64[N]

    @property
    def n_scaffolds(self):
        return int(self.length.shape[0])

    @property
    def total_bp(self):
        return int(self.length.sum())


def _draw_lengths(rng, n, min_len, max_len, law):
    if law == 'loguniform':
        draw = lambda m: np.floor(np.exp(rng.uniform(np.log(min_len), np.log(max_len + 1), m))).astype(np.int64)
    else:
        draw = lambda m: rng.randint(min_len, max_len + 1, m).astype(np.int64)
    L = np.clip(draw(n), min_len, max_len)
    bad = np.isin(L, list(_MIDPOINTS))
    while bad.any():
        L[bad] = np.clip(draw(int(bad.sum())), min_len, max_len)
        bad = np.isin(L, list(_MIDPOINTS))
    return L


def make_metagenome(seed, n_scaffolds, min_len, max_len, n_genomes, law='loguniform',
                    frac_n=0.01, frac_lower=0.01, extra_lengths=()):
    rng = np.random.RandomState(seed)
    m = Metagenome()
    m.seed = int(seed)
    G = int(n_genomes)
    m.genome_gc = rng.uniform(0.25, 0.75, G)
    base_p = np.stack([(1 - m.genome_gc) / 2, m.genome_gc / 2, m.genome_gc / 2, (1 - m.genome_gc) / 2], axis=1)
    probs = np.empty((G, 64, 4))
    for g in range(G):
        probs[g] = rng.dirichlet(80.0 * base_p[g], 64)
    cum = np.cumsum(probs, axis=2)[:, :, :3]
    m.cum = np.minimum(np.floor(cum * (1 << 24)), (1 << 24) - 1).astype(np.uint32)
    cov_g = rng.lognormal(3.0, 1.0, G)

    N = int(n_scaffolds)
    L = _draw_lengths(rng, N, min_len, max_len, law)
    if len(extra_lengths):
        L = np.concatenate([L, np.asarray(extra_lengths, dtype=np.int64)])
        N = L.shape[0]
    m.length = L
    m.genome = rng.randint(0, G, N).astype(np.int32)
    m.coverage = cov_g[m.genome] * (1.0 + rng.normal(0.0, 0.05, N))
    gg = m.genome_gc[m.genome]
    m.gc_obs = np.clip(gg + rng.normal(0.0, 1.0, N) * np.sqrt(gg * (1 - gg) / L), 0.0, 1.0)

    m.n_start = np.full(N, -1, dtype=np.int64)
    m.n_len = np.zeros(N, dtype=np.int64)
    m.lc_start = np.full(N, -1, dtype=np.int64)
    m.lc_len = np.zeros(N, dtype=np.int64)
    pick = rng.uniform(size=N)
    for arr_s, arr_l, sel, lo, hi in ((m.n_start, m.n_len, pick < frac_n, 1, 50),
                                      (m.lc_start, m.lc_len, (pick >= frac_n) & (pick < frac_n + frac_lower), 10, 500)):
        ids = np.nonzero(sel)[0]
        run = np.minimum(rng.randint(lo, hi + 1, ids.size), L[ids])
        start = np.floor(rng.uniform(size=ids.size) * (L[ids] - run + 1)).astype(np.int64)
        arr_s[ids] = start
        arr_l[ids] = run
    return m


def generate_block(meta, scaffold_ids):
    """ASCII bytes of the given scaffolds -> list of uint8 arrays.  Vectorised over chunks."""
    ids = np.asarray(scaffold_ids, dtype=np.int64)
    if ids.size == 0:
        return []
    L = meta.length[ids]
    nchunk = (L + CHUNK - 1) // CHUNK
    tot = int(nchunk.sum())
    sc = np.repeat(ids, nchunk)                                     # scaffold of each chunk
    first = np.cumsum(nchunk) - nchunk
    ck = np.arange(tot, dtype=np.int64) - np.repeat(first, nchunk)   # chunk index within scaffold
    seedk = mix64(np.uint64(meta.seed) + np.uint64(0x1234567))
    with np.errstate(over='ignore'):
        skey = seedk ^ (sc.astype(np.uint64) * _K_S)
        ctx = (mix64(skey ^ (ck.astype(np.uint64) * _K_C)) >> np.uint64(58)).astype(np.int64)   # 6 bits
    g = meta.genome[sc].astype(np.int64)
    out = np.empty((tot, CHUNK), dtype=np.uint8)
    cum = meta.cum
    for t in range(CHUNK):
        p = (ck * CHUNK + t).astype(np.uint64)
        with np.errstate(over='ignore'):
            u = (mix64(skey + p * _K_P) >> np.uint64(40)).astype(np.uint32)
        th = cum[g, ctx]                                            # [tot, 3]
        b = (u >= th[:, 0]).astype(np.int64) + (u >= th[:, 1]) + (u >= th[:, 2])
        out[:, t] = b
        ctx = ((ctx << 2) | b) & 63
    ascii_lut = np.frombuffer(b'ACGT', dtype=np.uint8)
    out = ascii_lut[out]
    seqs = []
    for n, s in enumerate(ids):
        a = out[first[n]:first[n] + nchunk[n]].reshape(-1)[:L[n]].copy()
        if meta.n_start[s] >= 0:
            a[meta.n_start[s]:meta.n_start[s] + meta.n_len[s]] = ord('N')
        if meta.lc_start[s] >= 0:
            a[meta.lc_start[s]:meta.lc_start[s] + meta.lc_len[s]] |= 0x20
        seqs.append(a)
    return seqs


def generate_sequences(meta, scaffold_ids=None, block_bp=8 << 20):
    """All (or the given) scaffolds as a list of uint8 ASCII arrays, generated in bounded blocks."""
    ids = np.arange(meta.n_scaffolds) if scaffold_ids is None else np.asarray(scaffold_ids)
    seqs, cur, cur_bp = [], [], 0
    for s in ids:
        cur.append(s)
        cur_bp += int(meta.length[s])
        if cur_bp >= block_bp:
            seqs.extend(generate_block(meta, cur))
            cur, cur_bp = [], 0
    seqs.extend(generate_block(meta, cur))
    return seqs


def concat(seqs):
    """list of uint8 arrays -> (uint8 concat, int64 offsets[N+1])."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.concatenate(seqs) if len(seqs) else np.zeros(0, dtype=np.uint8)
    return flat, off


# ---- synthetic distribution tables (upstream gc_dist.txt is missing; coverage_dist.txt is a
# ---- run-time product of preprocess.py:319-337) ----------------------------------------------

def _pct_keys():
    return np.arange(0, 100 + 0.5, 0.5).tolist()            # preprocess.py:328


def synthetic_gc_dist(length_keys=TD_LEN_KEYS):
    """gcDist[meanGC][len][pct] = z_pct * sqrt(GC(1-GC)/len); meanGC 0.20..0.80 step 0.01.
    Flagged synthetic: same nesting as the dict distributions.py:57-63,75-87 reads."""
    from scipy.stats import norm
    q = _pct_keys()
    z = np.clip(norm.ppf(np.array(q) / 100.0), -8.0, 8.0)
    d = {}
    for m in range(20, 81):
        gc = round(m / 100.0, 2)
        d[gc] = {}
        for L in length_keys:
            sd = np.sqrt(gc * (1 - gc) / L)
            d[gc][L] = dict(zip(q, (z * sd).tolist()))
    return d


def synthetic_cov_dist(seed=0, percent=0.2, min_len=2500, max_len=1000000):
    """covDist[testLen][pct] = percentile of % coverage error; test lengths grow as
    ``int(testLen + testLen * percent)`` (preprocess.py:196), percentiles as preprocess.py:328-331."""
    rng = np.random.RandomState(seed + 991)
    q = _pct_keys()
    d = {}
    testLen = min_len
    while testLen <= max_len:
        sd = 100.0 * min(0.5, np.sqrt(60.0 / testLen) + 0.05)
        pts = rng.normal(0.0, sd, 1000)
        d[testLen] = dict(zip(q, np.percentile(pts, q).tolist()))
        testLen = int(testLen + testLen * percent)
    return d


# ---- the five BASELINE.json configs ---------------------------------------------------------------

CONFIGS = {
    'C1': dict(seed=1, n_scaffolds=2000, min_len=1000, max_len=50000, n_genomes=20),
    'C2': dict(seed=2, n_scaffolds=50000, min_len=1000, max_len=100000, n_genomes=200),
    'C3': dict(seed=3, n_scaffolds=250000, min_len=1000, max_len=100000, n_genomes=200),
    'C4': dict(seed=4, n_scaffolds=1000000, min_len=1000, max_len=14000, n_genomes=2000),
    'C5': dict(seed=5, n_scaffolds=10000000, min_len=1000, max_len=2000, n_genomes=2000, law='uniform',
               extra_lengths=(5000000,) * 8),
}


def make_config(name, scale=1.0):
    kw = dict(CONFIGS[name])
    if scale != 1.0:
        kw['n_scaffolds'] = max(1, int(kw['n_scaffolds'] * scale))
    return make_metagenome(**kw)
