"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# tolerances stated by BASELINE.md / SURVEY.md section 8d
FREQ_RTOL = 1.2e-7        # |f32 - f64| <= 1.2e-7 * f64  (1 ulp FP32)
DIST_ATOL = 2e-6          # |d32 - d64| <= 2e-6 absolute


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def unpack_mask(bits, n):
    return np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), axis=1, bitorder='little')[:, :n]


def random_dna(rng, length, p_invalid=0.0):
    if length == 0:
        return np.zeros(0, dtype=np.uint8)
    a = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.randint(0, 4, length)].copy()
    if p_invalid > 0:
        bad = rng.uniform(size=length) < p_invalid
        a[bad] = np.frombuffer(b'NacgtnRY', dtype=np.uint8)[rng.randint(0, 8, int(bad.sum()))]
    return a


def check_mask_against_oracle(mask_rows, oracle_rows, d64=None, bound=None, atol=DIST_ATOL):
    """0/1 rows must agree everywhere except (TD only) pairs whose float64 distance lies within
    ``atol`` of the bound.  Returns the number of such tolerated flips."""
    diff = mask_rows != oracle_rows
    if not diff.any():
        return 0
    assert d64 is not None, 'mask mismatch: %d entries' % int(diff.sum())
    near = np.abs(d64 - bound) <= atol
    assert not (diff & ~near).any(), 'mask mismatch outside the %g band: %d' % (atol, int((diff & ~near).sum()))
    return int(diff.sum())


def check_topk(idx, dist, d64_rows, rows, k, allowed=None, atol=DIST_ATOL):
    """GPU top-k (idx, dist) vs float64 distance rows: distances within atol of float64, and the index
    set equal to the oracle's up to swaps among candidates within atol of the oracle's k-th distance."""
    n = d64_rows.shape[1]
    for r, i in enumerate(rows):
        ok = ~np.isnan(d64_rows[r])
        ok[i] = False
        if allowed is not None:
            ok &= allowed[r].astype(bool)
        cand = np.nonzero(ok)[0]
        m = min(k, cand.size)
        got = idx[r]
        assert (got[m:] == -1).all() and np.isinf(dist[r][m:]).all(), 'row %d: padding' % i
        got = got[:m]
        assert len(set(got.tolist())) == m and ok[got].all(), 'row %d: invalid/duplicate neighbours' % i
        assert np.abs(dist[r][:m] - d64_rows[r][got]).max(initial=0) <= atol, 'row %d: distance error' % i
        assert (np.diff(dist[r][:m]) >= 0).all(), 'row %d: not sorted' % i
        if m == 0:
            continue
        order = cand[np.lexsort((cand, d64_rows[r][cand]))]
        kth = d64_rows[r][order[m - 1]]
        must = set(order[d64_rows[r][order] < kth - atol].tolist())        # clearly inside
        may = set(order[d64_rows[r][order] <= kth + atol].tolist())        # inside or in the tie band
        g = set(got.tolist())
        assert must <= g <= may, 'row %d: neighbour set differs beyond the tie band' % i
