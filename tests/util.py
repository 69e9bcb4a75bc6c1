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


from oracle.checks import check_topk   # noqa: E402,F401  (shared with bench.py's parity check)
