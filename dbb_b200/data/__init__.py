"""Empirical tetranucleotide-distance table used by the TD predicate.

``td_dist.npz`` holds the reference's ``dbb/data/td_dist.txt`` (41 sequence lengths x 201 percentiles of
the within-genome L1 distance; read at distributions.py:30,40-42) re-encoded as arrays by
tests/golden/make_golden.py.  ``load_td_dist()`` rebuilds the nested dict the reference API takes:
``tdDist[length][percentile] -> bound``."""
import os

import numpy as np


def load_td_dist():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'td_dist.npz'))
    lens, pcts, val = z['lengths'], z['percentiles'], z['values']
    return {int(l): {float(p): float(val[i, j]) for j, p in enumerate(pcts)} for i, l in enumerate(lens)}
