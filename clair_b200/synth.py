"""Synthetic pile-up tensors shaped like CreateTensor.generate_tensor output
(reference dataPrepScripts/CreateTensor.py:29-65), after the generator's
``channels 1..3 -= channel 0`` transform (clair/utils.py:96-98).  Recipe: SURVEY.md section 8d.

Channel meaning per (position, row=ACGTacgt): 0 reference-base count, 1 query-base count
(+ insertions), 2 reference-base count (+ deletions), 3 query-base count (SNP evidence).
"""
import numpy as np

T, R, C = 33, 8, 4


def synthetic_counts(n, seed=20240607, depth_lo=8, depth_hi=80):
    """Raw integer counts [n,33,8,4] (int16), before channel subtraction."""
    rng = np.random.default_rng(seed)
    depth = rng.integers(depth_lo, depth_hi + 1, size=(n, 1))
    ref_base = rng.integers(0, 4, size=(n, T))
    alt_base = (ref_base + rng.integers(1, 4, size=(n, T))) % 4
    fwd = rng.binomial(depth, 0.5, size=(n, T))            # forward-strand share of the depth
    rev = depth - fwd
    # mismatch rate: sequencing noise everywhere, a het/hom variant at the centre of ~half the sites
    rate = rng.beta(1.0, 12.0, size=(n, T)) * 0.6
    centre = rng.choice([0.0, 0.5, 1.0], size=n, p=[0.5, 0.3, 0.2])
    rate[:, T // 2] = np.maximum(rate[:, T // 2], centre * rng.uniform(0.8, 1.0, size=n))
    mis_f = rng.binomial(fwd, rate)
    mis_r = rng.binomial(rev, rate)
    dele_f = rng.binomial(fwd, 0.03)
    dele_r = rng.binomial(rev, 0.03)
    x = np.zeros((n, T, R, C), dtype=np.int32)
    ii, tt = np.meshgrid(np.arange(n), np.arange(T), indexing="ij")
    for strand, cov, mis, dele in ((0, fwd, mis_f, dele_f), (4, rev, mis_r, dele_r)):
        x[ii, tt, ref_base + strand, 0] += cov
        x[ii, tt, ref_base + strand, 2] += cov + dele
        x[ii, tt, ref_base + strand, 1] += cov - mis
        x[ii, tt, ref_base + strand, 3] += cov - mis
        x[ii, tt, alt_base + strand, 1] += mis
        x[ii, tt, alt_base + strand, 3] += mis
    ins = rng.poisson(0.4, size=(n, T, R)) * (rng.random((n, T, R)) < 0.15)
    x[..., 1] += ins
    return x.astype(np.int16)


def synthetic_tensors(n, seed=20240607, **kw):
    """float32 [n,33,8,4] as ``tensor_generator_from`` would yield them (small signed integers)."""
    x = synthetic_counts(n, seed, **kw).astype(np.float32)
    x[..., 1:] -= x[..., 0:1]
    return x
