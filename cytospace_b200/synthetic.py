"""Synthetic N-cell x S-spot x G-gene inputs for the assignment hot path.

The reference ships no datasets or tests; SURVEY.md section 8(d) fixes the
generators used by the parity tests and bench.py.  Structured (primary): K cell
types with sparse marker genes, Poisson counts at the reference's down-sample
depth of 1500 transcripts per cell (/root/reference/cytospace/common/
argument_parser.py:58, common.py:149-173); each ST spot is the re-Poissoned sum
of `cn[s]` cells of random types at 10x depth.  Output is RAW counts laid out
like the reference's DataFrames after ``to_numpy()``: genes x cells.
"""
from __future__ import annotations

import numpy as np


def structured_counts(n_cells, n_spots, n_genes, cells_per_spot=1, seed=1001, n_types=20,
                      depth=1500.0, spot_depth_factor=10.0):
    """Returns (sc_counts [G x N] float64, st_counts [G x S] float64, cn int64[S])."""
    rng = np.random.default_rng(seed)
    cn = np.full(n_spots, cells_per_spot, dtype=np.int64) if np.isscalar(cells_per_spot) \
        else np.asarray(cells_per_spot, dtype=np.int64)
    base = rng.normal(-2.0, 1.5, size=n_genes)
    mu = np.tile(base, (n_types, 1))
    for k in range(n_types):
        marker = rng.random(n_genes) < 0.02
        mu[k, marker] += rng.normal(2.0, 0.5, size=int(marker.sum()))
    rate = np.exp(mu - mu.max(axis=1, keepdims=True))
    rate /= rate.sum(axis=1, keepdims=True)           # softmax per type  [K x G]
    sc_type = rng.integers(0, n_types, size=n_cells)
    sc = rng.poisson(depth * rate[sc_type]).T.astype(np.float64)       # [G x N]
    st = np.empty((n_genes, n_spots), dtype=np.float64)
    for s in range(n_spots):
        k = max(int(cn[s]), 1)
        types = rng.integers(0, n_types, size=k)
        lam = depth * spot_depth_factor * rate[types].sum(axis=0)
        st[:, s] = rng.poisson(lam)
    return sc, st, cn


def unstructured_counts(n_cells, n_spots, n_genes, lam=0.3, seed=7):
    """iid Poisson stress input: near-zero correlations, many near-ties."""
    rng = np.random.default_rng(seed)
    sc = rng.poisson(lam, size=(n_genes, n_cells)).astype(np.float64)
    st = rng.poisson(lam * 10, size=(n_genes, n_spots)).astype(np.float64)
    return sc, st, np.ones(n_spots, dtype=np.int64)


def uniform_cost_i32(n, seed=11, high=2_000_000):
    """LAP-only input: int32 uniform on [0, high)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, high, size=(n, n), dtype=np.int32)
