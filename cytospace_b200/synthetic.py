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


def structured_counts_torch(n_cells, n_spots, n_genes, cells_per_spot=1, seed=1001, n_types=20, depth=1500.0,
                            spot_depth_factor=10.0, device="cuda", dtype=None, block=4096):
    """Same distribution as ``structured_counts`` sampled with torch on ``device`` (bench-sized
    inputs in seconds instead of minutes; NOT the same random stream).  Returns
    (sc_counts [G x N], st_counts [G x S], cn int64 ndarray[S])."""
    import torch
    dtype = dtype or torch.float64
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    cn = np.full(n_spots, cells_per_spot, dtype=np.int64) if np.isscalar(cells_per_spot) \
        else np.asarray(cells_per_spot, dtype=np.int64)
    base = torch.randn(n_genes, generator=gen, device=device) * 1.5 - 2.0
    mu = base.repeat(n_types, 1)
    marker = torch.rand((n_types, n_genes), generator=gen, device=device) < 0.02
    mu = mu + marker * (torch.randn((n_types, n_genes), generator=gen, device=device) * 0.5 + 2.0)
    rate = torch.softmax(mu, dim=1)                                      # [K x G]
    sc = torch.empty((n_genes, n_cells), dtype=dtype, device=device)
    sc_type = torch.randint(0, n_types, (n_cells,), generator=gen, device=device)
    for c0 in range(0, n_cells, block):
        t = sc_type[c0:c0 + block]
        sc[:, c0:c0 + block] = torch.poisson(depth * rate[t], generator=gen).T.to(dtype)
    st = torch.empty((n_genes, n_spots), dtype=dtype, device=device)
    cn_t = torch.from_numpy(np.maximum(cn, 1)).to(device)
    kmax = int(cn_t.max().item())
    for s0 in range(0, n_spots, block):
        k = cn_t[s0:s0 + block]
        types = torch.randint(0, n_types, (k.numel(), kmax), generator=gen, device=device)
        w = (torch.arange(kmax, device=device)[None, :] < k[:, None]).to(rate.dtype)     # first cn[s] draws count
        mix = torch.zeros((k.numel(), n_types), device=device, dtype=rate.dtype)
        mix.scatter_add_(1, types, w)
        lam = depth * spot_depth_factor * (mix @ rate)
        st[:, s0:s0 + block] = torch.poisson(lam, generator=gen).T.to(dtype)
    return sc, st, cn


def normalize_data_torch(x):
    """``normalize_data`` (cytospace/common/common.py:142-147) with torch ops -- bench input
    preparation only (the product's fused version is the ``log_tpm`` flag of cyb_standardise)."""
    import torch
    x = torch.nan_to_num(x).to(torch.float64)
    x = x * (1e6 / x.sum(0, keepdim=True))
    return torch.nan_to_num(torch.log2(x + 1))
