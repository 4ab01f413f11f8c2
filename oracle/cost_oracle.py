"""TEST INFRASTRUCTURE ONLY -- float64 numpy restatement of the cost build.

Each function cites the reference lines it follows.  Pinned against the
reference's own code (imported with stubs) by ``tests/golden/make_golden.py``;
the resulting vectors are committed under ``tests/golden/``.
"""
from __future__ import annotations

import numpy as np

COST_SCALE = 10 ** 6   # integer scale; mirrors /root/reference/cytospace/cytospace.py:337


def normalize_data(data):
    """cytospace/common/common.py:142-147 -- nan->0, per-column TPM, log2(x+1), nan->0."""
    data = np.nan_to_num(np.asarray(data)).astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        data = data * (10 ** 6 / np.sum(data, axis=0, dtype=float))
    data = np.log2(data + 1)
    return np.nan_to_num(data)


def matrix_correlation_pearson(v1, v2):
    """cytospace/common/common.py:190-199 -- returns [cols(v2) x cols(v1)] Pearson r
    (population sigma, ddof=0); ValueError when the gene counts differ (:191-192)."""
    v1 = np.asarray(v1, dtype=np.float64); v2 = np.asarray(v2, dtype=np.float64)
    if v1.shape[0] != v2.shape[0]:
        raise ValueError("The two matrices v1 and v2 must have equal dimensions; "
                         "ST and scRNA data must have the same genes")
    g = v1.shape[0]
    sums = np.multiply.outer(v2.sum(0), v1.sum(0))
    stds = np.multiply.outer(v2.std(0), v1.std(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (v2.T.dot(v1) - sums / g) / stds / g


def calculate_cost(sc, st, cell_number_to_node_assignment):
    """linear_assignment_solvers.py:42-69, Pearson / non-CSPR branch (:53-55) and the
    slot expansion (:63-66).  Returns (distance_repeat float64[n x N], location_repeat int[n])."""
    cost = -matrix_correlation_pearson(sc, st)
    cn = np.asarray(cell_number_to_node_assignment)
    location_repeat = np.repeat(np.arange(len(cn)), cn).astype(int)
    return cost[location_repeat, :], location_repeat


def quantise(cost_f64):
    """Integer cost the device LAP solves: round-half-even(cost * 1e6) as int32.
    (Scale precedent: cytospace.py:337; rounding mode = the GPU epilogue's
    cvt.rni.s32.f32, so the two agree wherever the float inputs agree.)"""
    return np.rint(np.asarray(cost_f64, dtype=np.float64) * COST_SCALE).astype(np.int32)


def cost_matrix_i32(sc, st):
    """Compact [S x N] integer cost (no slot expansion): quantise(-pearson)."""
    return quantise(-matrix_correlation_pearson(sc, st))
