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


def matrix_correlation_spearman(v1, v2):
    """cytospace/common/common.py:202-215 -- per-column average ranks (pd.DataFrame.rank()
    defaults), then the Pearson formula on the ranks."""
    import pandas as pd
    v1 = np.asarray(v1, dtype=np.float64); v2 = np.asarray(v2, dtype=np.float64)
    if v1.shape[0] != v2.shape[0]:
        raise ValueError("The two matrices v1 and v2 must have equal dimensions; "
                         "ST and scRNA data must have the same genes")
    return matrix_correlation_pearson(pd.DataFrame(v1).rank().values, pd.DataFrame(v2).rank().values)


def average_ranks(x):
    """`pd.DataFrame(x).rank().values` (common.py:207-208) restated without pandas:
    rank = #less + (#equal + 1) / 2 per column."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    for c in range(x.shape[1]):
        col = np.sort(x[:, c])
        out[:, c] = (np.searchsorted(col, x[:, c], "left") + np.searchsorted(col, x[:, c], "right") + 1) / 2.0
    return out


def euclidean_distance(sc, st):
    """linear_assignment_solvers.py:51,59 -- transpose(cdist(sc.T, st.T, 'euclidean')): [S x N]."""
    sc = np.asarray(sc, dtype=np.float64); st = np.asarray(st, dtype=np.float64)
    diff = st.T[:, None, :] - sc.T[None, :, :] if sc.size * st.shape[1] < 5e7 else None
    if diff is not None:
        return np.sqrt((diff * diff).sum(-1))
    from scipy.spatial import distance
    return np.transpose(distance.cdist(sc.T, st.T, "euclidean"))


METRICS = ("Pearson_correlation", "Spearman_correlation", "Euclidean")


def metric_cost(sc, st, distance_metric="Pearson_correlation"):
    """The compact [S x N] float64 `cost` of calculate_cost (linear_assignment_solvers.py:46-59;
    the lap_CSPR branch :46-52 builds the same matrix through a double transpose)."""
    if distance_metric == "Pearson_correlation":
        return -matrix_correlation_pearson(sc, st)
    if distance_metric == "Spearman_correlation":
        return -matrix_correlation_spearman(sc, st)
    if distance_metric == "Euclidean":
        return euclidean_distance(sc, st)
    raise ValueError(f"unknown distance metric {distance_metric}")


def calculate_cost(sc, st, cell_number_to_node_assignment, distance_metric="Pearson_correlation"):
    """linear_assignment_solvers.py:42-69: the metric branches (:46-59) and the slot expansion
    (:63-66).  Returns (distance_repeat float64[n x N], location_repeat int[n])."""
    cost = metric_cost(sc, st, distance_metric)
    cn = np.asarray(cell_number_to_node_assignment)
    location_repeat = np.repeat(np.arange(len(cn)), cn).astype(int)
    return cost[location_repeat, :], location_repeat


def quantise(cost_f64):
    """Integer cost the device LAP solves: round-half-even(cost * 1e6) as int32.
    (Scale precedent: cytospace.py:337; rounding mode = the GPU epilogue's
    cvt.rni.s32.f32, so the two agree wherever the float inputs agree.)"""
    return np.rint(np.asarray(cost_f64, dtype=np.float64) * COST_SCALE).astype(np.int32)


def cost_matrix_i32(sc, st, distance_metric="Pearson_correlation"):
    """Compact [S x N] integer cost (no slot expansion): quantise(metric cost)."""
    return quantise(metric_cost(sc, st, distance_metric))


# ---- lap_CSPR (cytospace.py:334-347) ------------------------------------------------------

def cspr_matrix_reference(distance_repeat, seed):
    """cytospace.py:335-340 verbatim in numpy: the integer matrix the reference hands to ortools
    (cells x slots after the transpose), noise from the global legacy MT19937 stream."""
    np.random.seed(seed)
    cost_scaled = 10 ** 6 * distance_repeat + 10 * np.random.rand(*distance_repeat.shape) + 1
    return np.transpose(cost_scaled).astype(int)


_M64 = (1 << 64) - 1


def _mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hash_noise(seed, n_rows, n_cols, lo=1, span=10):
    """The device's counter-based replacement for `10 * np.random.rand(..) + 1` (cytospace.py:337):
    lo + floor(span * U) with U = top 32 bits of splitmix64(splitmix64(seed ^ i*C) + j) / 2^32
    (cytospace_b200/csrc/metrics.cu: expand_rows_noise_kernel)."""
    with np.errstate(over="ignore"):
        i = np.arange(n_rows, dtype=np.uint64)
        rowkey = _mix64(np.uint64(seed & _M64) ^ (i * np.uint64(0xD1B54A32D192ED03)))
        h = _mix64(rowkey[:, None] + np.arange(n_cols, dtype=np.uint64)[None, :])
        u32 = h >> np.uint64(32)
        return (lo + ((u32 * np.uint64(span)) >> np.uint64(32)).astype(np.int64)).astype(np.int32)


def cspr_matrix_i32(cost_i32, cell_number_to_node_assignment, seed):
    """Integer lap_CSPR matrix of the device path: quantised cost rows repeated per slot plus the
    hash noise, [n slots x N cells] (the reference's matrix is its transpose)."""
    cn = np.asarray(cell_number_to_node_assignment)
    location_repeat = np.repeat(np.arange(len(cn)), cn).astype(int)
    expanded = np.asarray(cost_i32, dtype=np.int32)[location_repeat, :]
    return expanded + hash_noise(seed, expanded.shape[0], expanded.shape[1])
