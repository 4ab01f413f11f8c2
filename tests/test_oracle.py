"""CPU: the oracle against the committed golden vectors and independent implementations."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from scipy.optimize import linear_sum_assignment

import oracle
from oracle import cost_oracle as co
from conftest import LAP_NAMES


def check_solution(cost, rowsol, colsol, total, u, v, row_map=None):
    n = len(rowsol)
    c = cost if row_map is None else cost[row_map]
    assert sorted(rowsol.tolist()) == list(range(n))
    assert np.array_equal(colsol[rowsol], np.arange(n))
    picked = c[np.arange(n), rowsol].astype(np.int64)
    assert int(picked.sum()) == total
    red = c.astype(np.int64) - u[:, None] - v[None, :]
    assert red.min() >= 0                      # dual feasibility
    assert np.all(red[np.arange(n), rowsol] == 0)   # tight on assigned pairs


@pytest.mark.parametrize("name", LAP_NAMES)
def test_jv_restatement_matches_golden(lap_golden, name):
    cost = lap_golden[f"{name}_cost"]
    rowsol, colsol, (total, u, v) = oracle.lapjv_i32(cost)
    assert total == int(lap_golden[f"{name}_opt"])          # optimal total from SciPy
    assert np.array_equal(rowsol, lap_golden[f"{name}_rowsol"])
    assert np.array_equal(colsol, lap_golden[f"{name}_colsol"])
    check_solution(cost, rowsol, colsol, total, u, v)


@pytest.mark.parametrize("name", LAP_NAMES)
def test_f64_restatement_and_auction_model(lap_golden, name):
    cost = lap_golden[f"{name}_cost"]
    opt = int(lap_golden[f"{name}_opt"])
    _, colsol, (total, _, _) = oracle.lapjv_f64(cost.astype(np.float64))
    assert int(round(total)) == opt
    for tail in (0, 2):
        po, so, tot, lam, stats, _ = oracle.auction_model(cost, tail_t=tail)       # persons = rows
        assert tot == opt and sorted(po.tolist()) == list(range(cost.shape[0]))
        assert np.array_equal(so[po], np.arange(cost.shape[0]))


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 24), st.integers(0, 2 ** 31 - 1), st.sampled_from([2, 10, 1000, 2_000_000]))
def test_jv_total_equals_scipy_property(n, seed, high):
    rng = np.random.default_rng(seed)
    cost = rng.integers(-high, high, (n, n), dtype=np.int32)
    rowsol, colsol, (total, u, v) = oracle.lapjv_i32(cost)
    ri, ci = linear_sum_assignment(cost.astype(np.float64))
    assert total == int(cost[ri, ci].astype(np.int64).sum())
    check_solution(cost, rowsol, colsol, total, u, v)
    assert oracle.auction_model(cost)[2] == total


@settings(max_examples=80, deadline=None)
@given(st.integers(1, 10), st.integers(0, 2 ** 31 - 1), st.sampled_from([2, 10, 1000, 2_000_000]), st.sampled_from([0, 2, 64]))
def test_capacitated_auction_model_equals_expanded_jv(n_obj, seed, high, tail):
    """The device algorithm (capacitated objects = spots with cn cells, transposed matrix) reaches the
    optimum of the reference's expanded LAP (LAS:63-66) -- including empty spots (cn == 0)."""
    rng = np.random.default_rng(seed)
    cap = rng.integers(0, 5, n_obj).astype(np.int32)
    if cap.sum() == 0:
        cap[0] = 1
    n = int(cap.sum())
    compact = rng.integers(-high, high, (n_obj, n), dtype=np.int32)           # spots x cells (reference orientation)
    row_map = np.repeat(np.arange(n_obj, dtype=np.int32), cap)
    want = oracle.lapjv_i32(compact, row_map)[2][0]
    po, so, tot, lam, stats, _ = oracle.auction_model(np.ascontiguousarray(compact.T), cap, tail_t=tail)
    assert tot == want
    assert np.array_equal(np.bincount(po, minlength=n_obj), cap)
    soff = np.concatenate([[0], np.cumsum(cap)])
    for o in range(n_obj):
        assert sorted(so[soff[o]:soff[o + 1]].tolist()) == sorted(np.nonzero(po == o)[0].tolist())
    # eps-CS certificate with eps = 1 in units of 1/(n+1)
    h = (compact.T.astype(object) - int(compact.min())) * (n + 1) + np.array([int(x) for x in lam], dtype=object)[None, :]
    assert max(h[i, po[i]] - min(h[i]) for i in range(n)) <= 1


def test_jv_row_map_equals_materialised_expansion():
    rng = np.random.default_rng(3)
    compact = rng.integers(-1000, 1000, (7, 21), dtype=np.int32)
    row_map = np.repeat(np.arange(7, dtype=np.int32), 3)
    r1, c1, (t1, u1, v1) = oracle.lapjv_i32(compact, row_map)
    r2, c2, (t2, _, _) = oracle.lapjv_i32(np.ascontiguousarray(compact[row_map]))
    assert t1 == t2 and np.array_equal(r1, r2)
    check_solution(compact, r1, c1, t1, u1, v1, row_map)
    assert oracle.assignment_cost_i32(compact, r1, row_map) == t1
    assert oracle.min_reduced_cost_i32(compact, u1, v1, row_map) == 0


def test_oracle_rejects_non_square():
    with pytest.raises(ValueError):
        oracle.lapjv_i32(np.zeros((3, 4), np.int32))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cost_oracle_matches_reference_outputs(cost_golden, tag):
    """Golden = outputs of the reference's own normalize_data / matrix_correlation_pearson /
    calculate_cost (tests/golden/make_golden.py)."""
    g = cost_golden
    sc_n = co.normalize_data(g[f"{tag}_sc"]); st_n = co.normalize_data(g[f"{tag}_st"])
    assert np.array_equal(sc_n, g[f"{tag}_sc_norm"]) and np.array_equal(st_n, g[f"{tag}_st_norm"])
    corr = co.matrix_correlation_pearson(sc_n, st_n)
    np.testing.assert_allclose(corr, g[f"{tag}_corr"], rtol=0, atol=1e-12)
    dist_rep, loc_rep = co.calculate_cost(sc_n, st_n, g[f"{tag}_cn"])
    np.testing.assert_allclose(dist_rep, g[f"{tag}_distance_repeat"], rtol=0, atol=1e-12)
    assert np.array_equal(loc_rep, g[f"{tag}_location_repeat"])
    # equals numpy's corrcoef block
    n_spots = st_n.shape[1]
    full = np.corrcoef(np.concatenate([st_n, sc_n], axis=1).T)[:n_spots, n_spots:]
    np.testing.assert_allclose(corr, full, atol=1e-12)


def test_cost_oracle_errors_and_degenerate():
    with pytest.raises(ValueError):
        co.matrix_correlation_pearson(np.zeros((3, 2)), np.zeros((4, 2)))
    x = np.zeros((5, 3)); x[:, 1] = [1, 2, 3, 4, 5]
    n = co.normalize_data(x)
    assert np.all(n[:, 0] == 0) and np.all(np.isfinite(n))       # all-zero column -> zeros (common.py:143-146)
    r = co.matrix_correlation_pearson(n, n)
    assert np.isnan(r[0, 1]) and abs(r[1, 1] - 1) < 1e-12         # sigma == 0 -> NaN (common.py:196-197)
