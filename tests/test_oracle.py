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


@settings(max_examples=80, deadline=None)
@given(st.integers(1, 10), st.integers(0, 2 ** 31 - 1), st.sampled_from([2, 10, 1000, 2_000_000]),
       st.sampled_from([(1, 1, 1), (4, 3, 1), (4, 8, 4), (64, 100000, 32), (2, 2, 32)]), st.sampled_from([2, 4, 64]))
def test_sap_finish_model_equals_expanded_jv(n_obj, seed, high, knobs, theta):
    """The hybrid the device runs (auction rounds + shortest-augmenting-path finish, oracle/sap_model.c) reaches
    the optimum of the reference's expanded LAP (LAS:63-66) for every search schedule (switch point, rows per
    round, paths per search), including empty spots, and leaves an eps = 1 certificate."""
    sap_t, K, multi = knobs
    rng = np.random.default_rng(seed)
    cap = rng.integers(0, 5, n_obj).astype(np.int32)
    if cap.sum() == 0:
        cap[0] = 1
    n = int(cap.sum())
    compact = rng.integers(-high, high, (n_obj, n), dtype=np.int32)           # spots x cells (reference orientation)
    row_map = np.repeat(np.arange(n_obj, dtype=np.int32), cap)
    want = oracle.lapjv_i32(compact, row_map)[2][0]
    # partial > 0: phases with eps > 1 stop early and hand their free persons to the next phase
    po, so, tot, lam, stats, _ = oracle.sap_model(np.ascontiguousarray(compact.T), cap, theta=theta, sap_t=sap_t, K=K, multi=multi,
                                                  warm=seed & 1, partial=(0, 2, 64)[(seed >> 1) % 3], chain=(seed >> 3) % 3)
    assert tot == want
    assert np.array_equal(np.bincount(po, minlength=n_obj), cap)
    soff = np.concatenate([[0], np.cumsum(cap)])
    for o in range(n_obj):
        assert sorted(so[soff[o]:soff[o + 1]].tolist()) == sorted(np.nonzero(po == o)[0].tolist())
    h = (compact.T.astype(object) - int(compact.min())) * (n + 1) + np.array([int(x) for x in lam], dtype=object)[None, :]
    assert max(h[i, po[i]] - min(h[i]) for i in range(n)) <= 1


@pytest.mark.parametrize("name", LAP_NAMES)
def test_sap_finish_model_on_golden_instances(lap_golden, name):
    cost = lap_golden[f"{name}_cost"]
    opt = int(lap_golden[f"{name}_opt"])
    for kw in (dict(), dict(sap_t=4, K=8, multi=1, theta=4), dict(sap_t=0),       # sap_t = 0: pure auction
               dict(theta=4, sap_t=64, K=296, multi=16, partial=64, warm=1),
               dict(theta=8, sap_t=64, K=296, multi=16, partial=64, warm=1, chain=2), dict(sap_t=4, K=8, multi=1, theta=4, chain=1)):
        po, so, tot, lam, stats, _ = oracle.sap_model(cost, **kw)
        assert tot == opt and sorted(po.tolist()) == list(range(cost.shape[0]))
        assert np.array_equal(so[po], np.arange(cost.shape[0]))


def test_sap_finish_needs_far_fewer_dependent_steps_than_the_auction_tail():
    """The measurement behind the device design: on a structured instance the Gauss-Seidel
    auction tail makes thousands of dependent single bids; the search finish replaces them by a few hundred
    grid-wide rounds."""
    from cytospace_b200 import synthetic as syn
    from oracle import cost_oracle as co
    sc, st_, cn = syn.structured_counts(1500, 1500, 1500, 1, seed=1002)
    cost = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st_))
    a = oracle.auction_model(cost, tail_t=8)
    s = oracle.sap_model(cost, theta=64, sap_t=148, K=296, multi=32)
    assert a[2] == s[2] == oracle.lapjv_i32(cost)[2][0]
    auction_dependent = int(a[4][1]) + int(a[4][5])           # grid rounds + tail bids
    sap_dependent = int(s[4][1]) + int(s[4][4])               # grid rounds + search rounds
    assert sap_dependent * 4 < auction_dependent, (sap_dependent, auction_dependent)


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


# ---- Spearman / Euclidean / lap_CSPR -------------------------------------------

@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("metric,key", [("Pearson_correlation", "pearson"), ("Spearman_correlation", "spearman"),
                                        ("Euclidean", "euclid")])
def test_metric_oracle_matches_reference_calculate_cost(cost_golden, tag, metric, key):
    """Golden = the reference's calculate_cost for every --distance-metric (LAS:46-59)."""
    g = cost_golden
    dist_rep, loc_rep = co.calculate_cost(g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"], g[f"{tag}_cn"], metric)
    np.testing.assert_allclose(dist_rep, g[f"{tag}_{key}_distance_repeat"], rtol=0, atol=1e-10)
    assert np.array_equal(loc_rep, g[f"{tag}_location_repeat"])


def test_average_ranks_restatement_equals_pandas():
    import pandas as pd
    rng = np.random.default_rng(5)
    x = rng.poisson(0.4, (300, 17)).astype(np.float64)         # heavy ties
    x[:, 3] = 2.5                                              # constant column
    x[::7, 5] = -0.0
    assert np.array_equal(co.average_ranks(x), pd.DataFrame(x).rank().values)
    y = rng.normal(size=(64, 5))
    assert np.array_equal(co.average_ranks(y), pd.DataFrame(y).rank().values)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("key,metric", [("pearson", "Pearson_correlation"), ("spearman", "Spearman_correlation"),
                                        ("euclid", "Euclidean")])
def test_oracle_solve_equals_reference_solve_end_to_end(cost_golden, tag, key, metric):
    """Golden `mapped` = the reference's own solve_linear_assignment_problem (CYT:304-351) with SciPy
    behind the lapjv convention.  The restated float64 JV on the reference's formulation
    (distance_repeat + 1e-16 * rand, CYT:325-327) reaches the same total cost."""
    g = cost_golden
    dist_rep, loc_rep = co.calculate_cost(g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"], g[f"{tag}_cn"], metric)
    n = dist_rep.shape[0]
    np.random.seed(1)
    cost_scaled = dist_rep + 1e-16 * np.random.rand(n, n)
    _, colsol, (total, _, _) = oracle.lapjv_f64(cost_scaled)
    mapped = loc_rep[colsol]
    ref = g[f"{tag}_{key}_mapped"]
    compact = co.metric_cost(g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"], metric)          # spots x cells
    tot_ours = compact[mapped, np.arange(n)].sum(); tot_ref = compact[ref, np.arange(n)].sum()
    assert abs(tot_ours - tot_ref) <= 1e-9 * max(1.0, abs(tot_ref))
    assert np.array_equal(np.bincount(mapped, minlength=len(g[f"{tag}_cn"])), g[f"{tag}_cn"])
    assert np.array_equal(np.bincount(ref, minlength=len(g[f"{tag}_cn"])), g[f"{tag}_cn"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cspr_matrix_restatement_equals_reference(cost_golden, tag):
    """Golden `cspr_int` = the arcs the reference's match_solution handed to (a recording fake of)
    ortools for seed 1: int(1e6 * d + 10 * rand + 1) transposed (CYT:335-340)."""
    g = cost_golden
    want = g[f"{tag}_cspr_int"]
    got = co.cspr_matrix_reference(g[f"{tag}_pearson_distance_repeat"], 1)
    assert got.shape == want.shape
    skipped = want == 2 ** 40                                   # arcs of cost 0 are skipped (LAS:79)
    assert np.array_equal(got[~skipped], want[~skipped]) and np.all(got[skipped] == 0)
    # JV on this integer matrix reproduces the reference's lap_CSPR assignment cost
    n = want.shape[0]
    rowsol, colsol, (total, _, _) = oracle.lapjv_i32(np.ascontiguousarray(got.astype(np.int32)))
    loc_rep = g[f"{tag}_location_repeat"]
    ref_mapped = g[f"{tag}_pearson_mapped_cspr"]
    assert np.array_equal(np.bincount(loc_rep[rowsol], minlength=len(g[f"{tag}_cn"])), g[f"{tag}_cn"])
    compact = co.metric_cost(g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"])
    mine = compact[loc_rep[rowsol], np.arange(n)].sum(); ref = compact[ref_mapped, np.arange(n)].sum()
    assert abs(mine - ref) <= n * 11e-6                          # both optimal up to the [1, 11) noise


def test_hash_noise_known_answers_and_range():
    """Known answers computed with an independent C transcription of the splitmix64 hash."""
    assert co.hash_noise(1, 3, 5).tolist() == [[4, 6, 8, 7, 9], [6, 5, 9, 9, 2], [10, 6, 10, 4, 9]]
    z = co.hash_noise(12345, 200, 300)
    assert z.min() == 1 and z.max() == 10
    assert abs(z.mean() - 5.5) < 0.05
    assert np.array_equal(co.hash_noise(12345, 200, 300), z)
    assert not np.array_equal(co.hash_noise(12346, 200, 300), z)
