"""GPU: the Spearman / Euclidean metrics and the lap_CSPR formulation against the
oracle and the reference's own outputs (tests/golden/cost_golden.npz)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import cost_oracle as co
import cytospace_b200
from cytospace_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

CORR_TOL = 3          # units of 1e-6 on a correlation (f16x3 operands, fp32 chunked accumulation)
EUCLID_RTOL = 1e-6    # end-to-end budget per cell, relative to the largest distance


def euclid_err_ok(got_i32, sc_n, st_n):
    """The Euclidean cost is the correlation GEMM seen through
    d^2 = G [(mu_a - mu_b)^2 + (sd_a - sd_b)^2 + 2 sd_a sd_b (1 - r)]: its tolerance is the correlation
    tolerance (CORR_TOL * 1e-6 on r) carried to the SQUARED distance, plus one unit of rounding of d."""
    want = co.euclidean_distance(sc_n, st_n)                               # spots x cells
    G = sc_n.shape[0]
    sd_c, sd_s = sc_n.std(0), st_n.std(0)
    bound_d2 = 2.0 * G * np.multiply.outer(sd_s, sd_c) * CORR_TOL * 1e-6
    got = got_i32.astype(np.float64) / 1e6
    slack = 2e-6 * (got + want) + 1e-12                                    # rounding of d to 1e-6, on d^2
    return bool(np.all(np.abs(got * got - want * want) <= bound_d2 + slack))


def dev(engine, x, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype).to(engine.device)


@pytest.mark.parametrize("G,n,kind", [(1, 3, "normal"), (2, 5, "ties"), (5, 4, "ties"), (333, 70, "ties"),
                                      (1025, 33, "normal"), (4096, 64, "counts"), (20000, 48, "counts"),
                                      (30011, 40, "counts"), (60000, 12, "ties"), (40000, 6, "normal")])
def test_rank_columns_equals_pandas_rank(engine, G, n, kind):
    """pd.DataFrame(x).rank() (COM:207-208): bit-exact, incl. columns longer than one shared-memory run."""
    rng = np.random.default_rng(G + n)
    if kind == "normal":
        x = rng.normal(size=(G, n))
    elif kind == "ties":
        x = rng.integers(-2, 3, (G, n)).astype(np.float64)
        x[:, 0] = 7.0
        x[::3, 1] = -0.0
    else:
        x = co.normalize_data(rng.poisson(0.2, (G, n)).astype(np.float64) + (rng.random((G, n)) < 0.01) * 50)
    want = co.average_ranks(x)
    got = engine.rank_columns(dev(engine, x)).cpu().numpy()
    assert np.array_equal(got.astype(np.float64), want)
    got32 = engine.rank_columns(dev(engine, x, torch.float32)).cpu().numpy()
    assert np.array_equal(got32.astype(np.float64), co.average_ranks(x.astype(np.float32)))


def test_rank_columns_fused_normalize_and_strided(engine):
    rng = np.random.default_rng(9)
    raw = rng.poisson(0.5, (700, 90)).astype(np.float64)
    raw[:, 4] = 0                                            # all-zero column: normalize_data -> zeros
    want = co.average_ranks(co.normalize_data(raw))
    big = dev(engine, np.concatenate([raw, raw], axis=1))
    got = engine.rank_columns(big[:, :90], log_tpm=True).cpu().numpy()      # ld_x = 180
    assert np.array_equal(got.astype(np.float64), want)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("metric,key", [("Spearman_correlation", "spearman"), ("Euclidean", "euclid")])
def test_metric_cost_matches_reference_golden(engine, cost_golden, tag, metric, key):
    g = cost_golden
    sc, st = dev(engine, g[f"{tag}_sc_norm"]), dev(engine, g[f"{tag}_st_norm"])
    S, N = st.shape[1], sc.shape[1]
    want = co.metric_cost(g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"], metric)          # [S x N], pinned on CPU
    lr = g[f"{tag}_location_repeat"]
    np.testing.assert_allclose(want[lr], g[f"{tag}_{key}_distance_repeat"], rtol=0, atol=1e-10)
    got = engine.cost_build(sc, st, layout="spots_x_cells", metric=metric)[:, :N].cpu().numpy()
    gotT = engine.cost_build(sc, st, layout="cells_x_spots", metric=metric)[:, :S].cpu().numpy()
    if metric == "Euclidean":
        assert euclid_err_ok(got, g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"])
        assert euclid_err_ok(np.ascontiguousarray(gotT.T), g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"])
    else:
        assert np.abs(got - np.rint(want * 1e6)).max() <= CORR_TOL
        assert np.abs(gotT.T - np.rint(want * 1e6)).max() <= CORR_TOL     # operands swapped: same tolerance


@pytest.mark.parametrize("metric", ["Spearman_correlation", "Euclidean"])
def test_metric_cost_structured_2k_genes(engine, metric):
    sc, st, cn = syn.structured_counts(384, 200, 2000, 1, seed=21)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    want = co.metric_cost(sc_n, st_n, metric)
    got = engine.cost_build(dev(engine, sc_n), dev(engine, st_n), layout="spots_x_cells", metric=metric)[:, :384]
    got = got.cpu().numpy()
    if metric == "Euclidean":
        assert euclid_err_ok(got, sc_n, st_n)
        assert np.abs(got - want * 1e6).max() <= 4 * EUCLID_RTOL * 1e6 * want.max()       # ~1 ppm on real-sized distances
    else:
        assert np.abs(got - np.rint(want * 1e6)).max() <= CORR_TOL
    # raw counts + fused normalize_data give the same matrix
    got2 = engine.cost_build(dev(engine, sc), dev(engine, st), log_tpm=True, layout="spots_x_cells", metric=metric)
    assert np.abs(got2[:, :384].cpu().numpy() - got).max() <= 1


def test_spearman_constant_column_is_rejected(engine):
    sc, st, cn = syn.structured_counts(40, 40, 200, 1, seed=3)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    sc_n[:, 5] = 1.25                                       # constant column: all ranks equal, sigma = 0
    with pytest.raises(ValueError, match="zero variance"):
        engine.cost_build(dev(engine, sc_n), dev(engine, st_n), metric="Spearman_correlation")


def test_euclidean_constant_columns_and_overflow(engine):
    rng = np.random.default_rng(2)
    x = rng.random((70, 40)); x[:, 3] = 1.0; x[:, 9] = 2.0            # constant columns are fine for a distance
    got = engine.cost_build(dev(engine, x), dev(engine, x), metric="Euclidean", layout="spots_x_cells")[:, :40]
    assert euclid_err_ok(got.cpu().numpy(), x, x)           # incl. the zero diagonal (sqrt amplifies: bound is on d^2)
    y = x * 1e3                                                        # 1e6 * distance no longer fits int32
    with pytest.raises(ValueError, match="cannot be represented"):
        engine.cost_build(dev(engine, y), dev(engine, y), metric="Euclidean")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("metric,key", [("Pearson_correlation", "pearson"), ("Spearman_correlation", "spearman"),
                                        ("Euclidean", "euclid")])
@pytest.mark.parametrize("solver_method", ["lapjv", "lap_CSPR"])
def test_entry_point_against_reference_end_to_end(cost_golden, tag, metric, key, solver_method):
    """Golden `mapped` = what the reference's own solve_linear_assignment_problem (CYT:304-351)
    returns on these inputs (SciPy standing in for the absent solver wheels).  The assignment is not
    unique under near-ties, its float64 cost is: ours must match within the quantisation budget
    (1e-6 per cell; + the [1, 11) integer noise per cell for lap_CSPR, CYT:337)."""
    g = cost_golden
    sc_n, st_n, cn = g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"], g[f"{tag}_cn"]
    n = sc_n.shape[1]
    mapped, pidx = cytospace_b200.solve_linear_assignment_problem(sc_n, st_n, cn, solver_method, None, 1, metric, 5)
    assert pidx == 5 and len(mapped) == n
    assert np.array_equal(np.bincount(mapped, minlength=len(cn)), cn)
    ref = g[f"{tag}_{key}_mapped" + ("_cspr" if solver_method == "lap_CSPR" else "")]
    compact = co.metric_cost(sc_n, st_n, metric)                       # spots x cells float64
    ours = compact[np.asarray(mapped), np.arange(n)].sum(); theirs = compact[ref, np.arange(n)].sum()
    per_cell = (EUCLID_RTOL * compact.max() + 2e-6) if metric == "Euclidean" else 4e-6
    if solver_method == "lap_CSPR":
        per_cell += 11e-6
    assert abs(ours - theirs) <= n * per_cell, (ours, theirs)


def test_cspr_matrix_and_total_bit_exact(engine):
    """The expanded lap_CSPR matrix equals the oracle's (same hash noise) entry for entry, and the
    device LAP total equals JV on that matrix."""
    sc, st, cn = syn.structured_counts(300, 60, 500, 5, seed=77)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    spot_of_cell, res, cost = engine.assign(sc_n, st_n, cn, cspr_seed=42)
    compact = engine.cost_build(dev(engine, sc_n), dev(engine, st_n), layout="spots_x_cells")[:, :300].cpu().numpy()
    want = co.cspr_matrix_i32(compact, cn, 42)
    got = cost[:, :300].cpu().numpy()
    assert np.array_equal(got, want)
    assert res.total == oracle.lapjv_i32(np.ascontiguousarray(want))[2][0]
    assert np.array_equal(np.bincount(spot_of_cell.cpu().numpy(), minlength=60), cn)
    # a different seed changes the noise, not the feasibility
    spot2, res2, _ = engine.assign(sc_n, st_n, cn, cspr_seed=43)
    assert np.array_equal(np.bincount(spot2.cpu().numpy(), minlength=60), cn)


def test_apply_linear_assignment_other_metric_chunked(engine):
    import pandas as pd
    sc, st, cn = syn.structured_counts(240, 240, 300, 1, seed=31)
    genes = [f"g{i}" for i in range(300)]
    sc_df = pd.DataFrame(sc, index=genes, columns=[f"c{i}" for i in range(240)])
    st_df = pd.DataFrame(st, index=genes, columns=[f"s{i}" for i in range(240)])
    coords = pd.DataFrame({"row": np.arange(240), "col": np.arange(240) * 2}, index=st_df.columns)
    idx_sc = [np.arange(0, 120), np.arange(120, 240)]
    idx_st = [np.arange(0, 120), np.arange(120, 240)]
    loc, ids = cytospace_b200.apply_linear_assignment(sc_df, st_df, coords, cn, "lapjv", None, 1,
                                                      "Spearman_correlation", 1, idx_sc, index_st_list=idx_st)
    assert len(loc) == 240 and len(ids) == 240
    assert sorted(loc.index.tolist()) == sorted(st_df.columns.tolist())       # every spot used once
    first = set(loc.index[:120])
    assert first == set(st_df.columns[:120])                                   # chunks stay matched
