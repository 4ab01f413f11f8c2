"""GPU: the LAP kernel (through the C ABI) against the CPU oracle.  Bit-exact bar: identical total
cost; the permutation must be a valid optimal assignment (certificate: eps-complementary
slackness <= 1 scaled unit, checked on the device and again on the host)."""
import numpy as np
import pytest
import torch

import oracle
from cytospace_b200 import synthetic as syn
from conftest import LAP_NAMES

pytestmark = pytest.mark.gpu


def solve_and_check(engine, cost_np, row_map_np=None, grid=0):
    n = cost_np.shape[1] if row_map_np is None else len(row_map_np)
    ld = (cost_np.shape[1] + 31) // 32 * 32
    dev = torch.full((cost_np.shape[0], ld), 2 ** 30 - 1, dtype=torch.int32, device=engine.device)
    dev[:, :cost_np.shape[1]] = torch.from_numpy(cost_np).to(engine.device)
    rm = None if row_map_np is None else torch.from_numpy(row_map_np.astype(np.int32)).to(engine.device)
    res = engine.lap_solve(dev, rm, n=n, grid=grid)
    rowsol = res.rowsol.cpu().numpy(); colsol = res.colsol.cpu().numpy()
    assert sorted(rowsol.tolist()) == list(range(n)), "rowsol is not a permutation"
    assert np.array_equal(colsol[rowsol], np.arange(n)), "colsol is not the inverse of rowsol"
    assert oracle.assignment_cost_i32(cost_np, rowsol, row_map_np) == res.total
    cert = engine.lap_check(dev, res, rm)
    assert cert["invalid_rows"] == 0 and cert["total"] == res.total
    assert cert["max_violation"] <= 1, cert
    return res, rowsol, colsol


@pytest.mark.parametrize("name", LAP_NAMES)
def test_golden_instances(engine, lap_golden, name):
    cost = lap_golden[f"{name}_cost"]
    res, rowsol, _ = solve_and_check(engine, cost)
    assert res.total == int(lap_golden[f"{name}_opt"])


@pytest.mark.parametrize("n,high,seed", [(1, 10, 0), (2, 10, 1), (3, 5, 2), (31, 100, 3), (33, 2_000_000, 4),
                                          (127, 50, 5), (257, 2_000_000, 6), (1000, 2_000_000, 7),
                                          (1023, 1000, 8), (2050, 2_000_000, 9)])
def test_uniform_random_vs_oracle(engine, n, high, seed):
    cost = np.random.default_rng(seed).integers(-high, high, (n, n), dtype=np.int32)
    res, _, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]


@pytest.mark.parametrize("grid", [1, 2, 7, 148])
def test_result_independent_of_grid_size(engine, grid):
    cost = np.random.default_rng(11).integers(0, 10_000, (300, 300), dtype=np.int32)
    ref, rowsol_ref, _ = solve_and_check(engine, cost, grid=0)
    res, rowsol, _ = solve_and_check(engine, cost, grid=grid)
    assert res.total == ref.total and np.array_equal(rowsol, rowsol_ref)      # deterministic tie-breaks


def test_row_map_expansion_cfg4_like(engine):
    """Visium-like: S spots x cn cells per spot, rows resolved by index (LAS:63-66 never materialised)."""
    from oracle import cost_oracle as co
    sc, st, cn = syn.structured_counts(600, 100, 500, 6, seed=1004)
    compact = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))
    row_map = np.repeat(np.arange(100, dtype=np.int32), 6)
    res, _, _ = solve_and_check(engine, compact, row_map)
    assert res.total == oracle.lapjv_i32(compact, row_map)[2][0]
    assert res.total == oracle.lapjv_i32(np.ascontiguousarray(compact[row_map]))[2][0]


def test_degenerate_inputs(engine):
    for cost in (np.zeros((64, 64), np.int32), np.full((5, 5), -7, np.int32),
                 np.tile(np.arange(40, dtype=np.int32), (40, 1)),            # identical rows
                 np.tile(np.arange(40, dtype=np.int32)[:, None], (1, 40)),   # identical columns
                 np.random.default_rng(1).integers(0, 2, (200, 200)).astype(np.int32)):
        res, _, _ = solve_and_check(engine, cost)
        assert res.total == oracle.lapjv_i32(cost)[2][0]


def test_extreme_cost_range(engine):
    lim = 2 ** 30 - 1
    cost = np.random.default_rng(5).integers(-lim, lim, (96, 96), dtype=np.int64).astype(np.int32)
    res, _, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]


def test_structured_cfg1_vs_oracle(engine):
    """BASELINE config 1 (1k x 1k x 2k genes): LAP on the oracle-built integer matrix."""
    from oracle import cost_oracle as co
    sc, st, cn = syn.structured_counts(1000, 1000, 2000, 1, seed=1001)
    cost = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))
    res, _, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]


def test_large_lap_properties_4k(engine):
    """Size-independent properties at a size the oracle still finishes quickly."""
    cost = syn.uniform_cost_i32(4096, seed=21)
    res, rowsol, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]
    # linearity: adding a row potential / column potential shifts the optimum by a known constant
    a = np.random.default_rng(1).integers(-1000, 1000, 4096).astype(np.int32)
    b = np.random.default_rng(2).integers(-1000, 1000, 4096).astype(np.int32)
    shifted = cost + a[:, None] + b[None, :]
    res2, _, _ = solve_and_check(engine, shifted)
    assert res2.total == res.total + int(a.sum()) + int(b.sum())


def test_lap_errors(engine):
    with pytest.raises(ValueError):
        engine.lap_solve(torch.zeros((4, 32), dtype=torch.float32, device=engine.device))
    with pytest.raises(ValueError):
        engine.lap_solve(torch.zeros((4, 32), dtype=torch.int32, device=engine.device), n=40)
