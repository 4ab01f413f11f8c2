"""GPU: the whole path behind the reference-facing entry points."""
import numpy as np
import pandas as pd
import pytest
import torch

import oracle
from oracle import cost_oracle as co
import cytospace_b200
from cytospace_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def oracle_total_on(cost_i32, row_map=None):
    return oracle.lapjv_i32(cost_i32, row_map)[2][0]


def test_solve_linear_assignment_problem_cfg1(engine):
    """cfg1: 1k x 1k x 2k genes.  LAP parity is defined on the GPU-built integer matrix:
    the device total must equal the CPU JV oracle's total on that same matrix."""
    sc, st, cn = syn.structured_counts(1000, 1000, 2000, 1, seed=1001)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    mapped, pidx = cytospace_b200.solve_linear_assignment_problem(
        sc_n, st_n, cn, "lapjv_b200", None, 1, "Pearson_correlation", process_idx=3)
    assert pidx == 3 and isinstance(mapped, list) and len(mapped) == 1000
    assert sorted(mapped) == list(range(1000))                     # cn == 1: a permutation of the spots
    spot_of_cell, res, cost = engine.assign(sc_n, st_n, cn)
    assert spot_of_cell.cpu().tolist() == mapped                   # deterministic
    cost_np = np.ascontiguousarray(cost[:, :1000].cpu().numpy())        # cn == 1: spots x cells on the device
    assert res.total == oracle_total_on(cost_np)
    # against the float64 reference formulation: the GPU assignment's float64 cost is within
    # n * 4e-6 of the float64 optimum (quantisation + fp16x3 tolerance)
    f64 = -co.matrix_correlation_pearson(sc_n, st_n)
    tot_f64_gpu = f64[np.array(mapped), np.arange(1000)].sum()
    tot_f64_opt = oracle.lapjv_f64(f64)[2][0]
    assert tot_f64_gpu >= tot_f64_opt - 1e-9 and tot_f64_gpu - tot_f64_opt <= 1000 * 4e-6


def test_visium_like_repeated_spots(engine):
    sc, st, cn = syn.structured_counts(900, 150, 800, 6, seed=1004)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    mapped, _ = cytospace_b200.solve_linear_assignment_problem(
        sc_n, st_n, cn, "lapjv", None, 1, "Pearson_correlation")
    assert np.array_equal(np.bincount(mapped, minlength=150), cn)   # every spot gets exactly cn cells
    _, res, cost = engine.assign(sc_n, st_n, cn)
    row_map = np.repeat(np.arange(150, dtype=np.int32), 6)
    assert res.total == oracle_total_on(np.ascontiguousarray(cost[:, :150].T.cpu().numpy()), row_map)


def test_uneven_capacities_and_empty_spots(engine):
    rng = np.random.default_rng(0)
    cn = rng.integers(0, 5, 80)
    sc, st, _ = syn.structured_counts(int(cn.sum()), 80, 400, cn, seed=12)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    mapped, _ = cytospace_b200.solve_linear_assignment_problem(
        sc_n, st_n, cn, "lapjv_compat", None, 1, "Pearson_correlation")
    assert np.array_equal(np.bincount(mapped, minlength=80), cn)


def test_non_square_raises(engine):
    sc, st, cn = syn.structured_counts(50, 40, 100, 1, seed=1)
    with pytest.raises(ValueError, match="square"):
        cytospace_b200.solve_linear_assignment_problem(co.normalize_data(sc), co.normalize_data(st), cn,
                                                       "lapjv", None, 1, "Pearson_correlation")


def test_lapjv_callable_conventions(engine):
    """P2: unmodified call_solver convention on a float64 host matrix (LAS:34-40)."""
    from cytospace_b200 import import_solver, call_solver
    rng = np.random.default_rng(4)
    cost = rng.random((120, 120)) * 2 - 1
    row_ind, col_ind, (total, u, v) = import_solver("lapjv")(cost)
    q = np.rint(cost * 1e6).astype(np.int32)
    assert int(q[np.arange(120), row_ind].sum()) == oracle_total_on(q)
    assert np.array_equal(col_ind[row_ind], np.arange(120))
    assert abs(total - cost[np.arange(120), row_ind].sum()) < 1e-9
    red = cost - u[:, None] - v[None, :]
    assert red.min() > -2e-6                                        # duals feasible up to quantisation
    y = call_solver(import_solver("lapjv"), "lapjv", cost)
    assert np.array_equal(y, col_ind)
    tot2, x2, y2 = import_solver("lapjv_compat")(cost)
    assert np.array_equal(call_solver(import_solver("lapjv_compat"), "lapjv_compat", cost), y2)
    assert np.array_equal(x2, row_ind)
    with pytest.raises(ValueError):
        import_solver("lapjv")(np.zeros((3, 4)))


def test_calculate_cost_compat(engine, cost_golden):
    g = cost_golden
    d, loc = cytospace_b200.calculate_cost(g["b_sc_norm"], g["b_st_norm"], g["b_cn"], "lapjv", "Pearson_correlation")
    assert np.array_equal(loc, g["b_location_repeat"])
    assert np.abs(d - g["b_distance_repeat"]).max() <= 5e-6


def test_apply_linear_assignment_chunked_single_cell(engine):
    """--single-cell -noss: matched blocks of spots and cells (cytospace.py:628-633)."""
    from cytospace_b200 import partition_indices, apply_linear_assignment
    n = 700
    sc, st, cn = syn.structured_counts(n, n, 600, 1, seed=77)
    sc_df = pd.DataFrame(sc, columns=[f"cell{i}" for i in range(n)])
    st_df = pd.DataFrame(st, columns=[f"spot{i}" for i in range(n)])
    coords = pd.DataFrame({"row": np.arange(n), "col": np.arange(n) * 2}, index=st_df.columns)
    np.random.seed(1)
    isc = partition_indices(np.arange(n), split_by_interval_int=300, shuffle=True)
    ist = partition_indices(np.arange(n), split_by_interval_int=300, shuffle=True)
    locs, cells = apply_linear_assignment(sc_df, st_df, coords, cn, "lapjv_b200", None, 1, "Pearson_correlation", 4,
                                          isc, index_st_list=ist)
    assert len(locs) == n and len(cells) == n
    assert sorted(locs.index.tolist()) == sorted(st_df.columns.tolist())     # each spot used once
    assert sorted(cells.tolist()) == sorted(sc_df.columns.tolist())
    # chunk k's cells only land on chunk k's spots
    pos = 0
    for a, b in zip(isc, ist):
        got = set(locs.index[pos:pos + len(a)]); pos += len(a)
        assert got == set(st_df.columns[b])


def test_apply_linear_assignment_sub_spots(engine):
    """--sampling-sub-spots: blocks of cells against all spots with sub-sampled capacities (:650-660)."""
    from cytospace_b200 import partition_indices, apply_linear_assignment
    sc, st, cn = syn.structured_counts(480, 80, 500, 6, seed=31)
    sc_df = pd.DataFrame(sc, columns=[f"cell{i}" for i in range(480)])
    st_df = pd.DataFrame(st, columns=[f"spot{i}" for i in range(80)])
    coords = pd.DataFrame({"row": np.arange(80)}, index=st_df.columns)
    np.random.seed(2)
    isc = partition_indices(np.arange(480), split_by_interval_int=200, shuffle=True)
    agg = np.repeat(np.arange(80), cn)
    parts = partition_indices(agg, split_by_interval_int=200, shuffle=True)
    subs = [np.bincount(p, minlength=80) for p in parts]
    locs, cells = apply_linear_assignment(sc_df, st_df, coords, cn, "lapjv", None, 1, "Pearson_correlation", 2,
                                          isc, subsampled_cell_number_to_node_assignment_list=subs)
    counts = locs.index.value_counts()
    assert all(counts[f"spot{s}"] == cn[s] for s in range(80))
    assert sorted(cells.tolist()) == sorted(sc_df.columns.tolist())


def test_more_spots_than_cells_and_integer_counts(engine):
    """--sampling-sub-spots hands every chunk a bincount(minlength=n_spots) capacity vector (cytospace.py:650-660):
    most spots take no cell and there can be more spots than cells in the chunk.  Count matrices read with
    read_csv are int64.  (both used to raise.)"""
    rng = np.random.default_rng(4)
    n_spots, n_cells = 300, 120
    sc, st, _ = syn.structured_counts(n_cells, n_spots, 500, 1, seed=31)
    cn = np.bincount(rng.integers(0, n_spots, n_cells), minlength=n_spots)
    assert (cn == 0).sum() > n_spots - n_cells - 1
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    mapped, _ = cytospace_b200.solve_linear_assignment_problem(sc_n, st_n, cn, "lapjv_b200", None, 1, "Pearson_correlation")
    assert np.array_equal(np.bincount(mapped, minlength=n_spots), cn)
    keep = np.flatnonzero(cn > 0)
    want_cost = co.cost_matrix_i32(sc_n, st_n[:, keep])
    row_map = np.repeat(np.arange(keep.size, dtype=np.int32), cn[keep])
    _, res, cost = engine.assign(sc_n, st_n, cn)
    assert res.total == oracle_total_on(np.ascontiguousarray(cost[:, :keep.size].T.cpu().numpy()), row_map)
    assert abs(res.total - oracle_total_on(want_cost, row_map)) <= 4 * n_cells
    # integer count matrices through the DataFrame entry point
    sc_i, st_i = sc.astype(np.int64), st.astype(np.int64)
    cells = [f"c{i}" for i in range(n_cells)]
    sc_df = pd.DataFrame(sc_i, columns=cells); st_df = pd.DataFrame(st_i, columns=[f"s{i}" for i in range(n_spots)])
    coords = pd.DataFrame({"row": np.arange(n_spots), "col": np.arange(n_spots)}, index=st_df.columns)
    loc_i, ids_i = cytospace_b200.apply_linear_assignment(sc_df, st_df, coords, cn, "lapjv_b200", None, 1,
                                                          "Pearson_correlation", 1, [np.arange(n_cells)])
    loc_f, ids_f = cytospace_b200.apply_linear_assignment(sc_df.astype(np.float64), st_df.astype(np.float64), coords, cn,
                                                          "lapjv_b200", None, 1, "Pearson_correlation", 1, [np.arange(n_cells)])
    assert loc_i.index.tolist() == loc_f.index.tolist() and list(ids_i) == list(ids_f)


def test_staged_upload_equals_plain_copy(engine):
    """Host arrays above STAGE_MIN_BYTES travel through the pinned ring.  With an explicit dtype (or
    ``stage_float32`` off) the device copy is byte-identical; by default a float64 array lands as float32,
    narrowed by the staging threads exactly like ``astype(np.float32)``, and the engine records whether every
    value survived."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((3000, 5001))                      # 120 MB, not a multiple of the piece size
    assert x.nbytes > engine.STAGE_MIN_BYTES
    d = engine.to_device(x, torch.float64)
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), torch.from_numpy(x))
    d32 = engine.to_device(x)
    assert d32.dtype == torch.float32 and torch.equal(d32.cpu(), torch.from_numpy(x.astype(np.float32)))
    assert engine.last_stage_exact is False
    engine.stage_float32 = False
    try:
        assert torch.equal(engine.to_device(x).cpu(), torch.from_numpy(x))
    finally:
        del engine.stage_float32                               # back to the class default
    xi = rng.integers(0, 50, (3000, 3000))                     # int64 counts -> float64 on the host -> exact float32 on the wire
    di = engine.to_device(xi)
    assert di.dtype == torch.float32 and torch.equal(di.cpu().double(), torch.from_numpy(xi.astype(np.float64)))
    assert engine.last_stage_exact is True
    xn = x.copy(); xn[5, 7] = np.nan; xn[9, 1] = np.inf       # NaN / inf pass through (NaN counts as inexact)
    dn = engine.to_device(xn).cpu().numpy()
    assert np.isnan(dn[5, 7]) and np.isinf(dn[9, 1])


def test_float32_staging_moves_the_cost_by_at_most_one_unit(engine):
    """The plugin call uploads large float64 matrices as float32 (half the PCIe bytes).  Tolerance, stated here: on
    log2(TPM+1) data the integer cost changes by at most 1 unit (1e-6 on r) in well under 1 % of the entries, and
    stays within the cost-build tolerance of the float64 oracle."""
    sc, st, cn = syn.structured_counts(900, 900, 5000, 1, seed=77)          # 36 MB each: above STAGE_MIN_BYTES
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    assert sc_n.nbytes > engine.STAGE_MIN_BYTES and sc_n.dtype == np.float64
    c32 = engine.cost_build(engine.to_device(sc_n), engine.to_device(st_n))[:, :900].cpu().numpy()
    assert engine.last_stage_exact is False
    c64 = engine.cost_build(engine.to_device(sc_n, torch.float64), engine.to_device(st_n, torch.float64))[:, :900].cpu().numpy()
    diff = np.abs(c32.astype(np.int64) - c64)
    assert diff.max() <= 1 and (diff != 0).mean() < 0.01
    ref = co.cost_matrix_i32(sc_n, st_n)                                    # spots x cells, float64 oracle
    assert np.abs(c32.T.astype(np.int64) - ref).max() <= 3


@pytest.mark.gpu
def test_stage_upload_ring_wraps_and_odd_sizes():
    """cyb_stage_upload with a ring much smaller than the array (3 pieces of 1 MB, 4 workers): every slot is reused
    many times, the last piece is ragged, sizes that are not a multiple of the 64-byte copy unit, back-to-back calls
    on the same stream.  (The ring is sized once per process, hence the child process.)"""
    import subprocess, sys, os
    child = r"""
import numpy as np, torch
from cytospace_b200 import _native
lib = _native.load(); ffi = _native.ffi()
rng = np.random.default_rng(5)
s = torch.cuda.current_stream().cuda_stream
for nbytes in (1, 63, 64, 65, (1 << 20) - 1, (1 << 20) + 1, 37 * (1 << 20) + 5):
    x = rng.integers(0, 256, nbytes + 3, dtype=np.uint8)[3:]                 # deliberately misaligned source
    outs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    for o in outs:                                                           # two calls in flight behind each other
        _native.check(lib.cyb_stage_upload(ffi.cast("const void *", x.ctypes.data), ffi.cast("void *", o.data_ptr()), nbytes, ffi.cast("void *", s)))
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), x), nbytes
assert lib.cyb_stage_upload(ffi.NULL, ffi.NULL, 8, ffi.cast("void *", s)) != 0
print("ok")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CYB_STAGE_THREADS="4", CYB_STAGE_PIECE_MB="1", CYB_STAGE_PIECES="3", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
