"""CPU: host-side logic and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import cytospace_b200
from cytospace_b200 import _native, chunking
from cytospace_b200.cytospace import partition_indices


def test_library_loads_and_exports_every_declared_symbol():
    _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _native.declared_symbols()
    assert len(declared) >= 12
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/cytospace_b200.h but not exported"
    assert _native.load().cyb_abi_version() == _native.load().CYB_ABI_VERSION


def test_size_queries_without_gpu():
    lib = _native.load()
    assert lib.cyb_operand_k(20000, lib.CYB_PREC_F16) == 20032
    assert lib.cyb_operand_k(20000, lib.CYB_PREC_F16X3) == 3 * 20032
    assert lib.cyb_operand_k(64, lib.CYB_PREC_F16) == 64
    assert lib.cyb_lap_workspace_bytes(1000, 1000) >= 1000 * 48
    assert lib.cyb_lap_workspace_bytes(0, 0) > 0
    need = lib.cyb_cost_build_workspace_bytes(2000, 1000, 1000, lib.CYB_PREC_F16X3)
    assert need >= 2 * 1000 * 3 * 2048 * 2


def test_native_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is checkable on CPU."""
    lib, ffi = _native.load(), _native.ffi()
    rc = lib.cyb_lap_solve_i32(ffi.NULL, 8, 8, 8, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL,
                               ffi.NULL, 0, 0, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"null" in ffi.string(lib.cyb_last_error())
    rc = lib.cyb_lap_solve_i32(ffi.NULL, 8, 0, 0, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL,
                               ffi.NULL, 0, 0, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID
    rc = lib.cyb_lap_solve_i32(ffi.NULL, 8, 8, 4, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL,
                               ffi.NULL, 0, 0, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"square" in ffi.string(lib.cyb_last_error())
    rc = lib.cyb_lap_solve_i32(ffi.NULL, 1 << 18, 1 << 18, 1 << 18, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL,
                               ffi.NULL, ffi.NULL, 0, 0, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"2^18" in ffi.string(lib.cyb_last_error())     # person index is 18 bits
    rc = lib.cyb_cost_gemm_i32(ffi.cast("void *", 16), ffi.cast("void *", 16), 8, 8, 70, 1.0,
                               ffi.cast("int32_t *", 16), 8, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID           # k not a multiple of 64
    assert lib.cyb_stage_upload(ffi.NULL, ffi.NULL, 0, ffi.NULL) == 0                        # nothing to copy
    rc = lib.cyb_stage_upload(ffi.NULL, ffi.NULL, 64, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"null" in ffi.string(lib.cyb_last_error())
    with pytest.raises(_native.CybError):
        _native.check(rc)


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from cytospace_b200.engine import AssignmentEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AssignmentEngine()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.dirname(cytospace_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_partition_indices_matches_reference_golden(cost_golden):
    g = cost_golden
    p1 = partition_indices(np.arange(1800), split_by_category_list=np.array([500, 1000, 300]),
                           split_by_interval_int=400, shuffle=False)
    assert [len(p) for p in p1] == g["part1_lens"].tolist() == [400, 100, 400, 400, 200, 300]
    assert np.array_equal(np.concatenate(p1), g["part1_cat"])
    p2 = partition_indices(np.arange(2500), split_by_interval_int=1000, shuffle=False)
    assert [len(p) for p in p2] == g["part2_lens"].tolist()
    np.random.seed(7)
    p3 = partition_indices(np.arange(103), split_by_interval_int=25, shuffle=True)
    assert [len(p) for p in p3] == g["part3_lens"].tolist()
    assert np.array_equal(np.concatenate(p3), g["part3_cat"])          # same global-RNG shuffle
    p4 = partition_indices(np.arange(10), shuffle=False)
    assert [len(p) for p in p4] == g["part4_lens"].tolist() == [10]


def test_plan_chunks_modes():
    cn = np.array([2, 0, 3, 1])
    one = chunking.plan_chunks(6, 4, cn, [np.arange(6)])
    assert len(one) == 1 and one[0].st_index is None and one[0].n == 6
    sc_l = [np.array([0, 1, 2]), np.array([3, 4, 5])]
    st_l = [np.array([0, 3]), np.array([2])]
    single = chunking.plan_chunks(6, 4, cn, sc_l, index_st_list=st_l)
    assert [c.cn.tolist() for c in single] == [[2, 1], [3]]
    sub = chunking.plan_chunks(6, 4, cn, sc_l, subsampled_cell_number_to_node_assignment_list=[
        np.array([1, 0, 2, 0]), np.array([1, 0, 1, 1])])
    assert all(c.st_index is None for c in sub) and sub[1].cn.sum() == 3
    with pytest.raises(ValueError):
        chunking.plan_chunks(6, 4, cn, sc_l, index_st_list=st_l,
                             subsampled_cell_number_to_node_assignment_list=[cn, cn])


def test_assign_ranks_balances_and_is_deterministic():
    owner = chunking.assign_ranks([25000] * 8, 8)
    assert sorted(owner) == list(range(8))
    owner = chunking.assign_ranks([25000] * 8, 2)
    assert owner.count(0) == owner.count(1) == 4
    owner = chunking.assign_ranks([10000, 10000, 10000, 3000], 2)
    assert owner == chunking.assign_ranks([10000, 10000, 10000, 3000], 2)
    assert owner[3] == owner[2] or owner.count(owner[3]) == 2      # the small chunk joins the lighter rank


def test_solver_surface_mirrors_reference():
    from cytospace_b200 import linear_assignment_solvers as las
    with pytest.raises(NotImplementedError, match="not a supported solver"):
        las.import_solver("lap_CSPR")                                # LAS:20-22: only the two lapjv names import
    assert callable(las.import_solver("lapjv")) and callable(las.import_solver("lapjv_compat"))
    assert las.call_solver(lambda c: (0, "y_lapjv", 1), "lapjv", None) == "y_lapjv"
    assert las.call_solver(lambda c: (0, 1, "y_lap"), "lapjv_compat", None) == "y_lap"
    from cytospace_b200.cytospace import solve_linear_assignment_problem
    with pytest.raises(ValueError, match="Invalid solver_method"):
        solve_linear_assignment_problem(None, None, None, "bogus", None, 1, "Pearson_correlation")
    with pytest.raises(ValueError, match="Invalid distance_metric"):
        solve_linear_assignment_problem(None, None, None, "lapjv", None, 1, "Manhattan")
    assert set(las.DISTANCE_METRICS) == {"Pearson_correlation", "Spearman_correlation", "Euclidean"}
    assert "lap_CSPR" in las.SOLVER_METHODS


def test_metric_entry_points_validate_arguments_without_gpu():
    lib, ffi = _native.load(), _native.ffi()
    assert lib.cyb_rank_workspace_bytes(20000, 10000) >= 10000 * 8
    assert lib.cyb_rank_workspace_bytes(40000, 100) > lib.cyb_rank_workspace_bytes(20000, 100)   # multi-run scratch
    pe = lib.cyb_cost_build_metric_workspace_bytes(lib.CYB_METRIC_PEARSON, 2000, 1000, 1000, lib.CYB_PREC_F16X3)
    assert pe == lib.cyb_cost_build_workspace_bytes(2000, 1000, 1000, lib.CYB_PREC_F16X3)
    sp = lib.cyb_cost_build_metric_workspace_bytes(lib.CYB_METRIC_SPEARMAN, 2000, 1000, 1000, lib.CYB_PREC_F16X3)
    assert sp >= pe + 2 * 2000 * 1000 * 4                              # float32 rank matrices
    assert lib.cyb_cost_build_metric_workspace_bytes(lib.CYB_METRIC_EUCLIDEAN, 2000, 1000, 1000, lib.CYB_PREC_F16X3) > pe
    rc = lib.cyb_rank_columns(ffi.NULL, lib.CYB_F64, 10, 10, 10, 0, ffi.NULL, 10, ffi.NULL, 0, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"null" in ffi.string(lib.cyb_last_error())
    p16 = ffi.cast("void *", 256)
    rc = lib.cyb_rank_columns(p16, 7, 10, 10, 10, 0, ffi.cast("float *", 256), 10, p16, 1 << 20, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"dtype" in ffi.string(lib.cyb_last_error())
    rc = lib.cyb_cost_build(9, p16, p16, lib.CYB_F64, 10, 10, 10, 10, 10, 0, lib.CYB_PREC_F16, 1e6,
                            ffi.cast("int32_t *", 256), 32, ffi.NULL, p16, 1 << 30, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID and b"metric" in ffi.string(lib.cyb_last_error())
    rc = lib.cyb_expand_rows_noise_i32(ffi.NULL, 8, 8, 8, ffi.NULL, 1, 1, 10, ffi.NULL, 8, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID
    rc = lib.cyb_cost_gemm_euclid_i32(p16, p16, 8, 8, 64, 100, 1.0, ffi.NULL, ffi.NULL, ffi.cast("int32_t *", 256), 8,
                                      ffi.NULL, ffi.NULL)
    assert rc == lib.CYB_ERR_INVALID
