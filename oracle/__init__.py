"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CytoSPACE assignment hot path.

Nothing under ``cytospace_b200/`` may import this package.  Allowed users:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs -- as the checker or the reported CPU baseline only.

Contents
--------
* ``lapjv_i32`` / ``lapjv_f64`` -- ctypes bindings of ``lapjv_oracle.c``, a
  restatement of the dense Jonker-Volgenant LAP that the reference reaches via
  ``lapjv.lapjv`` (``/root/reference/cytospace/linear_assignment_solvers/
  linear_assignment_solvers.py:34-40``).  PARITY UNPINNED against the wheel
  (absent here); pinned against SciPy and committed golden vectors instead.
* ``cost_oracle`` -- numpy float64 restatement of the cost build
  (``matrix_correlation_pearson`` ``cytospace/common/common.py:190-199``,
  ``calculate_cost`` ``linear_assignment_solvers.py:42-69``,
  ``normalize_data`` ``common.py:142-147``); pinned against the reference's
  own functions imported with stubs (``tests/golden/make_golden.py``).
* ``sap_model`` -- sequential model of the device solver (``sap_model.c``: auction rounds, incomplete
  phases, warm-started shortest-augmenting-path searches); the GPU tests require identical assignments,
  prices and counters.  ``auction_model`` is the round-1 model (pure auction), kept as the comparator of
  the experiment behind the redesign.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -O3)."""
    srcs = [os.path.join(_HERE, f) for f in ("lapjv_oracle.c", "lapjv_body.inc", "auction_model.c", "sap_model.c", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        lib.lapjv_i32_solve.restype = ctypes.c_int
        lib.lapjv_i32_solve.argtypes = [ctypes.c_int, i32p, ctypes.c_int64, i32p, i32p, i32p, i64p, i64p, i64p, i64p]
        lib.lapjv_f64_solve.restype = ctypes.c_int
        lib.lapjv_f64_solve.argtypes = [ctypes.c_int, f64p, ctypes.c_int64, i32p, i32p, i32p, f64p, f64p, f64p, i64p]
        lib.lapjv_i32_assignment_cost.restype = ctypes.c_int64
        lib.lapjv_i32_assignment_cost.argtypes = [ctypes.c_int, i32p, ctypes.c_int64, i32p, i32p]
        lib.lapjv_i32_min_reduced_cost.restype = ctypes.c_int64
        lib.lapjv_i32_min_reduced_cost.argtypes = [ctypes.c_int, i32p, ctypes.c_int64, i32p, i64p, i64p]
        lib.auction_model_i32.restype = ctypes.c_int
        lib.auction_model_i32.argtypes = [ctypes.c_int, ctypes.c_int, i32p, ctypes.c_int64, i32p, i32p, i32p, i64p, i64p,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, i64p, i32p, ctypes.c_int64, ctypes.c_int]
        lib.sap_model_i32.restype = ctypes.c_int
        lib.sap_model_i32.argtypes = [ctypes.c_int, ctypes.c_int, i32p, ctypes.c_int64, i32p, i32p, i32p, i64p, i64p,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, i64p]
        _lib = lib
    return _lib


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def _prep(cost, row_map, dtype):
    cost = np.ascontiguousarray(cost, dtype=dtype)
    if cost.ndim != 2:
        raise ValueError("cost must be 2-D")
    if row_map is not None:
        row_map = np.ascontiguousarray(row_map, dtype=np.int32)
        n = int(row_map.shape[0])
        if row_map.size and (row_map.min() < 0 or row_map.max() >= cost.shape[0]):
            raise ValueError("row_map out of range")
    else:
        n = int(cost.shape[0])
    if cost.shape[1] != n:
        raise ValueError(f"LAP must be square: {n} rows vs {cost.shape[1]} columns")
    return cost, row_map, n


def lapjv_i32(cost, row_map=None, return_stats=False):
    """Dense JV on an int32 matrix.  Returns (row_ind, col_ind, (total, u, v)),
    the ``lapjv.lapjv`` return convention; CytoSPACE uses [1]."""
    lib = _load()
    cost, row_map, n = _prep(cost, row_map, np.int32)
    rowsol = np.empty(n, np.int32); colsol = np.empty(n, np.int32)
    u = np.zeros(n, np.int64); v = np.zeros(n, np.int64)
    total = np.zeros(1, np.int64); stats = np.zeros(4, np.int64)
    rc = lib.lapjv_i32_solve(n, _ptr(cost, ctypes.c_int32), cost.shape[1], _ptr(row_map, ctypes.c_int32),
                             _ptr(rowsol, ctypes.c_int32), _ptr(colsol, ctypes.c_int32),
                             _ptr(u, ctypes.c_int64), _ptr(v, ctypes.c_int64), _ptr(total, ctypes.c_int64),
                             _ptr(stats, ctypes.c_int64))
    if rc != 0:
        raise RuntimeError(f"lapjv_i32_solve failed rc={rc}")
    out = (rowsol, colsol, (int(total[0]), u, v))
    return out + (stats,) if return_stats else out


def lapjv_f64(cost, row_map=None):
    lib = _load()
    cost, row_map, n = _prep(cost, row_map, np.float64)
    rowsol = np.empty(n, np.int32); colsol = np.empty(n, np.int32)
    u = np.zeros(n, np.float64); v = np.zeros(n, np.float64)
    total = np.zeros(1, np.float64); stats = np.zeros(4, np.int64)
    rc = lib.lapjv_f64_solve(n, _ptr(cost, ctypes.c_double), cost.shape[1], _ptr(row_map, ctypes.c_int32),
                             _ptr(rowsol, ctypes.c_int32), _ptr(colsol, ctypes.c_int32),
                             _ptr(u, ctypes.c_double), _ptr(v, ctypes.c_double), _ptr(total, ctypes.c_double),
                             _ptr(stats, ctypes.c_int64))
    if rc != 0:
        raise RuntimeError(f"lapjv_f64_solve failed rc={rc}")
    return rowsol, colsol, (float(total[0]), u, v)


def assignment_cost_i32(cost, rowsol, row_map=None) -> int:
    lib = _load()
    cost, row_map, n = _prep(cost, row_map, np.int32)
    rowsol = np.ascontiguousarray(rowsol, dtype=np.int32)
    return int(lib.lapjv_i32_assignment_cost(n, _ptr(cost, ctypes.c_int32), cost.shape[1],
                                             _ptr(row_map, ctypes.c_int32), _ptr(rowsol, ctypes.c_int32)))


def min_reduced_cost_i32(cost, u, v, row_map=None) -> int:
    lib = _load()
    cost, row_map, n = _prep(cost, row_map, np.int32)
    u = np.ascontiguousarray(u, dtype=np.int64); v = np.ascontiguousarray(v, dtype=np.int64)
    return int(lib.lapjv_i32_min_reduced_cost(n, _ptr(cost, ctypes.c_int32), cost.shape[1],
                                              _ptr(row_map, ctypes.c_int32), _ptr(u, ctypes.c_int64),
                                              _ptr(v, ctypes.c_int64)))


def auction_model(m, cap=None, theta=4, eps0_div=4, tail_t=0, round_cap=0, variant=1, early_stop=0):
    """Sequential model of the device auction.  ``m`` is persons x objects (cells x spots: the
    TRANSPOSE of the reference's cost), ``cap`` the object capacities (None: all 1).
    Returns (person_obj, slot_owner, total, lambda, stats, round_log)."""
    lib = _load()
    m = np.ascontiguousarray(m, dtype=np.int32)
    P, O = m.shape
    capa = None if cap is None else np.ascontiguousarray(cap, dtype=np.int32)
    if (O if capa is None else int(capa.sum())) != P:
        raise ValueError("capacities must sum to the number of persons")
    person_obj = np.empty(P, np.int32); slot_owner = np.empty(P, np.int32)
    lam = np.zeros(O, np.int64); total = np.zeros(1, np.int64); stats = np.zeros(6, np.int64)
    rlog = np.zeros(max(round_cap, 1), np.int32)
    ctypes.c_int.in_dll(lib, "auction_early_stop").value = int(early_stop)     # device knob CYB_LAP_EARLY (default 0)
    rc = lib.auction_model_i32(P, O, _ptr(m, ctypes.c_int32), m.shape[1], _ptr(capa, ctypes.c_int32),
                               _ptr(person_obj, ctypes.c_int32), _ptr(slot_owner, ctypes.c_int32),
                               _ptr(lam, ctypes.c_int64), _ptr(total, ctypes.c_int64),
                               theta, eps0_div, tail_t, _ptr(stats, ctypes.c_int64),
                               _ptr(rlog, ctypes.c_int32), round_cap, variant)
    if rc != 0:
        raise RuntimeError(f"auction_model_i32 failed rc={rc}")
    return person_obj, slot_owner, int(total[0]), lam, stats, rlog[:min(round_cap, int(stats[1]))]


def sap_model(m, cap=None, theta=4, eps0_div=4, sap_t=64, K=148, multi=0, warm=0, partial=0, chain=0):
    """Sequential model of the hybrid device solver (auction rounds + shortest-augmenting-path finish,
    ``sap_model.c``).  Same orientation as ``auction_model``.  Returns (person_obj, slot_owner, total,
    lambda, stats, phase_log) -- stats / phase_log columns are documented in sap_model.c."""
    lib = _load()
    m = np.ascontiguousarray(m, dtype=np.int32)
    P, O = m.shape
    capa = None if cap is None else np.ascontiguousarray(cap, dtype=np.int32)
    if (O if capa is None else int(capa.sum())) != P:
        raise ValueError("capacities must sum to the number of persons")
    person_obj = np.empty(P, np.int32); slot_owner = np.empty(P, np.int32)
    lam = np.zeros(O, np.int64); total = np.zeros(1, np.int64); stats = np.zeros(8, np.int64)
    ctypes.c_int.in_dll(lib, "sap_warm").value = int(warm)
    ctypes.c_int.in_dll(lib, "sap_partial").value = int(partial)
    ctypes.c_int.in_dll(lib, "sap_chain").value = int(chain)
    rc = lib.sap_model_i32(P, O, _ptr(m, ctypes.c_int32), m.shape[1], _ptr(capa, ctypes.c_int32),
                           _ptr(person_obj, ctypes.c_int32), _ptr(slot_owner, ctypes.c_int32),
                           _ptr(lam, ctypes.c_int64), _ptr(total, ctypes.c_int64),
                           theta, eps0_div, sap_t, K, multi, _ptr(stats, ctypes.c_int64))
    if rc != 0:
        raise RuntimeError(f"sap_model_i32 failed rc={rc}")
    log = np.ctypeslib.as_array((ctypes.c_int64 * 8 * 64).in_dll(lib, "sap_phase_log")).copy()
    return person_obj, slot_owner, int(total[0]), lam, stats, log[:int(stats[0])]
