"""Drop-in callables with the call/return conventions of the two third-party solvers
CytoSPACE can import (linear_assignment_solvers.py:11-31), served by the B200 LAP kernel.

* ``lapjv(cost)``        -- convention of ``lapjv==1.3.14``: returns
  ``(row_ind, col_ind, (total_cost, u, v))``; CytoSPACE consumes ``[1]`` (:38).
* ``lapjv_compat(cost)`` -- convention of ``lap==0.4.0``: returns ``(total_cost, x, y)``;
  CytoSPACE consumes ``[2]`` (:36).

Entry P2 of SURVEY section 8(b): the float64 host matrix an unmodified CytoSPACE built
(``distance_repeat + 1e-16 * rand``, cytospace.py:326-327) is copied to the device, integerised
as ``rint(scale * cost)`` (scale 1e6 for |cost| <= 1000, precedent cytospace.py:337) and solved
exactly on that integer matrix.  Documented deviation: tie-noise below 1/scale is not
representable, ties are broken by index (lowest column for a row, lowest row for a column).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import AssignmentEngine, COST_SCALE

_engine = None


def _get_engine() -> AssignmentEngine:
    global _engine
    if _engine is None:
        _engine = AssignmentEngine()
    return _engine


def _solve(cost_matrix):
    cost = np.asarray(cost_matrix)
    if cost.ndim != 2 or cost.shape[0] != cost.shape[1]:
        raise ValueError("\"cost_matrix\" must be a square 2D numpy array")
    if cost.shape[0] == 0:
        raise ValueError("\"cost_matrix\" is empty")
    eng = _get_engine()
    c64 = np.ascontiguousarray(cost, dtype=np.float64)
    amax = float(np.max(np.abs(c64))) if c64.size else 0.0
    if not np.isfinite(amax):
        raise ValueError("cost matrix contains NaN or inf")
    scale = float(COST_SCALE) if amax <= 1000.0 else float(2 ** 29) / amax
    dev = eng.quantise(eng.to_device(c64, torch.float64), scale)       # (explicit dtype: a cost matrix is never narrowed)
    res = eng.lap_solve(dev, n_persons=cost.shape[0], n_objects=cost.shape[0])    # persons = rows, objects = columns
    rowsol = res.person_obj.cpu().numpy()
    colsol = res.slot_owner.cpu().numpy()
    n = cost.shape[0]
    # duals in the caller's units: v_j = -price_j / ((n+1) * scale), u_i = c[i, x_i] - v[x_i]
    v = -(res.price.cpu().numpy().astype(np.float64)) / ((n + 1) * scale)
    picked = c64[np.arange(n), rowsol]
    u = picked - v[rowsol]
    return rowsol, colsol, float(picked.sum()), u, v


def lapjv(cost_matrix, verbose=0, force_doubles=False):
    """``lapjv.lapjv`` convention."""
    rowsol, colsol, total, u, v = _solve(cost_matrix)
    return rowsol, colsol, (total, u, v)


def lapjv_compat(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
    """``lap.lapjv`` convention."""
    if extend_cost or np.isfinite(cost_limit):
        raise NotImplementedError("extend_cost / cost_limit are not used by CytoSPACE and not supported")
    rowsol, colsol, total, _, _ = _solve(cost)
    return (total, rowsol, colsol) if return_cost else (rowsol, colsol)
