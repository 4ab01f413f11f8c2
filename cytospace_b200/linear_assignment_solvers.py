"""Host-side mirror of ``cytospace/linear_assignment_solvers/linear_assignment_solvers.py``
(every solver method and distance metric): same function names, argument meaning and error behaviour,
with the arithmetic done by the sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .engine import AssignmentEngine, COST_SCALE

#: ``--solver-method`` values served here.  ``lapjv`` / ``lapjv_compat`` are the reference's own
#: names (argument_parser.py:69-71); ``lapjv_b200`` is the new value a maintainer adds to that
#: ``choices`` list to route cost build + solve through this package.
SOLVER_METHODS = ("lapjv", "lapjv_compat", "lapjv_b200", "lap_CSPR")
DISTANCE_METRICS = ("Pearson_correlation", "Spearman_correlation", "Euclidean")

_engine = None


def get_engine() -> AssignmentEngine:
    global _engine
    if _engine is None:
        _engine = AssignmentEngine()
    return _engine


def import_solver(solver_method):
    """linear_assignment_solvers.py:11-31 -- returns the solver callable for ``solver_method``."""
    if solver_method == "lapjv_compat":
        from .lapjv import lapjv_compat
        return lapjv_compat
    if solver_method in ("lapjv", "lapjv_b200"):
        from .lapjv import lapjv
        return lapjv
    raise NotImplementedError(f"The solver {solver_method} is not a supported solver "
                              "for the shortest augmenting path method, choose between "
                              "'lapjv' and 'lapjv_compat'.")


def call_solver(solver, solver_method, cost_scaled):
    """linear_assignment_solvers.py:34-40 -- the row (slot) assigned to each column (cell)."""
    if solver_method == "lapjv_compat":
        _, _, y = solver(cost_scaled)
    elif solver_method in ("lapjv", "lapjv_b200"):
        _, y, _ = solver(cost_scaled)
    else:
        raise ValueError("Invalid solver_method provided")
    return y


def calculate_cost(expressions_tpm_scRNA_log, expressions_tpm_st_log, cell_number_to_node_assignment,
                   solver_method, distance_metric):
    """linear_assignment_solvers.py:42-69, all three metrics (:46-59) and the slot expansion (:63-66).

    Compatibility form: returns ``(distance_repeat float64 [n x N], location_repeat int [n])`` on
    the host like the reference, where ``distance_repeat`` is the device-built integer matrix
    divided by the integer scale (so it is quantised to 1e-6).  ``solver_method`` only selects the
    reference's branch (:46 vs :53); both build the same matrix.  The fast path
    (``cytospace.solve_linear_assignment_problem``) never materialises this expansion."""
    if distance_metric not in DISTANCE_METRICS:
        raise ValueError(f"Invalid distance_metric provided: {distance_metric}")
    print("Building cost matrix ...")
    t0 = time.perf_counter()
    eng = get_engine()
    sc = eng.to_device(np.asarray(expressions_tpm_scRNA_log, dtype=np.float64))
    st = eng.to_device(np.asarray(expressions_tpm_st_log, dtype=np.float64))
    cost_i32 = eng.cost_build(sc, st, metric=distance_metric)                      # cells x spots on the device
    n_spots = st.shape[1]
    cost = cost_i32[:, :n_spots].T.cpu().numpy().astype(np.float64) / COST_SCALE    # spots x cells like the reference
    location_repeat = np.repeat(np.arange(len(cell_number_to_node_assignment)),
                                cell_number_to_node_assignment).astype(int)
    distance_repeat = cost[location_repeat, :]
    print(f"Time to build cost matrix: {round(time.perf_counter() - t0, 2)} seconds")
    return distance_repeat, location_repeat
