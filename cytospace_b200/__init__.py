"""cytospace_b200 -- B200-native (sm_100a) implementation of CytoSPACE's cell-to-spot
assignment hot path: Pearson-correlation cost build + exact dense linear assignment.

Importing the package does not need a GPU; every compute entry does (no CPU fallback).

Layout
------
csrc/                        CUDA kernels + the C ABI (include/cytospace_b200.h)
_native.py                   cffi (ABI mode) loader of libcytospace_b200.so
engine.py                    device driver: cost_build / lap_solve / assign
linear_assignment_solvers.py mirror of the reference module (import_solver, call_solver, calculate_cost)
cytospace.py                 mirror of solve_linear_assignment_problem / partition_indices / apply_linear_assignment
lapjv.py                     lapjv.lapjv / lap.lapjv compatible callables
chunking.py                  chunk planner (single GPU or one rank per GPU)
synthetic.py                 synthetic N-cell x S-spot x G-gene generators
"""
from .cytospace import apply_linear_assignment, partition_indices, solve_linear_assignment_problem
from .linear_assignment_solvers import SOLVER_METHODS, calculate_cost, call_solver, import_solver

__all__ = ["apply_linear_assignment", "partition_indices", "solve_linear_assignment_problem",
           "SOLVER_METHODS", "calculate_cost", "call_solver", "import_solver"]
__version__ = "0.1.0"
