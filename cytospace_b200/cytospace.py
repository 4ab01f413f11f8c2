"""Host-side mirror of the assignment path of ``cytospace/cytospace.py``: the functions a
maintainer re-points at this package to make ``--solver-method lapjv_b200`` a drop-in
(INTEGRATION.md).  Same names, argument meaning, return values and exceptions as the
reference; the arithmetic runs in the sm_100a kernels behind the C ABI.

* ``solve_linear_assignment_problem``  <- cytospace/cytospace.py:304-351
* ``partition_indices``                <- cytospace/cytospace.py:150-209
* ``apply_linear_assignment``          <- cytospace/cytospace.py:354-469 (the process pool of
  :430-467 becomes the GPU chunk planner of ``chunking.py``: CUDA cannot live behind fork)
"""
from __future__ import annotations

import time

import numpy as np

from . import chunking
from .linear_assignment_solvers import DISTANCE_METRICS, SOLVER_METHODS, get_engine

_GPU_METHODS = SOLVER_METHODS


def solve_linear_assignment_problem(scRNA_norm_data, st_norm_data, cell_number_to_node_assignment,
                                    solver_method, solver, seed, distance_metric, process_idx=None):
    """cytospace.py:304-351 for ``solver_method`` in {lapjv, lapjv_compat, lapjv_b200, lap_CSPR} and
    ``distance_metric`` in {Pearson_correlation, Spearman_correlation, Euclidean}.

    Parameters as the reference: normalised genes x cells arrays, cell count per spot, the
    solver name, the solver callable (ignored: cost build and solve are fused on the device),
    ``seed`` (accepted for signature parity; the integer path breaks ties by index instead of
    the reference's 1e-16 * U(0,1) noise, cytospace.py:325-327; the lap_CSPR path uses it to seed
    its integer tie noise, cytospace.py:336-337, drawn from a counter-based hash instead of
    MT19937), the distance metric and a ``process_idx`` returned as is.
    Returns ``(mapped_st_index: List[int] of length n_cells, process_idx)`` -- element c is the
    column index in ``st_norm_data`` of the spot that cell c is mapped to."""
    if solver_method not in _GPU_METHODS:
        raise ValueError("Invalid solver_method provided")
    if distance_metric not in DISTANCE_METRICS:
        raise ValueError(f"Invalid distance_metric provided: {distance_metric}")
    eng = get_engine()
    t0 = time.perf_counter()
    spot_of_cell, res, _ = eng.assign(np.asarray(scRNA_norm_data), np.asarray(st_norm_data),
                                      cell_number_to_node_assignment, metric=distance_metric,
                                      cspr_seed=(seed if solver_method == "lap_CSPR" else None), progress=print)
    mapped_st_index = spot_of_cell.cpu().numpy().tolist()
    print(f"Time to build cost matrix and solve linear assignment problem: "
          f"{round(time.perf_counter() - t0, 2)} seconds")
    return mapped_st_index, process_idx


def partition_indices(indices, split_by_category_list=None, split_by_interval_int=None, shuffle=True):
    """cytospace.py:150-209 -- split ``indices`` into consecutive blocks.

    Breakpoints: the category boundaries ``cumsum(split_by_category_list)`` (blocks never straddle
    a category), and inside every block longer than ``split_by_interval_int`` one breakpoint every
    ``split_by_interval_int`` entries.  ``shuffle`` permutes ``indices`` in place first with the
    global NumPy RNG, as the reference does (:181-182)."""
    indices = np.asarray(indices) if not isinstance(indices, np.ndarray) else indices
    total = len(indices)
    if shuffle:
        np.random.shuffle(indices)
    cuts = {0, total}
    if split_by_category_list is not None:
        if np.sum(split_by_category_list) != total:
            print('Warning: sum of counts in each category does not match the full length')
        cuts.update(int(c) for c in np.cumsum(split_by_category_list))
    ordered = sorted(cuts)
    if split_by_interval_int is not None:
        step = int(split_by_interval_int)
        for lo, hi in zip(ordered[:-1], ordered[1:]):
            if hi - lo > step:
                cuts.update(range(lo, hi, step))
    inner = [c for c in sorted(cuts) if 0 < c < total]
    return np.array_split(indices, inner)


def apply_linear_assignment(scRNA_data, st_data, coordinates_data, cell_number_to_node_assignment,
                            solver_method, solver, seed, distance_metric, number_of_processors,
                            index_sc_list, index_st_list=None,
                            subsampled_cell_number_to_node_assignment_list=None):
    """cytospace.py:354-469 with the process pool replaced by the GPU chunk planner.

    ``scRNA_data`` / ``st_data`` / ``coordinates_data`` are pandas DataFrames formatted as the
    reference's ``read_data`` returns them (genes x cells, genes x spots, spots x coords), RAW
    (un-normalised): ``normalize_data`` (common.py:142-147, called at cytospace.py:398-399) is
    fused into the device standardise pre-pass.  ``number_of_processors`` is accepted for
    signature parity (chunks run back to back on this rank's GPU, or are spread over the ranks of
    an initialised ``torch.distributed`` group).  Returns ``(assigned_locations DataFrame,
    cell_ids_selected ndarray)`` with rows paired as in the reference; chunks are concatenated in
    chunk order (the reference's completion order, :453-467, is non-deterministic)."""
    import pandas as pd

    if (index_st_list is not None) and (subsampled_cell_number_to_node_assignment_list is not None):
        raise ValueError("index_st_list and subsampled_cell_number_to_node_assignment_list cannot both be specified")
    if solver_method not in _GPU_METHODS:
        raise ValueError("Invalid solver_method provided")
    if distance_metric not in DISTANCE_METRICS:
        raise ValueError(f"Invalid distance_metric provided: {distance_metric}")
    sc_np = scRNA_data.to_numpy()
    st_np = st_data.to_numpy()
    cn = np.asarray(cell_number_to_node_assignment)
    plan = chunking.plan_chunks(sc_np.shape[1], st_np.shape[1], cn, index_sc_list, index_st_list,
                                subsampled_cell_number_to_node_assignment_list)
    if len(plan) > 1:
        print(f"Number of required processors: {len(plan)}")
    assign_kw = {}
    if distance_metric != "Pearson_correlation":
        assign_kw["metric"] = distance_metric
    if solver_method == "lap_CSPR":
        assign_kw["cspr_seed"] = seed
    mapped = chunking.solve_chunks(get_engine(), sc_np, st_np, plan, log_tpm=True, **assign_kw)
    locations, cell_ids = [], []
    for chunk, mapped_st_index in zip(plan, mapped):
        coords = coordinates_data if chunk.st_index is None else coordinates_data.iloc[chunk.st_index]
        locations.append(coords.iloc[mapped_st_index])
        cell_ids.append(scRNA_data.columns.values[chunk.sc_index])
    if len(plan) == 1:
        return locations[0], cell_ids[0]
    return pd.concat(locations), np.concatenate(cell_ids, axis=0)
