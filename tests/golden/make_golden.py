"""Generates the committed golden fixtures under tests/golden/ (run in the BUILD container, where
/root/reference exists; the GPU box never runs this).

1. cost_golden.npz -- outputs of the REFERENCE's own code (imported from /root/reference with the
   four absent third-party modules stubbed): normalize_data (common.py:142-147),
   matrix_correlation_pearson (common.py:190-199), calculate_cost
   (linear_assignment_solvers.py:42-69) and partition_indices (cytospace.py:150-209) on small
   seeded inputs.
2. lap_golden.npz  -- small integer LAP instances with their optimal total from
   scipy.optimize.linear_sum_assignment (independent implementation) and the permutation of the
   repo's JV restatement (oracle/lapjv_oracle.c) so the restatement itself cannot drift.
   The reference ships no LAP vectors and its lapjv wheel is absent: PARITY UNPINNED vs the wheel.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


class _Stub(types.ModuleType):
    """Absent third-party module: any attribute is a dummy class (never called on this path)."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return type(item, (), {})


def import_reference():
    for name in ("scanpy", "datatable", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.patches", "matplotlib.lines", "matplotlib.collections", "matplotlib.cm",
                 "ortools", "ortools.graph", "ortools.graph.pywrapgraph", "seaborn"):
        m = _Stub(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    sys.modules["ortools.graph"].pywrapgraph = sys.modules["ortools.graph.pywrapgraph"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, "/root/reference")
    from cytospace.common.common import normalize_data, matrix_correlation_pearson
    from cytospace.linear_assignment_solvers.linear_assignment_solvers import calculate_cost
    from cytospace.cytospace import partition_indices
    return normalize_data, matrix_correlation_pearson, calculate_cost, partition_indices


def main():
    from cytospace_b200 import synthetic as syn
    import oracle
    from scipy.optimize import linear_sum_assignment

    normalize_data, pearson, calculate_cost, partition_indices = import_reference()
    out = {}
    # --- cost build: three shapes incl. repeated spots and a ragged gene count
    cases = [("a", 48, 48, 200, 1), ("b", 60, 20, 333, 3), ("c", 37, 37, 64, 1)]
    for tag, n_cells, n_spots, n_genes, cps in cases:
        sc, st, cn = syn.structured_counts(n_cells, n_spots, n_genes, cps, seed=100 + len(tag) + n_genes)
        sc_n, st_n = normalize_data(sc.copy()), normalize_data(st.copy())
        corr = pearson(sc_n, st_n)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            dist_rep, loc_rep = calculate_cost(sc_n, st_n, cn, "lapjv", "Pearson_correlation")
        out[f"{tag}_sc"] = sc; out[f"{tag}_st"] = st; out[f"{tag}_cn"] = cn
        out[f"{tag}_sc_norm"] = sc_n; out[f"{tag}_st_norm"] = st_n
        out[f"{tag}_corr"] = corr; out[f"{tag}_distance_repeat"] = dist_rep; out[f"{tag}_location_repeat"] = loc_rep
    # --- partition_indices (no shuffle, and shuffle under a fixed global seed)
    p1 = partition_indices(np.arange(1800), split_by_category_list=np.array([500, 1000, 300]),
                           split_by_interval_int=400, shuffle=False)
    out["part1_lens"] = np.array([len(p) for p in p1]); out["part1_cat"] = np.concatenate(p1)
    p2 = partition_indices(np.arange(2500), split_by_interval_int=1000, shuffle=False)
    out["part2_lens"] = np.array([len(p) for p in p2])
    np.random.seed(7)
    p3 = partition_indices(np.arange(103), split_by_interval_int=25, shuffle=True)
    out["part3_lens"] = np.array([len(p) for p in p3]); out["part3_cat"] = np.concatenate(p3)
    p4 = partition_indices(np.arange(10), shuffle=False)
    out["part4_lens"] = np.array([len(p) for p in p4])
    np.savez_compressed(os.path.join(HERE, "cost_golden.npz"), **out)

    # --- LAP instances
    rng = np.random.default_rng(2024)
    lap = {}
    mats = {
        "uniform16": rng.integers(0, 100, (16, 16)),
        "uniform64": rng.integers(0, 2_000_000, (64, 64)),
        "negative33": rng.integers(-1_000_000, 1_000_000, (33, 33)),
        "ties24": rng.integers(0, 3, (24, 24)),
        "constant9": np.full((9, 9), 7),
        "one": np.array([[5]]),
        "two": np.array([[4, 1], [2, 8]]),
        "duprows30": np.repeat(rng.integers(-500_000, 500_000, (6, 30)), 5, axis=0),
        "dupcols28": np.repeat(rng.integers(0, 1000, (28, 7)), 4, axis=1),
        "diag40": (np.ones((40, 40)) * 1000 - np.eye(40) * 999).astype(np.int64),
    }
    a = out["a_distance_repeat"]
    mats["pearson48"] = np.rint(a * 1e6)
    for name, m in mats.items():
        m = np.ascontiguousarray(m, dtype=np.int32)
        ri, ci = linear_sum_assignment(m.astype(np.float64))
        opt = int(m[ri, ci].astype(np.int64).sum())
        rowsol, colsol, (tot, u, v) = oracle.lapjv_i32(m)
        assert tot == opt, (name, tot, opt)
        lap[f"{name}_cost"] = m; lap[f"{name}_opt"] = np.int64(opt)
        lap[f"{name}_rowsol"] = rowsol; lap[f"{name}_colsol"] = colsol
    np.savez_compressed(os.path.join(HERE, "lap_golden.npz"), **lap)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
