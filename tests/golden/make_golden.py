"""Generates the committed golden fixtures under tests/golden/ (run in the BUILD container, where
/root/reference exists; the GPU box never runs this).

1. cost_golden.npz -- outputs of the REFERENCE's own code (imported from /root/reference with the
   four absent third-party modules stubbed): normalize_data (common.py:142-147),
   matrix_correlation_pearson (common.py:190-199), calculate_cost
   (linear_assignment_solvers.py:42-69) and partition_indices (cytospace.py:150-209) on small
   seeded inputs; the Spearman / Euclidean branches of calculate_cost (common.py:202-215,
   linear_assignment_solvers.py:51,59); and the reference's own solve_linear_assignment_problem
   (cytospace.py:304-351) run END TO END for every metric and for both solver branches, with the
   absent third-party solvers replaced by SciPy's linear_sum_assignment behind their call
   conventions (a `lapjv.lapjv`-shaped callable; a recording fake of
   ortools.graph.pywrapgraph.LinearSumAssignment) -> mapped_st_index vectors.
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


def scipy_lapjv(cost):
    """A callable with the `lapjv.lapjv` return convention backed by SciPy."""
    from scipy.optimize import linear_sum_assignment
    ri, ci = linear_sum_assignment(cost)
    row_ind = np.empty(len(ri), np.int32); col_ind = np.empty(len(ri), np.int32)
    row_ind[ri] = ci; col_ind[ci] = ri
    return row_ind, col_ind, (float(cost[ri, ci].sum()), None, None)


class FakeLinearSumAssignment:
    """Records the arcs match_solution adds (linear_assignment_solvers.py:72-96) and solves them with
    SciPy; a skipped arc (integer cost exactly 0, :79) is forbidden, as in ortools."""
    OPTIMAL, INFEASIBLE, POSSIBLE_OVERFLOW = 0, 1, 2
    last = None

    def __init__(self):
        self.arcs = {}
        FakeLinearSumAssignment.last = self

    def AddArcWithCost(self, worker, task, cost):
        self.arcs[(worker, task)] = cost

    def Solve(self):
        from scipy.optimize import linear_sum_assignment
        n = 1 + max(max(w for w, _ in self.arcs), max(t for _, t in self.arcs))
        big = 2 ** 40
        m = np.full((n, n), big, dtype=np.int64)
        for (w, t), c in self.arcs.items():
            m[w, t] = c
        ri, ci = linear_sum_assignment(m)
        self.n, self.mate, self.m = n, ci, m
        return self.OPTIMAL if (m[ri, ci] < big).all() else self.INFEASIBLE

    def NumNodes(self):
        return self.n

    def RightMate(self, i):
        return int(self.mate[i])

    def AssignmentCost(self, i):
        return int(self.m[i, self.mate[i]])

    def OptimalCost(self):
        return int(sum(self.AssignmentCost(i) for i in range(self.n)))


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
        # the other metrics and the reference's own solve, end to end
        import cytospace.linear_assignment_solvers.linear_assignment_solvers as las
        from cytospace.cytospace import solve_linear_assignment_problem
        las.pywrapgraph.LinearSumAssignment = FakeLinearSumAssignment
        for metric, key in (("Pearson_correlation", "pearson"), ("Spearman_correlation", "spearman"),
                            ("Euclidean", "euclid")):
            with contextlib.redirect_stdout(io.StringIO()):
                d_rep, _ = calculate_cost(sc_n, st_n, cn, "lapjv", metric)
                d_cspr, _ = calculate_cost(sc_n, st_n, cn, "lap_CSPR", metric)
                mapped, _ = solve_linear_assignment_problem(sc_n, st_n, cn, "lapjv", scipy_lapjv, 1, metric)
                mapped_cspr, _ = solve_linear_assignment_problem(sc_n, st_n, cn, "lap_CSPR", None, 1, metric)
            assert np.array_equal(d_rep, d_cspr)          # both branches build the same matrix
            out[f"{tag}_{key}_distance_repeat"] = d_rep
            out[f"{tag}_{key}_mapped"] = np.asarray(mapped, dtype=np.int64)
            out[f"{tag}_{key}_mapped_cspr"] = np.asarray(mapped_cspr, dtype=np.int64)
            if key == "pearson":      # the integer matrix the reference handed to the solver (cells x slots)
                out[f"{tag}_cspr_int"] = FakeLinearSumAssignment.last.m.copy()
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
