"""CPU: the N>1 chunk distribution (broadcast / point-to-point / all-gather) with gloo, world_size 2.
The device engine is replaced by a stub that solves with the oracle -- this tests the plumbing,
not the kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """Test double with the AssignmentEngine.assign contract, CPU tensors, oracle arithmetic."""
    device = torch.device("cpu")

    def to_device(self, x, dtype=None):
        t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
        if dtype is None and t.dtype not in (torch.float32, torch.float64):
            dtype = torch.float64
        return t if dtype is None else t.to(dtype)

    def assign(self, sc, st, cn, log_tpm=False, metric="Pearson_correlation", cspr_seed=None):
        import oracle
        from oracle import cost_oracle as co
        sc = sc.numpy() if torch.is_tensor(sc) else np.asarray(sc)
        st = st.numpy() if torch.is_tensor(st) else np.asarray(st)
        if log_tpm:
            sc, st = co.normalize_data(sc), co.normalize_data(st)
        cost = co.cost_matrix_i32(sc, st, metric)
        row_map = np.repeat(np.arange(len(cn), dtype=np.int32), np.asarray(cn))
        if cspr_seed is not None:
            cost, row_map_solve = co.cspr_matrix_i32(cost, cn, cspr_seed), None
        else:
            row_map_solve = row_map
        _, colsol, _ = oracle.lapjv_i32(cost, row_map_solve)
        return torch.from_numpy(row_map[colsol].astype(np.int64)), None, None


def _problem(mode):
    from cytospace_b200 import synthetic as syn
    from cytospace_b200.cytospace import partition_indices
    from cytospace_b200 import chunking
    if mode == "single_cell":
        sc, st, cn = syn.structured_counts(90, 90, 150, 1, seed=5)
        isc = partition_indices(np.arange(90), split_by_interval_int=40, shuffle=False)
        ist = partition_indices(np.arange(90), split_by_interval_int=40, shuffle=False)
        plan = chunking.plan_chunks(90, 90, cn, isc, index_st_list=ist)
    else:
        sc, st, cn = syn.structured_counts(60, 20, 150, 3, seed=6)
        isc = partition_indices(np.arange(60), split_by_interval_int=25, shuffle=False)
        agg = np.repeat(np.arange(20), cn)
        parts = partition_indices(agg, split_by_interval_int=25, shuffle=False)
        subs = [np.bincount(p, minlength=20) for p in parts]
        plan = chunking.plan_chunks(60, 20, cn, isc, subsampled_cell_number_to_node_assignment_list=subs)
    return sc, st, plan


def test_integer_count_matrices_travel_as_float64():
    """read_csv count matrices are int64: the distributed path must not down-cast them."""
    sys.path.insert(0, ROOT)
    from cytospace_b200 import chunking
    sc, st, plan = _problem("single_cell")
    a = chunking.solve_chunks(OracleEngine(), sc, st, plan, log_tpm=True)
    b = chunking.solve_chunks(OracleEngine(), sc.astype(np.int64), st.astype(np.int64), plan, log_tpm=True)
    assert a == b


KW = {"single_cell": {}, "sub_spots": {}, "single_cell_spearman": {"metric": "Spearman_correlation"},
      "sub_spots_cspr": {"metric": "Euclidean", "cspr_seed": 7}}
BASE = {"single_cell": "single_cell", "sub_spots": "sub_spots", "single_cell_spearman": "single_cell",
        "sub_spots_cspr": "sub_spots"}


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cytospace_b200 import chunking
        sc, st, plan = _problem(BASE[mode])
        if rank == 0:
            out = chunking.solve_chunks(OracleEngine(), sc, st, plan, log_tpm=True, **KW[mode])
        else:
            out = chunking.solve_chunks(OracleEngine(), None, None, None, log_tpm=True)   # kwargs travel from rank 0
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["single_cell", "sub_spots", "single_cell_spearman", "sub_spots_cspr"])
def test_two_rank_chunk_distribution_matches_single_process(mode):
    sys.path.insert(0, ROOT)
    from cytospace_b200 import chunking
    sc, st, plan = _problem(BASE[mode])
    expect = chunking.solve_chunks(OracleEngine(), sc, st, plan, log_tpm=True, **KW[mode])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + sorted(KW).index(mode)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == expect and got[1] == expect
    assert len(expect) == len(plan) and all(len(e) == c.n for e, c in zip(expect, plan))
