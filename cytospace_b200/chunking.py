"""GPU chunk planner: CytoSPACE's independent sub-LAPs mapped onto ranks / GPUs.

Replaces the ``ProcessPoolExecutor`` fan-out of ``apply_linear_assignment``
(cytospace/cytospace.py:430-467).  The reference already splits the problem into independent
sub-problems (``--single-cell -noss K``: matched blocks of spots and cells, :605-633;
``--sampling-sub-spots -nosss K``: blocks of cells against all spots with sub-sampled
capacities, :650-660); each chunk is one ``solve_linear_assignment_problem`` call with no
cross-chunk coupling, so chunks are simply dealt to ranks (one process per GPU).

Data movement when a ``torch.distributed`` group is up (NCCL over NVLink on GPUs, gloo in the
CPU tests): rank 0 owns the expression matrices; the ST block every chunk shares
(``--sampling-sub-spots``, cytospace.py:438) goes out with ONE broadcast, per-chunk column blocks
go point-to-point to their owner, and the assignment vectors come back with one all-gather of
int32 indices.  Nothing is exchanged during a solve.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Chunk:
    idx: int
    sc_index: np.ndarray            # columns of the scRNA matrix in this chunk
    st_index: np.ndarray | None     # columns of the ST matrix (None: every spot, cytospace.py:438)
    cn: np.ndarray                  # cells per selected spot (sums to len(sc_index))

    @property
    def n(self) -> int:
        return int(len(self.sc_index))


def plan_chunks(n_cells, n_spots, cell_number_to_node_assignment, index_sc_list, index_st_list=None,
                subsampled_cell_number_to_node_assignment_list=None) -> list[Chunk]:
    """The chunk list of cytospace.py:406-443: one chunk (no lists), ``--single-cell`` (index_st_list)
    or ``--sampling-sub-spots`` (per-chunk capacity vectors)."""
    cn = np.asarray(cell_number_to_node_assignment)
    if index_st_list is not None and subsampled_cell_number_to_node_assignment_list is not None:
        raise ValueError("index_st_list and subsampled_cell_number_to_node_assignment_list cannot both be specified")
    if index_st_list is None and subsampled_cell_number_to_node_assignment_list is None:
        return [Chunk(0, np.asarray(index_sc_list[0]), None, cn)]
    chunks = []
    if index_st_list is not None:
        if len(index_st_list) != len(index_sc_list):
            raise ValueError("index_sc_list and index_st_list must have the same number of partitions")
        for i, (isc, ist) in enumerate(zip(index_sc_list, index_st_list)):
            ist = np.asarray(ist)
            chunks.append(Chunk(i, np.asarray(isc), ist, cn[ist]))
    else:
        subs = subsampled_cell_number_to_node_assignment_list
        if len(subs) != len(index_sc_list):
            raise ValueError("index_sc_list and the sub-sampled capacity list must have the same length")
        for i, (isc, sub) in enumerate(zip(index_sc_list, subs)):
            chunks.append(Chunk(i, np.asarray(isc), None, np.asarray(sub)))
    return chunks


def assign_ranks(sizes, world: int) -> list[int]:
    """Owner rank per chunk: longest-processing-time first on the weight n^2.5 (the measured growth
    of the LAP), ties to the lowest rank -- deterministic on every rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-float(sizes[i]) ** 2.5, i))
    load = [0.0] * world
    owner = [0] * len(sizes)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += float(sizes[i]) ** 2.5
    return owner


def _dist_state(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(group), dist.get_world_size(group)
    return None, 0, 1


def solve_chunks(engine, sc_np, st_np, plan: list[Chunk], log_tpm: bool = False, group=None, **assign_kw):
    """Solve every chunk; returns ``mapped_st_index`` (list of int per cell, chunk-local spot index
    as in cytospace.py:455-459) for each chunk, on every rank.

    Single process: chunks run back to back on ``engine``.  With a process group: rank 0 passes the
    matrices (other ranks may pass ``None``), chunks are dealt by ``assign_ranks``.
    ``assign_kw`` (``metric=``, ``cspr_seed=``) goes to ``engine.assign`` unchanged (rank 0's values)."""
    dist, rank, world = _dist_state(group)
    if world == 1:
        out = []
        for ch in plan:
            sc = sc_np[:, ch.sc_index] if len(ch.sc_index) != sc_np.shape[1] or not _is_arange(ch.sc_index) else sc_np
            st = st_np if ch.st_index is None else st_np[:, ch.st_index]
            spot_of_cell, _, _ = engine.assign(np.ascontiguousarray(sc), np.ascontiguousarray(st), ch.cn,
                                               log_tpm=log_tpm, **assign_kw)
            out.append(spot_of_cell.cpu().numpy().tolist())
        return out
    return _solve_chunks_distributed(dist, rank, world, engine, sc_np, st_np, plan, log_tpm, group, assign_kw)


def _is_arange(idx) -> bool:
    idx = np.asarray(idx)
    return idx.size > 0 and idx[0] == 0 and np.array_equal(idx, np.arange(idx.size))


def _solve_chunks_distributed(dist, rank, world, engine, sc_np, st_np, plan, log_tpm, group, assign_kw):
    dev = engine.device
    # plan and shapes travel as one small object broadcast (host metadata, not the data path)
    meta = [None]
    if rank == 0:
        meta[0] = (plan, int(sc_np.shape[0]), int(st_np.shape[1]), str(sc_np.dtype), dict(assign_kw))
    dist.broadcast_object_list(meta, src=0, group=group)
    plan, n_genes, n_spots, dt, assign_kw = meta[0]
    tdt = torch.float64 if dt == "float64" else torch.float32
    owner = assign_ranks([c.n for c in plan], world)
    shared_st = any(c.st_index is None for c in plan)

    st_all = None
    if shared_st:
        # ONE broadcast of the ST block every chunk reads (cytospace.py:438)
        if rank == 0:
            st_all = torch.from_numpy(np.ascontiguousarray(st_np)).to(tdt).to(dev)
        else:
            st_all = torch.empty((n_genes, n_spots), dtype=tdt, device=dev)
        dist.broadcast(st_all, src=0, group=group)

    # per-chunk column blocks: point-to-point from rank 0 to the owner
    mine = {}
    pending = []
    for ch in plan:
        o = owner[ch.idx]
        need_st = ch.st_index is not None
        if rank == 0:
            sc_blk = torch.from_numpy(np.ascontiguousarray(sc_np[:, ch.sc_index])).to(tdt).to(dev)
            st_blk = torch.from_numpy(np.ascontiguousarray(st_np[:, ch.st_index])).to(tdt).to(dev) if need_st else None
            if o == 0:
                mine[ch.idx] = (sc_blk, st_blk)
            else:
                pending.append((dist.isend(sc_blk, dst=o, group=group), sc_blk))
                if need_st:
                    pending.append((dist.isend(st_blk, dst=o, group=group), st_blk))
        elif o == rank:
            sc_blk = torch.empty((n_genes, ch.n), dtype=tdt, device=dev)
            dist.recv(sc_blk, src=0, group=group)
            st_blk = None
            if need_st:
                st_blk = torch.empty((n_genes, len(ch.st_index)), dtype=tdt, device=dev)
                dist.recv(st_blk, src=0, group=group)
            mine[ch.idx] = (sc_blk, st_blk)
    for req, _keep in pending:
        req.wait()

    # solve: no communication
    results = {}
    for ch in plan:
        if owner[ch.idx] != rank:
            continue
        sc_blk, st_blk = mine.pop(ch.idx)
        spot_of_cell, _, _ = engine.assign(sc_blk, st_all if st_blk is None else st_blk, ch.cn, log_tpm=log_tpm,
                                           **assign_kw)
        results[ch.idx] = spot_of_cell.to(torch.int32)

    # one all-gather of the assignment indices (padded to the largest per-rank total)
    per_rank = [sum(c.n for c in plan if owner[c.idx] == r) for r in range(world)]
    width = max(per_rank) if per_rank else 0
    send = torch.full((max(width, 1),), -1, dtype=torch.int32, device=dev)
    pos = 0
    for ch in plan:
        if owner[ch.idx] == rank:
            send[pos:pos + ch.n] = results[ch.idx]
            pos += ch.n
    gathered = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(gathered, send, group=group)
    host = [g.cpu().numpy() for g in gathered]
    cursor = [0] * world
    out = []
    for ch in plan:
        r = owner[ch.idx]
        out.append(host[r][cursor[r]:cursor[r] + ch.n].tolist())
        cursor[r] += ch.n
    return out
