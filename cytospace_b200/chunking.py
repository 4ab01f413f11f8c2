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


class TorchTransport:
    """Data plane over a ``torch.distributed`` process group (NCCL on GPUs, gloo in the CPU tests).
    Ranks are GROUP ranks; they are translated to global ranks where the API wants them."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def _g(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def bcast_object(self, obj, root=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=self._g(root), group=self.group)
        return box[0]

    def broadcast(self, t, root=0):
        self.dist.broadcast(t, src=self._g(root), group=self.group)

    def isend(self, t, dst):
        return self.dist.isend(t, dst=self._g(dst), group=self.group)

    def recv(self, t, src):
        self.dist.recv(t, src=self._g(src), group=self.group)

    def all_gather(self, send):
        out = [torch.empty_like(send) for _ in range(self.world)]
        self.dist.all_gather(out, send, group=self.group)
        return out


#: owners fed concurrently by one NCCL group of sends (rank 0 holds the blocks of two groups at most)
SEND_GROUP = 4

#: send float64 blocks as float32 when every value survives the round trip exactly (raw counts do)
NARROW_EXACT = True


#: pass as ``transport=`` to run every chunk on the calling rank although a process group is up
SOLO = "solo"


_native_transports = {}


def _transport(group=None, transport=None, engine=None):
    """The data plane of a distributed solve: an explicit transport, else -- with a process group of more than one
    rank -- NCCL through the C ABI (``dist_native.NativeTransport``, bootstrapped over the group) when the group's
    backend is NCCL and the engine lives on a GPU, ``TorchTransport`` otherwise (gloo in the CPU tests)."""
    if transport is SOLO:
        return None
    if transport is not None:
        return transport
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return None
    dev = getattr(engine, "device", None)
    if dist.get_backend(group) == "nccl" and dev is not None and dev.type == "cuda":
        key = (id(engine), id(group))
        if key not in _native_transports:
            from . import dist_native
            _native_transports[key] = dist_native.NativeTransport.from_torch_group(engine, group)
        return _native_transports[key]
    return TorchTransport(dist, group)


#: bytes moved by the last distributed solve on this rank (bench.py reports them)
last_traffic = {"bcast_bytes": 0, "p2p_bytes": 0, "gather_bytes": 0}


def solve_chunks(engine, sc_np, st_np, plan: list[Chunk], log_tpm: bool = False, group=None, transport=None,
                 **assign_kw):
    """Solve every chunk; returns ``mapped_st_index`` (list of int per cell, chunk-local spot index
    as in cytospace.py:455-459) for each chunk, on every rank.

    Single process: chunks run back to back on ``engine``.  With a process group (or an explicit
    ``transport``, e.g. ``dist_native.NativeTransport``: NCCL through the C ABI): rank 0 passes the
    matrices -- host arrays or tensors already on its GPU; other ranks may pass ``None`` -- and chunks are
    dealt by ``assign_ranks``.  ``assign_kw`` (``metric=``, ``cspr_seed=``) goes to ``engine.assign``
    unchanged (rank 0's values)."""
    tp = _transport(group, transport, engine)
    if tp is None:
        out = []
        sc_dev = st_dev = None
        # (explicit dtypes: the chunk paths, distributed or not, never narrow on upload -- same bits on every route)
        keep = lambda x: None if (torch.is_tensor(x) or np.asarray(x).dtype == np.float32) else torch.float64
        if len(plan) > 1:
            # upload once, slice the chunks' columns on the device
            sc_dev, st_dev = engine.to_device(sc_np, keep(sc_np)), engine.to_device(st_np, keep(st_np))
        for ch in plan:
            if sc_dev is None:
                sc, st = sc_np, st_np
                if len(ch.sc_index) != sc_np.shape[1] or not _is_arange(ch.sc_index):
                    sc = _columns(engine, engine.to_device(sc_np, keep(sc_np)), ch.sc_index)
                if ch.st_index is not None:
                    st = _columns(engine, engine.to_device(st_np, keep(st_np)), ch.st_index)
            else:
                sc = _columns(engine, sc_dev, ch.sc_index)
                st = st_dev if ch.st_index is None else _columns(engine, st_dev, ch.st_index)
            spot_of_cell, _, _ = engine.assign(sc, st, ch.cn, log_tpm=log_tpm, **assign_kw)
            out.append(spot_of_cell.cpu().numpy().tolist())
        return out
    return _solve_chunks_distributed(tp, engine, sc_np, st_np, plan, log_tpm, assign_kw)


def _is_arange(idx) -> bool:
    idx = np.asarray(idx)
    return idx.size > 0 and idx[0] == 0 and np.array_equal(idx, np.arange(idx.size))


def _columns(engine, x_dev, idx, tp=None):
    """x_dev[:, idx] as a contiguous device matrix (device gather: the host never touches the columns);
    through the C ABI (``cyb_gather_columns``) when the transport offers it."""
    idx = np.asarray(idx)
    if idx.size == x_dev.shape[1] and _is_arange(idx):
        return x_dev
    if tp is not None and hasattr(tp, "gather_columns") and x_dev.is_cuda and x_dev.stride(1) == 1:
        return tp.gather_columns(x_dev, idx)
    return x_dev.index_select(1, torch.from_numpy(idx.astype(np.int64)).to(x_dev.device))


def _solve_chunks_distributed(tp, engine, sc_np, st_np, plan, log_tpm, assign_kw):
    dev = engine.device
    rank, world = tp.rank, tp.world
    traffic = {"bcast_bytes": 0, "p2p_bytes": 0, "gather_bytes": 0}
    # rank 0: ONE upload of each matrix (pinned staging ring); chunk columns are gathered on the device
    sc_dev = st_dev = None
    meta = None
    if rank == 0:
        # float32 stays float32; anything else (float64, integer counts) is float64 as in the reference
        dt = "float32" if str(sc_np.dtype).endswith("float32") and str(st_np.dtype).endswith("float32") else "float64"
        tdt = torch.float64 if dt == "float64" else torch.float32
        sc_dev = engine.to_device(sc_np, tdt)
        st_dev = engine.to_device(st_np, tdt)
        # Lossless narrowing of the wire format: raw count matrices (what apply_linear_assignment hands over,
        # normalize_data being fused into the device pre-pass) are small integers stored as float64 -- every value
        # is exactly a float32, so the blocks travel as float32 (half the NVLink bytes) and the owners, whose kernels
        # widen every element back to double on load, compute bit-identical statistics.  Checked, not assumed.
        if tdt == torch.float64 and NARROW_EXACT and sc_dev.is_cuda and hasattr(engine, "narrow_exact"):
            sc32, st32 = engine.narrow_exact(sc_dev), engine.narrow_exact(st_dev)       # one fused pass each
            if sc32 is not None and st32 is not None:
                dt, sc_dev, st_dev = "float32", sc32, st32
            del sc32, st32
        # plan, shapes and wire dtype travel as one small object broadcast (host metadata, not the data path)
        meta = (plan, int(sc_np.shape[0]), int(st_np.shape[1]), dt, dict(assign_kw))
    plan, n_genes, n_spots, dt, assign_kw = tp.bcast_object(meta, root=0)
    tdt = torch.float64 if dt == "float64" else torch.float32
    esz = 8 if dt == "float64" else 4
    owner = assign_ranks([c.n for c in plan], world)
    shared_st = any(c.st_index is None for c in plan)

    st_all = None
    if shared_st:
        # ONE broadcast of the ST block every chunk reads (cytospace.py:438)
        st_all = st_dev if rank == 0 else torch.empty((n_genes, n_spots), dtype=tdt, device=dev)
        tp.broadcast(st_all, root=0)
        traffic["bcast_bytes"] += n_genes * n_spots * esz

    # per-chunk column blocks: point-to-point from rank 0 to the owner.  With a grouping transport the blocks of up to
    # SEND_GROUP chunks (different owners) leave as one NCCL group, so that rank 0's whole NVLink egress is used; at
    # most two groups (or two single blocks) are in flight, each freed as soon as its send has completed.
    mine = {}
    inflight = []
    batch = []

    def flush():
        if batch:
            inflight.append((tp.send_group(list(batch)), list(batch)))
            batch.clear()
        while len(inflight) > 2:
            req, _keep = inflight.pop(0)
            req.wait()

    for ch in plan:
        o = owner[ch.idx]
        need_st = ch.st_index is not None
        if rank == 0:
            blocks = [_columns(engine, sc_dev, ch.sc_index, tp)]
            if need_st:
                blocks.append(_columns(engine, st_dev, ch.st_index, tp))
            if o == 0:
                mine[ch.idx] = (blocks[0], blocks[1] if need_st else None)
            elif hasattr(tp, "send_group"):
                for blk in blocks:
                    batch.append((blk.contiguous(), o))
                    traffic["p2p_bytes"] += blk.numel() * esz
                if len({d for _b, d in batch}) >= SEND_GROUP:
                    flush()
            else:
                for blk in blocks:
                    blk = blk.contiguous()
                    inflight.append((tp.isend(blk, o), blk))
                    traffic["p2p_bytes"] += blk.numel() * esz
                while len(inflight) > 2:
                    req, _blk = inflight.pop(0)
                    req.wait()
        elif o == rank:
            sc_blk = torch.empty((n_genes, ch.n), dtype=tdt, device=dev)
            tp.recv(sc_blk, 0)
            st_blk = None
            if need_st:
                st_blk = torch.empty((n_genes, len(ch.st_index)), dtype=tdt, device=dev)
                tp.recv(st_blk, 0)
            mine[ch.idx] = (sc_blk, st_blk)
            traffic["p2p_bytes"] += (sc_blk.numel() + (st_blk.numel() if need_st else 0)) * esz
    if rank == 0:
        flush()
    for req, _blk in inflight:
        req.wait()
    inflight.clear()
    if rank == 0 and not any(owner[c.idx] == 0 and c.st_index is None for c in plan):
        st_dev = None
    sc_dev = None                                  # rank 0 keeps only its own chunks' blocks

    # solve: no communication
    results = {}
    for ch in plan:
        if owner[ch.idx] != rank:
            continue
        sc_blk, st_blk = mine.pop(ch.idx)
        spot_of_cell, _, _ = engine.assign(sc_blk, st_all if st_blk is None else st_blk, ch.cn, log_tpm=log_tpm,
                                           **assign_kw)
        results[ch.idx] = spot_of_cell.to(torch.int32)
        del sc_blk, st_blk

    # one all-gather of the assignment indices (padded to the largest per-rank total)
    per_rank = [sum(c.n for c in plan if owner[c.idx] == r) for r in range(world)]
    width = max(per_rank) if per_rank else 0
    send = torch.full((max(width, 1),), -1, dtype=torch.int32, device=dev)
    pos = 0
    for ch in plan:
        if owner[ch.idx] == rank:
            send[pos:pos + ch.n] = results[ch.idx]
            pos += ch.n
    gathered = tp.all_gather(send)
    traffic["gather_bytes"] += send.numel() * 4 * world
    host = [g.cpu().numpy() for g in gathered]
    cursor = [0] * world
    out = []
    for ch in plan:
        r = owner[ch.idx]
        out.append(host[r][cursor[r]:cursor[r] + ch.n].tolist())
        cursor[r] += ch.n
    last_traffic.update(traffic)
    return out
