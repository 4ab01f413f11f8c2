"""GPU: the chunk distribution over NCCL on real devices (needs >= 2 GPUs; skipped on a 1-GPU box) -- the same
broadcast / point-to-point / all-gather plumbing the gloo CPU test drives with an oracle stub, here with the device
engine on every rank and compared with the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _problem(mode):
    sys.path.insert(0, ROOT)
    from cytospace_b200 import synthetic as syn, chunking
    from cytospace_b200.cytospace import partition_indices
    if mode == "single_cell":
        sc, st, cn = syn.structured_counts(900, 900, 600, 1, seed=5)
        isc = partition_indices(np.arange(900), split_by_interval_int=250, shuffle=False)
        ist = partition_indices(np.arange(900), split_by_interval_int=250, shuffle=False)
        plan = chunking.plan_chunks(900, 900, cn, isc, index_st_list=ist)
    else:
        sc, st, cn = syn.structured_counts(600, 200, 600, 3, seed=6)
        isc = partition_indices(np.arange(600), split_by_interval_int=250, shuffle=False)
        parts = partition_indices(np.repeat(np.arange(200), cn), split_by_interval_int=250, shuffle=False)
        subs = [np.bincount(p, minlength=200) for p in parts]
        plan = chunking.plan_chunks(600, 200, cn, isc, subsampled_cell_number_to_node_assignment_list=subs)
    return sc, st, plan


def _worker(rank, world, port, mode, native, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cytospace_b200 import chunking
        from cytospace_b200.engine import AssignmentEngine
        eng = AssignmentEngine(device=f"cuda:{rank}")
        # default on GPUs: NCCL through the C ABI (dist_native.NativeTransport); the other variant sends the same
        # tensors through torch.distributed's own NCCL process group
        tp = None if native else chunking.TorchTransport(dist)
        sc, st, plan = _problem(mode)
        if rank == 0:
            out = chunking.solve_chunks(eng, sc, st, plan, log_tpm=True, transport=tp)
        else:
            out = chunking.solve_chunks(eng, None, None, None, log_tpm=True, transport=tp)
        q.put((rank, out, dict(chunking.last_traffic), type(chunking._transport(None, tp, eng)).__name__))
        for t_ in chunking._native_transports.values():
            t_.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("native", [False, True], ids=["torch_nccl", "c_abi_nccl"])
@pytest.mark.parametrize("mode", ["single_cell", "sub_spots"])
def test_two_gpu_chunk_distribution_matches_single_process(engine, mode, native):
    sys.path.insert(0, ROOT)
    from cytospace_b200 import chunking
    sc, st, plan = _problem(mode)
    expect = chunking.solve_chunks(engine, sc, st, plan, log_tpm=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + (2 if native else 0) + (1 if mode == "sub_spots" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, native, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, out, traffic, kind = q.get(timeout=300)
        assert kind == ("NativeTransport" if native else "TorchTransport")
        got[r] = (out, traffic)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][0] == expect and got[1][0] == expect
    assert got[0][1]["p2p_bytes"] > 0 and got[0][1]["gather_bytes"] > 0
    if mode == "sub_spots":
        # ONE broadcast of the shared ST block; raw counts are exact in float32, so that is the wire format
        assert got[0][1]["bcast_bytes"] == sc.shape[0] * st.shape[1] * 4
