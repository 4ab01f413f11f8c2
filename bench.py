#!/usr/bin/env python
"""bench.py -- cell-spot assignments/sec of the CytoSPACE assignment hot path on B200.

A "step" is one pass of the hot path (Pearson cost build + exact LAP) over one synthetic
N-cell x S-spot x G-gene problem.  Default workload = BASELINE.json configs[1]
(10k x 10k x 20k genes).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4] [--impl reference]

* ``value``  : whole-job assignments/s, inputs resident in HBM when the timed region starts.
* ``e2e``    : same metric through the reference-facing call with HOST buffers (pinned H2D of both
               expression matrices + D2H of the assignment inside the timed region).
* ``roofline``: the dominant kernel (the LAP auction kernel): algorithmic row-scan bytes / its
               CUDA-event time vs the measured HBM peak (MEASURED_PEAKS.json).
* ``cpu_baseline``: the oracle (restated JV + numpy cost build) on this box's host cores.
* ``--impl reference``: the reference's own CPU formulation (float64 numpy cost build, 1e-16 tie-noise,
               restated float64 JV -- the lapjv wheel is absent) on a bounded sample of the workload.
N > 1 (torchrun): one rank per GPU, each rank solves its own independent sub-problem (CytoSPACE's
chunks are independent, cytospace.py:430-451), one all-gather of the assignment indices per step.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": dict(n_cells=1000, n_spots=1000, n_genes=2000, cps=1, seed=1001,
                 desc="synthetic 1k cells x 1k spots x 2k genes"),
    "cfg2": dict(n_cells=10000, n_spots=10000, n_genes=20000, cps=1, seed=1002,
                 desc="synthetic 10k cells x 10k spots x 20k genes, cost-build + LAP"),
    "cfg3": dict(n_cells=50000, n_spots=50000, n_genes=20000, cps=1, seed=1003,
                 desc="synthetic 50k cells x 50k spots x 20k genes"),
    "cfg4": dict(n_cells=30000, n_spots=5000, n_genes=30000, cps=6, seed=1004,
                 desc="Visium-shaped 30k cells x 5k spots (6 cells/spot) x 30k genes"),
    "chunk25k": dict(n_cells=25000, n_spots=25000, n_genes=20000, cps=1, seed=1005,
                     desc="one 25k x 25k x 20k-gene sub-LAP of the 200k chunked problem (cfg5)"),
    "cfg5": dict(n_cells=25000, n_spots=25000, n_genes=20000, cps=1, seed=1005, chunks=8,
                 desc="synthetic 200k cells x 200k spots chunked into 8 sub-LAPs of 25k (--single-cell -noss 25000), "
                      "strong scaling over the ranks"),
}
METRIC = "cell-spot assignments/sec"
UNIT = "assignments/s"


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch measured once with `ncu --set full` (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get(kernel)
    except (OSError, ValueError):
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def reference_step(sc_n, st_n, cn, seed=1):
    """The reference's formulation of one solve (cytospace.py:304-332) on the CPU: float64 numpy cost
    build (LAS:42-69 / COM:190-199), 1e-16 * U(0,1) tie-noise (CYT:325-327), dense JV on float64
    (restated: the lapjv wheel is absent), location_repeat[assignment] (CYT:331)."""
    import oracle
    from oracle import cost_oracle as co
    distance_repeat, location_repeat = co.calculate_cost(sc_n, st_n, cn)
    np.random.seed(seed)
    cost_scaled = distance_repeat + 1e-16 * np.random.rand(*distance_repeat.shape)
    _, colsol, (total, _, _) = oracle.lapjv_f64(cost_scaled)
    return location_repeat[colsol], total


def _reference_worker(q, n, n_spots, wl, steps, warmup, threads):
    """One ``-nop`` worker of the reference arm: its own copy of the sample, BLAS limited to its share of the cores."""
    from threadpoolctl import threadpool_limits
    from oracle import cost_oracle as co
    from cytospace_b200 import synthetic as syn
    with threadpool_limits(limits=threads):
        sc, st, cn = syn.structured_counts(n, n_spots, wl["n_genes"], wl["cps"], seed=wl["seed"])
        sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
        del sc, st
        for _ in range(warmup):
            reference_step(sc_n, st_n, cn)
        q.put(("ready", None))
        t0 = time.perf_counter()
        total = None
        for _ in range(steps):
            _, total = reference_step(sc_n, st_n, cn)
        q.put(("done", (time.perf_counter() - t0, float(total))))


def run_reference(args, wl):
    """The reference's own CPU formulation on this box's host cores.  N > 1 mirrors ``-nop N``
    (cytospace.py:430: one worker process per chunk, ``min(num_chunks, nop)`` at a time): N independent copies of
    the sample on ``min(N, cores)`` worker processes sharing the cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    n_units = max(1, args.gpus)
    workers = min(n_units, cores)
    threads = max(1, cores // workers)
    # CPU code needs no warm-up: at most one untimed step, so that the timed ones can be as large as possible
    warmup = min(args.warmup, 1)
    # one full-size cfg2 step measured here: 42 s on 8 cores (float64 cost build ~25 s, tie noise, float64 JV ~12 s)
    budget_s = 330.0 / max(1, args.steps + warmup) / -(-n_units // workers)
    est_full = 42.0 * (wl["n_cells"] / 10000.0) ** 2.6 * (wl["n_genes"] / 20000.0) ** 0.5 * (8.0 / min(16, max(threads, 1))) ** 0.5
    n = min(500, wl["n_cells"])
    for cand in (wl["n_cells"], 8000, 7000, 6000, 5000, 4000, 3000, 2000, 1000, 500):
        if cand <= wl["n_cells"] and est_full * (cand / wl["n_cells"]) ** 2.6 <= budget_s:
            n = cand
            break
    n_spots = max(1, n // wl["cps"])
    n = n_spots * wl["cps"]
    # torchrun exports OMP_NUM_THREADS=1 to its children: the reference arm is a CPU job and uses the cores
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(threads)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    dt, total = 0.0, None
    for wave in range(0, n_units, workers):
        procs = [ctx.Process(target=_reference_worker, args=(q, n, n_spots, wl, args.steps, warmup, threads))
                 for _ in range(min(workers, n_units - wave))]
        for p_ in procs:
            p_.start()
        got = [q.get() for _ in range(2 * len(procs))]
        for p_ in procs:
            p_.join()
        times = [v[0] for k, v in got if k == "done"]
        total = [v[1] for k, v in got if k == "done"][0]
        dt += max(times)                                   # the waves run one after the other
    dt /= max(1, args.steps)
    value = n * n_units / dt
    sample = (f"{n} cells x {n_spots} spots x {wl['n_genes']} genes block of {args.workload} "
              f"(full problem {wl['n_cells']} x {wl['n_spots']}); per step: float64 numpy cost build + 1e-16 tie noise + "
              f"restated float64 JV (single-threaded like lapjv; the wheel is absent); {n_units} independent unit(s) on "
              f"{workers} worker process(es) x {threads} BLAS thread(s) (mirrors -nop {n_units}); {warmup} untimed warm-up step(s)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "sample_n": n, "same_config": bool(n == wl["n_cells"]),
                       "units": n_units, "worker_processes": workers, "blas_threads_per_worker": threads},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "total_cost_f64": total}
    emit(line)


# ----------------------------------------------------------------------------------- B200 arm
def cpu_baseline(cost_np, row_map, sc_host, st_host, wl, total_gpu, with_scipy=False):
    """Oracle on this box's host cores: restated JV (int32, 1 thread) on the GPU-built matrix (full
    size up to 10k, else its leading 8k block) + numpy float64 cost build on a 2k x 2k x G block, which is also
    compared entry by entry with the same block of the GPU-built matrix (cost-build tolerance at the full G)."""
    import oracle
    from oracle import cost_oracle as co
    oracle.build()
    cores = os.cpu_count() or 1
    n = cost_np.shape[1]
    b = min(2000, sc_host.shape[1], st_host.shape[1])
    t0 = time.perf_counter()
    blk = co.cost_matrix_i32(sc_host[:, :b], st_host[:, :b])          # spots x cells (reference orientation)
    t_cost_blk = time.perf_counter() - t0
    t_cost = t_cost_blk * (sc_host.shape[1] / b) * (st_host.shape[1] / b)
    out = {"unit": UNIT, "cores": cores, "kind": "port", "lap_threads": 1, "cost_build_threads": cores}
    if row_map is None:
        diff = np.abs(blk.astype(np.int64) - cost_np[:b, :b].astype(np.int64))
        out["cost_block_max_abs_diff"] = int(diff.max())
        out["cost_block_frac_within_1"] = float((diff <= 1).mean())
    if n <= 10000:
        t0 = time.perf_counter()
        total_cpu = oracle.lapjv_i32(cost_np, row_map)[2][0]
        t_lap = time.perf_counter() - t0
        try:
            if with_scipy and n <= 10000 and row_map is None:           # ~90 s at 10k: opt-in (--cpu-scipy)
                from scipy.optimize import linear_sum_assignment
                t0 = time.perf_counter()
                ri, ci = linear_sum_assignment(cost_np.astype(np.float64))
                out["scipy_lap_s"] = time.perf_counter() - t0
                out["scipy_total_equal"] = bool(int(cost_np[ri, ci].astype(np.int64).sum()) == total_gpu)
        except Exception as e:            # a reported extra, never allowed to cost the bench line
            out["scipy_error"] = repr(e)[:200]
        out.update(value=n / (t_cost + t_lap), lap_s=t_lap, cost_build_s_scaled=t_cost,
                   total_cost=int(total_cpu), total_equal=bool(total_cpu == total_gpu),
                   sample=f"LAP: restated JV (int32, 1 thread) on the full GPU-built {n}x{n} matrix; cost build: "
                          f"numpy float64 on a {b}x{b}x{wl['n_genes']} block ({cores} BLAS threads) scaled by "
                          f"{(sc_host.shape[1] / b) * (st_host.shape[1] / b):.1f}")
    else:
        m = 8000
        sub = np.ascontiguousarray(cost_np[:m, :m]) if row_map is None else None
        if sub is None:
            rm = row_map[:m]
            t0 = time.perf_counter(); oracle.lapjv_i32(np.ascontiguousarray(cost_np[:, :m]), rm); t_blk = time.perf_counter() - t0
        else:
            t0 = time.perf_counter(); oracle.lapjv_i32(sub); t_blk = time.perf_counter() - t0
        t_lap = t_blk * (n / m) ** 2.6
        out.update(value=n / (t_cost + t_lap), lap_s=t_lap, cost_build_s_scaled=t_cost, total_equal=None,
                   sample=f"LAP: restated JV on the leading {m}x{m} block, extrapolated with n^2.6 (full-size totals: "
                          f"profiles/r02_parity_*.json); cost build: numpy float64 on a {b}x{b} block scaled")
    return out


def strong_cfg5(eng, dev, rank, world, args):
    """BASELINE configs[4]: 200k cells x 200k spots as 8 matched sub-LAPs of 25k (--single-cell -noss 25000,
    cytospace.py:605-633) through chunking.solve_chunks -- rank 0 holds the expression matrices in HBM, the blocks
    are gathered on its GPU and travel point-to-point over NCCL INSIDE the timed region, one all-gather returns the
    indices.  Also timed: the same plan on rank 0's GPU alone (the 1-GPU reference of the speed-up)."""
    import torch
    import torch.distributed as dist
    from cytospace_b200 import chunking, synthetic as syn
    from cytospace_b200.cytospace import partition_indices
    wl = WORKLOADS["cfg5"]
    n_chunk, n_chunks, G = wl["n_cells"], wl["chunks"], wl["n_genes"]
    n = n_chunk * n_chunks
    if args.cfg5_cells:
        n = int(args.cfg5_cells); n_chunk = n // n_chunks
    plan = None
    sc = st = None
    if rank == 0:
        # RAW counts (float64, as read_data returns them): normalize_data is fused into the device pre-pass
        # (log_tpm), exactly the flow of cytospace_b200.apply_linear_assignment
        sc, st, cn = syn.structured_counts_torch(n, n, G, 1, seed=wl["seed"], device=dev)
        isc = partition_indices(np.arange(n), split_by_interval_int=n_chunk, shuffle=False)
        ist = partition_indices(np.arange(n), split_by_interval_int=n_chunk, shuffle=False)
        plan = chunking.plan_chunks(n, n, cn, isc, index_st_list=ist)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, out

    dist_fn = lambda: chunking.solve_chunks(eng, sc, st, plan, log_tpm=True)
    dist_fn()                                                      # warm-up (NCCL connections, workspaces)
    ms_n, out_n = timed(dist_fn, 2)
    res = {"workload": f"{n} cells x {n} spots x {G} genes, {n_chunks} sub-LAPs of {n_chunk} (raw float64 count matrices resident "
                       f"in rank 0's HBM, normalize_data fused on the owners; distribution inside the timed region; the blocks "
                       f"travel as float32 when every value is exactly representable -- checked on the device)",
           "n_gpus": world, "ms_per_step": ms_n, "value": n / (ms_n / 1e3), "unit": UNIT, "scaling": "strong",
           **{k: int(v) for k, v in chunking.last_traffic.items()}} if world > 1 else \
          {"workload": f"{n} cells x {n} spots x {G} genes, {n_chunks} sub-LAPs of {n_chunk} back to back on one GPU (raw float64 "
                       f"count matrices, normalize_data fused)",
           "n_gpus": 1, "ms_per_step": ms_n, "value": n / (ms_n / 1e3), "unit": UNIT, "scaling": "strong",
           "bcast_bytes": 0, "p2p_bytes": 0, "gather_bytes": 0}
    if world > 1:
        # the 1-GPU time of the same plan, on rank 0 (the other ranks wait at the barrier)
        ms_1 = None
        if rank == 0:
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            t0.record()
            out_1 = chunking.solve_chunks(eng, sc, st, plan, log_tpm=True, transport=chunking.SOLO)
            t1.record(); torch.cuda.synchronize(dev)
            ms_1 = t0.elapsed_time(t1)
            res["ms_per_step_1gpu"] = ms_1
            res["speedup_vs_1gpu"] = ms_1 / ms_n
            res["same_assignment_as_1gpu"] = bool(out_1 == out_n)
        barrier()
    return res


def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import cytospace_b200
    from cytospace_b200 import synthetic as syn
    from cytospace_b200 import linear_assignment_solvers as las
    from cytospace_b200.engine import AssignmentEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = AssignmentEngine(device=dev, precision=args.precision)
    las._engine = eng                      # the engine behind the plugin entry points of this process
    if world > 1:                          # the host cores are shared by the ranks' staging threads (read at first use)
        os.environ.setdefault("CYB_STAGE_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // world))))
    eng.profile = True
    n_cells, n_spots, n_genes, cps = wl["n_cells"], wl["n_spots"], wl["n_genes"], wl["cps"]
    n_chunks = wl.get("chunks", 0)          # --workload cfg5: sub-LAPs dealt round-robin, inputs resident per rank
    strong = n_chunks > 0
    my_chunks = [c for c in range(n_chunks) if c % world == rank] if strong else [rank]

    # ---- synthetic inputs (sampled on the device, normalised like CYT:398-399) -- not timed
    units = []
    for c in my_chunks:
        # weak scaling: every rank gets the SAME instance (per-GPU work fixed as N grows; the solve time of an
        # instance is data-dependent); the strong-scaling chunks are all different
        sc_raw, st_raw, cn = syn.structured_counts_torch(n_cells, n_spots, n_genes, cps,
                                                         seed=wl["seed"] + (17 * c if strong else 0), device=dev)
        sc_dev = syn.normalize_data_torch(sc_raw); del sc_raw
        st_dev = syn.normalize_data_torch(st_raw); del st_raw
        units.append((sc_dev, st_dev, cn))
    # host image of one unit: plain (pageable) numpy arrays, what solve_linear_assignment_problem receives from
    # apply_linear_assignment (cytospace.py:398-409)
    sc_np = units[0][0].cpu().numpy()
    st_np = units[0][1].cpu().numpy()
    h2d = (sc_np.nbytes + st_np.nbytes) * len(units)      # host bytes handed to the call; the bytes on PCIe are counted by the engine
    d2h = n_cells * 8 * len(units)
    n_out = n_cells * max(1, len(units))
    gathered = [torch.empty(n_out, dtype=torch.int64, device=dev) for _ in range(world)] if world > 1 else None
    even = (not strong) or (n_chunks % world == 0)

    def finish(spots):
        if world > 1 and even:
            dist.all_gather(gathered, torch.cat(spots) if len(spots) > 1 else spots[0])   # gather of assignment indices

    def step_resident():
        spots = []
        for sc_d, st_d, cn in units:
            spot, res, cost = eng.assign(sc_d, st_d, cn, metric=args.distance_metric)
            spots.append(spot)
        finish(spots)
        return spot, res, cost

    def step_e2e():
        # the call a user of the reference makes: host numpy arrays in, Python list out
        out = None
        for _sc_d, _st_d, cn in units:
            out = cytospace_b200.solve_linear_assignment_problem(sc_np, st_np, cn, "lapjv_b200", None, 1,
                                                                 args.distance_metric)
        return out

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, track=True):
        for _ in range(warmup):
            out = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lap_ms, cost_ms = [], []
        l0 = eng.launches
        e0.record()
        for _ in range(steps):
            out = fn()
            if track:
                lap_ms.append(eng.last_ms("lap")); cost_ms.append(eng.last_ms("cost"))
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, (float(np.mean(lap_ms)) if track else 0.0), (float(np.mean(cost_ms)) if track else 0.0), eng.launches - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tot_ms, (spot, res, cost), lap_ms, cost_ms, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    with contextlib.redirect_stdout(sys.stderr):
        e2e_ms, _, _, _, _ = timed(step_e2e, args.steps, min(args.warmup, 1), track=False)
    # bytes that actually crossed PCIe in ONE e2e step (the engine counts the device tensors it fills)
    b0 = eng.h2d_bytes
    with contextlib.redirect_stdout(sys.stderr):
        step_e2e()
    torch.cuda.synchronize(dev)
    h2d_wire = eng.h2d_bytes - b0
    # the same call with float64 on the wire (CYB_STAGE_F32=0 / engine.stage_float32 = False): reported next to the default
    eng.stage_float32 = False
    try:
        with contextlib.redirect_stdout(sys.stderr):
            e2e64_ms, _, _, _, _ = timed(step_e2e, max(1, min(args.steps, 3)), 1, track=False)
        b0 = eng.h2d_bytes
        with contextlib.redirect_stdout(sys.stderr):
            step_e2e()
        torch.cuda.synchronize(dev)
        h2d_wire64 = eng.h2d_bytes - b0
    finally:
        del eng.stage_float32
    e2e64_steps = max(1, min(args.steps, 3))
    # certificate = one coalesced read of the whole cost matrix: the row-scan bandwidth probe, and a
    # full-size optimality proof of the last solve
    for _ in range(3):
        cert = eng.lap_check(cost, res)
    check_ms = eng.last_ms("check")
    cfg5 = None
    if not strong and not args.no_cfg5:
        units.clear()
        del sc_dev, st_dev
        torch.cuda.empty_cache()
        cfg5 = strong_cfg5(eng, dev, rank, world, args)

    if rank == 0:
        hbm, tf_burst, tf_sus, peak_src = measured_peaks()
        ms_per_step = tot_ms / args.steps
        total_cells = (n_chunks if strong else world) * n_cells
        value = total_cells / (ms_per_step / 1e3)
        e2e_value = total_cells / (e2e_ms / args.steps / 1e3)
        n_per, n_obj = int(res.person_obj.numel()), int(res.price.numel())
        # row scans of one solve: every auction bid scans one row (n_obj int32), every relaxed row of a search is
        # one scan, every phase start re-checks every row; + the min/max pass over every row
        scans = res.row_scans + n_per
        lap_bytes = scans * n_obj * 4
        ach = lap_bytes / (lap_ms / 1e3) / 1e9
        kop = n_genes if args.precision == "f16" else 3 * n_genes
        gemm_flop_alg = 2.0 * n_spots * n_cells * n_genes
        scan_bytes = n_per * n_obj * 4
        traffic = ncu_traffic(args.workload, "lap_sap_kernel")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "int32/int64 LAP; fp16(hi/lo split)->fp32 cost GEMM",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "n_cells": n_cells, "n_spots": n_spots,
                       "n_genes": n_genes, "cells_per_spot": cps, "precision": args.precision,
                       "distance_metric": args.distance_metric, "input_dtype": "float64",
                       "l2": "inputs larger than L2 (expression matrices %.1f GB, cost matrix %.2f GB per sub-problem)"
                             % ((sc_np.nbytes + st_np.nbytes) / 1e9, n_spots * n_cells * 4 / 1e9),
                       "per_rank": (f"{n_chunks} independent sub-LAPs dealt round-robin to {world} rank(s)" if strong else
                                    ("each rank solves its own independent copy of the same sub-problem instance" if world > 1 else "single GPU")),
                       "e2e": "cytospace_b200.solve_linear_assignment_problem(sc_np, st_np, cn, 'lapjv_b200', ...) with pageable "
                              "host numpy arrays (native pinned ring inside the call; the float64 matrices cross PCIe as float32, "
                              "narrowed by the staging threads: h2d_bytes_per_step counts the bytes on the wire, "
                              "host_bytes_per_step the arrays handed over), Python list out"},
            "roofline": {"kernel": "lap_sap_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                         "frac": ach / hbm, "traffic": traffic, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                         "row_scans": scans, "bytes_per_scan": n_obj * 4, "kernel_ms": lap_ms,
                         "algorithmic_bytes": lap_bytes,
                         "traffic_over_algorithmic": (traffic / lap_bytes if traffic else None),
                         "note": "dependency-bound: ~%d dependent grid rounds per solve; see DESIGN.md 4.3"
                                 % (int(res.stats["rounds"]) + int(res.stats["list_hits"]))},
            "roofline_row_scan": {"kernel": ("lap_rowcheck_whole_kernel" if n_obj <= 12288 else "lap_rowmin_kernel") +
                                            " (certificate: one pass over the cost matrix)",
                                  "bound": "hbm", "bytes": scan_bytes, "ms": check_ms,
                                  "achieved": scan_bytes / (check_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                                  "frac": scan_bytes / (check_ms / 1e3) / 1e9 / hbm,
                                  "note": "ms covers the whole cyb_lap_check_i32 call (memset + row pass + publish kernels)"},
            "roofline_cost_build": {"bound": "tensor", "algorithmic_flop": gemm_flop_alg, "ms": cost_ms,
                                    "achieved": gemm_flop_alg / (cost_ms / 1e3) / 1e12, "peak": tf_burst,
                                    "unit": "TFLOP/s", "frac": gemm_flop_alg / (cost_ms / 1e3) / 1e12 / tf_burst,
                                    "executed_flop": 2.0 * n_spots * n_cells * kop,
                                    "executed_frac": 2.0 * n_spots * n_cells * kop / (cost_ms / 1e3) / 1e12 / tf_burst,
                                    "note": "ms covers standardise pre-pass + GEMM; f16x3 executes 3x the algorithmic flop"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_wire, "d2h_bytes_per_step": d2h,
                    "host_bytes_per_step": h2d, "ms_per_step": e2e_ms / args.steps,
                    "float64_on_the_wire": {"value": total_cells / (e2e64_ms / e2e64_steps / 1e3), "unit": UNIT,
                                            "ms_per_step": e2e64_ms / e2e64_steps, "h2d_bytes_per_step": h2d_wire64}},
            "gpu_launches": launches,
            "clocks": clocks,
            "certificate": cert,
            "lap_stats": {"phases": res.stats["phases"], "auction_rounds": res.stats["rounds"], "auction_bids": res.stats["bids"],
                          "searches": res.stats["tails"], "search_rounds": res.stats["list_hits"],
                          "search_rows": res.stats["tail_bids"], "paths": res.stats["paths"],
                          "grid": res.stats["grid"], "smem_prices": res.stats["smem_prices"], "variant": res.stats["tail_mode"]},
            "total_cost": res.total, "lap_ms": lap_ms, "cost_build_ms": cost_ms,
        }
        if cfg5 is not None:
            line["strong_cfg5"] = cfg5
        if world == 1 and not strong and not args.no_cpu_baseline and args.distance_metric == "Pearson_correlation":
            row_map = None if cps == 1 else np.repeat(np.arange(n_spots, dtype=np.int32), cn)
            # the oracle takes the reference's orientation (spots x cells)
            cost_np = np.ascontiguousarray((cost[:, :n_cells] if cps == 1 else cost[:, :n_spots].T).cpu().numpy())
            line["cpu_baseline"] = cpu_baseline(cost_np, row_map, sc_np, st_np, wl, res.total, args.cpu_scipy)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries chat on stdout (NCCL prints "NCCL version ..." at communicator creation): everything except
    # the JSON line goes to stderr, at the file-descriptor level so that native code is covered too.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f16x3", choices=["f16", "f16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-scipy", action="store_true",
                    help="also time scipy.optimize.linear_sum_assignment on the GPU-built matrix (n <= 10k; ~90 s)")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the strong_cfg5 object (200k x 200k chunked problem)")
    ap.add_argument("--cfg5-cells", type=int, default=0, help="total cells of the strong_cfg5 problem (default 200000)")
    ap.add_argument("--distance-metric", default="Pearson_correlation",
                    choices=["Pearson_correlation", "Spearman_correlation", "Euclidean"])
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
