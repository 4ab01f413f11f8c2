"""GPU diagnostics: cost-build error vs the float64 oracle for several G, and kernel timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine

dev = torch.device("cuda:0")
for G in (2000, 20000, 30000):
    n = 1500
    sc, st, cn = syn.structured_counts_torch(n, n, G, 1, seed=5, device=dev)
    sc_n, st_n = syn.normalize_data_torch(sc), syn.normalize_data_torch(st)
    # float64 oracle on the device (torch fp64 matmul) -- diagnostics only
    zs = (st_n - st_n.mean(0)) / st_n.std(0, unbiased=False); zc = (sc_n - sc_n.mean(0)) / sc_n.std(0, unbiased=False)
    want = torch.round(-(zs.T @ zc) / G * 1e6)
    for prec in ("f16", "f16x3"):
        eng = AssignmentEngine(device=dev, precision=prec)
        got = eng.cost_build(sc_n, st_n)[:, :n].double()
        d = (got - want)
        rel = d / want.abs().clamp(min=1)
        print(f"G={G} {prec}: max|d|={d.abs().max().item():.0f} mean d={d.mean().item():.2f} rms={d.pow(2).mean().sqrt().item():.2f} "
              f"frac<=1={(d.abs()<=1).double().mean().item():.4f} mean rel={rel.mean().item():.2e}", flush=True)

# timings at cfg2 size
n, G = 10000, 20000
sc, st, cn = syn.structured_counts_torch(n, n, G, 1, seed=7, device=dev)
sc_n, st_n = syn.normalize_data_torch(sc), syn.normalize_data_torch(st); del sc, st
for prec in ("f16", "f16x3"):
    eng = AssignmentEngine(device=dev, precision=prec); eng.profile = True
    for _ in range(3):
        cost = eng.cost_build(sc_n, st_n)
    ms = eng.last_ms("cost")
    k = G if prec == "f16" else 3 * G
    print(f"cost build {prec} {n}x{n}x{G}: {ms:.2f} ms  -> executed {2*n*n*k/ms/1e9:.0f} TFLOP/s incl. standardise", flush=True)
    # GEMM alone
    lib, ffi = eng.lib, eng.ffi
    from cytospace_b200 import _native
    kop = lib.cyb_operand_k(G, 0 if prec == "f16" else 1)
    za = torch.randn((n, kop), device=dev, dtype=torch.float16); zb = torch.randn((n, kop), device=dev, dtype=torch.float16)
    out = torch.empty((n, n), dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(4):
        if it == 1: e0.record()
        _native.check(lib.cyb_cost_gemm_i32(_native.ptr("void *", za), _native.ptr("void *", zb), n, n, kop, 1.0,
                                            _native.ptr("int32_t *", out), n, eng._stream()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"GEMM alone {prec}: {ms:.2f} ms = {2*n*n*kop/ms/1e9:.0f} TFLOP/s", flush=True)
    ref = (za[:256].float() @ zb[:512].float().T)
    print("   gemm check max rel err", ((out[:256, :512].float() + ref).abs().max() / ref.abs().max()).item())
    del za, zb, out
eng = AssignmentEngine(device=dev); eng.profile = True
t0 = time.time(); spot, res, cost = eng.assign(sc_n, st_n, cn); torch.cuda.synchronize()
print("assign cfg2: wall", time.time() - t0, "lap ms", eng.last_ms("lap"), "cost ms", eng.last_ms("cost"), res.stats, flush=True)
cert = eng.lap_check(cost, res)
print("certificate", cert)
