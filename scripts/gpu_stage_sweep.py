"""GPU: host -> device staging rate of engine.to_device (pageable numpy, 1.6 GB float64) per thread count / slab size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytospace_b200.engine import AssignmentEngine
x = np.random.default_rng(0).standard_normal((20000, 10000))
print("cores", os.cpu_count(), "GB", x.nbytes / 1e9, flush=True)
t0 = time.perf_counter(); d = torch.from_numpy(x).cuda(); torch.cuda.synchronize(); print("plain pageable .cuda(): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); d = torch.from_numpy(x).cuda(); torch.cuda.synchronize(); print("plain pageable .cuda(): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
xp = torch.from_numpy(x).pin_memory()
t0 = time.perf_counter(); d = xp.cuda(non_blocking=True); torch.cuda.synchronize(); print("pinned: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
for thr in (4, 8, 12, 16):
    for slab_mb in (16, 32, 64, 128):
        for slabs in (3, 4):
            eng = AssignmentEngine()
            eng._stage_threads = thr
            eng.STAGE_SLAB_BYTES = slab_mb << 20
            eng.STAGE_SLABS = slabs
            eng.to_device(x); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); d = eng.to_device(x); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
            print(f"threads={thr} slab={slab_mb}MB x{slabs}: {min(ts):.1f} ms = {x.nbytes / min(ts) / 1e6:.1f} GB/s", flush=True)
            del eng
