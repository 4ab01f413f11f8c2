"""GPU: host -> device rate of engine.to_device (pageable numpy, 1.6 GB float64) through cyb_stage_upload per thread
count / piece size / ring depth (one process per setting: the stager is sized once per process)."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
from cytospace_b200.engine import AssignmentEngine
x = np.random.default_rng(0).standard_normal((20000, 10000))
eng = AssignmentEngine()
if os.environ.get("PLAIN"):
    t0 = time.perf_counter(); d = torch.from_numpy(x).cuda(); torch.cuda.synchronize(); t1 = time.perf_counter()
    d = torch.from_numpy(x).cuda(); torch.cuda.synchronize(); t2 = time.perf_counter()
    xp = torch.from_numpy(x).pin_memory(); t3 = time.perf_counter(); d = xp.cuda(non_blocking=True); torch.cuda.synchronize()
    print("cores %%d  plain pageable .cuda(): %%.1f / %%.1f ms   pinned: %%.1f ms" %% (os.cpu_count(), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t3) * 1e3))
else:
    for narrow in (False, True):
        eng.stage_float32 = narrow
        eng.to_device(x); torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); d = eng.to_device(x); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        assert torch.equal(d[::997].cpu(), torch.from_numpy(x[::997].astype(np.float32 if narrow else np.float64)))
        print("threads=%%s piece=%%sMB pieces=%%s %%s: %%.1f ms = %%.1f GB/s of host float64" %% (os.environ.get("CYB_STAGE_THREADS"), os.environ.get("CYB_STAGE_PIECE_MB"), os.environ.get("CYB_STAGE_PIECES"), "float32 on the wire" if narrow else "float64 on the wire", min(ts), x.nbytes / min(ts) / 1e6))
""" % ROOT


def run(**env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=300)
    print((r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)


if __name__ == "__main__":
    run(PLAIN=1)
    for thr in (4, 8, 12, 16):
        for piece in (4, 8):
            run(CYB_STAGE_THREADS=thr, CYB_STAGE_PIECE_MB=piece, CYB_STAGE_PIECES=3 * thr)
