import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from cytospace_b200.engine import AssignmentEngine
eng = AssignmentEngine()
rng = np.random.default_rng(3)
for n in [int(a) for a in sys.argv[1:]] or [5]:
    m = rng.integers(-1000, 1000, (n, n), dtype=np.int32)
    ld = (n + 31) // 32 * 32
    dev = torch.full((n, ld), 2 ** 30 - 1, dtype=torch.int32, device=eng.device)
    dev[:, :n] = torch.from_numpy(m).to(eng.device)
    try:
        res = eng.lap_solve(dev, None, n_persons=n, n_objects=n)
        print(n, "total", res.total, "jv", oracle.lapjv_i32(m)[2][0], res.stats, flush=True)
    except Exception as e:
        torch.cuda.synchronize()
        print(n, "FAILED", e, flush=True)
