"""GPU: the certificate / row-scan probe in isolation (10k x 10k int32, 400 MB)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytospace_b200.engine import AssignmentEngine
eng = AssignmentEngine(); eng.profile = True
for n in (10000, 12000):
    cost = torch.randint(0, 2000000, (n, (n + 31) // 32 * 32), dtype=torch.int32, device="cuda")
    res = eng.lap_solve(cost, None, n_persons=n, n_objects=n)
    ms = []
    for _ in range(6):
        c = eng.lap_check(cost, res); ms.append(eng.last_ms("check"))
    print(n, "check ms", [round(x, 4) for x in ms], "GB/s", round(n * n * 4 / min(ms) / 1e6, 1), c["max_violation"], c["total"] == res.total)
