"""GPU: LAP-only timing sweeps on a device-built structured cost matrix (tuning aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
tails = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "8"])]
dev = torch.device("cuda:0")
eng = AssignmentEngine(device=dev); eng.profile = True
sc, st, cn = syn.structured_counts_torch(n, n, G, 1, seed=1019, device=dev)
cost = eng.cost_build(syn.normalize_data_torch(sc), syn.normalize_data_torch(st)); del sc, st
ref = None
for T in tails:
    os.environ["CYB_LAP_TAIL"] = str(T)
    for rep in range(2):
        res = eng.lap_solve(cost, n_persons=n, n_objects=n)
    ms = eng.last_ms("lap")
    st_ = res.stats
    if ref is None: ref = res.total
    print(f"n={n} tail_t={T}: {ms:.1f} ms total_ok={res.total == ref} rounds={st_['rounds']} bids={st_['bids']} "
          f"tail_bids={st_['tail_bids']} tails={st_['tails']} phases={st_['phases']} "
          f"us/round~{1e3 * ms / max(1, st_['rounds'] + st_['tail_bids']):.2f}", flush=True)
cert = eng.lap_check(cost, res)
print("certificate", cert)
