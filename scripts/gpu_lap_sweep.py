"""GPU: LAP-only timing sweeps on a device-built structured cost matrix (tuning aid).
usage: gpu_lap_sweep.py n G tails(comma) lists(comma) [cps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
tails = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "8"])]
lists = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["1"])]
cps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda:0")
eng = AssignmentEngine(device=dev); eng.profile = True
sc, st, cn = syn.structured_counts_torch(n, n // cps, G, cps, seed=1002, device=dev)
layout = "spots_x_cells" if cps == 1 else "cells_x_spots"
cost = eng.cost_build(syn.normalize_data_torch(sc), syn.normalize_data_torch(st), layout=layout); del sc, st
cap = None if cps == 1 else cn
ref = None
for L in lists:
    for T in tails:
        os.environ["CYB_LAP_TAIL"] = str(T); os.environ["CYB_LAP_LISTS"] = str(L)
        for rep in range(2):
            res = eng.lap_solve(cost, cap, n_persons=n, n_objects=n // cps)
        ms = eng.last_ms("lap")
        s = res.stats
        if ref is None: ref = (res.total, res.person_obj.clone())
        same = bool((res.person_obj == ref[1]).all())
        print(f"n={n} cps={cps} lists={L} tail_t={T}: {ms:.1f} ms total_ok={res.total == ref[0]} same_assignment={same} rounds={s['rounds']} "
              f"bids={s['bids']} tail_bids={s['tail_bids']} list_hits={s['list_hits']} phases={s['phases']}", flush=True)
cert = eng.lap_check(cost, res)
print("certificate", cert)
