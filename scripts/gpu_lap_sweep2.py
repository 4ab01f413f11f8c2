"""GPU: sweep the eps schedule (theta, eps0_div) and tail threshold on a structured matrix."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine
n = int(sys.argv[1]); G = int(sys.argv[2]); cps = int(sys.argv[3])
thetas = [int(x) for x in sys.argv[4].split(",")]; e0s = [int(x) for x in sys.argv[5].split(",")]; tails = [int(x) for x in sys.argv[6].split(",")]
seeds = [int(x) for x in (sys.argv[7].split(",") if len(sys.argv) > 7 else ["1002"])]
dev = torch.device("cuda:0")
eng = AssignmentEngine(device=dev); eng.profile = True
for seed in seeds:
    sc, st, cn = syn.structured_counts_torch(n, n // cps, G, cps, seed=seed, device=dev)
    layout = "spots_x_cells" if cps == 1 else "cells_x_spots"
    cost = eng.cost_build(syn.normalize_data_torch(sc), syn.normalize_data_torch(st), layout=layout); del sc, st
    cap = None if cps == 1 else cn
    ref = None
    for th, e0, T in itertools.product(thetas, e0s, tails):
        os.environ.update(CYB_LAP_THETA=str(th), CYB_LAP_EPS0=str(e0), CYB_LAP_TAIL=str(T))
        res = eng.lap_solve(cost, cap, n_persons=n, n_objects=n // cps)
        res = eng.lap_solve(cost, cap, n_persons=n, n_objects=n // cps)
        ms = eng.last_ms("lap"); s = res.stats
        if ref is None: ref = res.total
        print(f"seed={seed} n={n} cps={cps} theta={th} eps0={e0} T={T}: {ms:.1f} ms ok={res.total == ref} phases={s['phases']} rounds={s['rounds']} "
              f"bids={s['bids']} tail={s['tail_bids']} hits={s['list_hits']}", flush=True)
