"""GPU: sweep LAP tuning knobs given as ENV=v1,v2,... arguments on device-built structured matrices.
usage: gpu_lap_sweep3.py n G cps seeds(comma) CYB_LAP_EARLY=0,8,16 CYB_LAP_THETA=4,8 ..."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine
n = int(sys.argv[1]); G = int(sys.argv[2]); cps = int(sys.argv[3]); seeds = [int(x) for x in sys.argv[4].split(",")]
knobs = [a.split("=") for a in sys.argv[5:]]
names = [k for k, _ in knobs]; vals = [v.split(",") for _, v in knobs]
dev = torch.device("cuda:0")
eng = AssignmentEngine(device=dev); eng.profile = True
for seed in seeds:
    sc, st, cn = syn.structured_counts_torch(n, n // cps, G, cps, seed=seed, device=dev)
    layout = "spots_x_cells" if cps == 1 else "cells_x_spots"
    cost = eng.cost_build(syn.normalize_data_torch(sc), syn.normalize_data_torch(st), layout=layout); del sc, st
    cap = None if cps == 1 else cn
    ref = None
    for combo in itertools.product(*vals):
        os.environ.update(dict(zip(names, combo)))
        ms = []
        for rep in range(3):
            res = eng.lap_solve(cost, cap, n_persons=n, n_objects=n // cps)
            ms.append(eng.last_ms("lap"))
        s = res.stats
        if ref is None: ref = res.total
        tag = " ".join(f"{k.replace('CYB_LAP_', '').lower()}={v}" for k, v in zip(names, combo))
        print(f"seed={seed} n={n} cps={cps} {tag}: {min(ms):.1f} ms ok={res.total == ref} phases={s['phases']} rounds={s['rounds']} "
              f"bids={s['bids']} tail={s['tail_bids']} hits={s['list_hits']} | small rounds {s['small_rounds']}: bid {s['ns_bid'] / max(1, s['small_rounds']) / 1e3:.2f} "
              f"bar {s['ns_barrier'] / max(1, s['small_rounds']) / 1e3:.2f} res {s['ns_resolve'] / max(1, s['small_rounds']) / 1e3:.2f} us; tails {s['ns_tail'] / 1e6:.1f} ms", flush=True)
    cert = eng.lap_check(cost, res)
    print("certificate", cert, flush=True)
