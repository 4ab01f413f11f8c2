"""GPU: one LAP solve on a device-built structured matrix (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine
n = int(sys.argv[1]); G = int(sys.argv[2]); cps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
eng = AssignmentEngine(device=dev); eng.profile = True
sc, st, cn = syn.structured_counts_torch(n, n // cps, G, cps, seed=1002, device=dev)
layout = "spots_x_cells" if cps == 1 else "cells_x_spots"
cost = eng.cost_build(syn.normalize_data_torch(sc), syn.normalize_data_torch(st), layout=layout); del sc, st
res = eng.lap_solve(cost, None if cps == 1 else cn, n_persons=n, n_objects=n // cps)
print(eng.last_ms("lap"), res.stats)
