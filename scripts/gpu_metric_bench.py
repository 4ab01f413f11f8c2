"""GPU: device time of the cost build per distance metric (and of the rank pre-pass alone)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dev = torch.device("cuda:0")
eng = AssignmentEngine(device=dev); eng.profile = True
sc, st, cn = syn.structured_counts_torch(n, n, G, 1, seed=1002, device=dev)
sc_n, st_n = syn.normalize_data_torch(sc), syn.normalize_data_torch(st); del sc, st
for metric in ("Pearson_correlation", "Spearman_correlation", "Euclidean"):
    for rep in range(3):
        cost = eng.cost_build(sc_n, st_n, layout="spots_x_cells", metric=metric)
    print(f"n={n} G={G} {metric}: cost build {eng.last_ms('cost'):.2f} ms", flush=True)
    res = eng.lap_solve(cost, None, n_persons=n, n_objects=n); res = eng.lap_solve(cost, None, n_persons=n, n_objects=n)
    print(f"   LAP {eng.last_ms('lap'):.1f} ms total={res.total} rounds={res.stats['rounds']} bids={res.stats['bids']} tail={res.stats['tail_bids']}", flush=True)
for rep in range(3):
    r = eng.rank_columns(sc_n)
ms = eng.last_ms("rank")
print(f"rank_columns [{G} x {n}] f64: {ms:.2f} ms = {G * n * 12 / ms / 1e6:.0f} GB/s algorithmic (8 B in + 4 B out per element)")
