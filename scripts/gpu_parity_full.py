"""Full-size oracle parity of one BASELINE config: the GPU-built integer cost matrix is solved on the device and by
the CPU oracle (restated JV, oracle/lapjv_oracle.c; capacitated configs through `row_map` = location_repeat) and the
two totals are compared.  Writes profiles/r02_parity_<name>.json.  CPU time: ~1 min at 25k, ~4 min at 30k x 5k,
~10 min at 50k (single thread; the oracle is the checker here, not the thing measured).

`--recheck`: after a change of the solver's schedule, solve the same seeded instance again and compare the device total
with the oracle total RECORDED in profiles/r02_parity_<name>.json (the JV run is not repeated; the on-device certificate
max_violation <= 1 is itself a proof of optimality) -> profiles/r02_parity_recheck_<name>.json."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine

CFG = {"cfg2": (10000, 10000, 20000, 1, 1002), "cfg3": (50000, 50000, 20000, 1, 1003), "cfg4": (30000, 5000, 30000, 6, 1004),
       "cfg5_chunk0": (25000, 25000, 20000, 1, 1005)}
name = sys.argv[1]
recheck = "--recheck" in sys.argv[2:]
N, S, G, cps, seed = CFG[name]
eng = AssignmentEngine()
eng.profile = True
sc, st, cn = syn.structured_counts_torch(N, S, G, cps, seed=seed, device=eng.device)
sc, st = syn.normalize_data_torch(sc), syn.normalize_data_torch(st)
spot, res, cost = eng.assign(sc, st, cn)
lap_ms = eng.last_ms("lap")
cert = eng.lap_check(cost, res)
if recheck:
    rec = json.load(open(os.path.join(ROOT, "profiles", f"r02_parity_{name}.json")))
    out = {"config": name, "n_cells": N, "n_spots": S, "n_genes": G, "cells_per_spot": cps, "total_gpu": int(res.total),
           "total_cpu_oracle_recorded": rec["total_cpu_oracle"], "total_equal": bool(rec["total_cpu_oracle"] == int(res.total)),
           "certificate": cert, "gpu_lap_ms": lap_ms, "lap_stats": {k: int(v) for k, v in res.stats.items()},
           "cost_matrix_sum": int(cost[:, :(N if cps == 1 else S)].sum(dtype=torch.int64).item()),
           "note": "same seeded instance as profiles/r02_parity_%s.json; oracle total taken from there" % name}
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_parity_recheck_{name}.json"), "w"), indent=1)
    print(json.dumps(out))
    sys.exit(0)
if cps == 1:
    cost_np = np.ascontiguousarray(cost[:, :N].cpu().numpy()); row_map = None          # spots x cells
else:
    cost_np = np.ascontiguousarray(cost[:, :S].T.cpu().numpy()); row_map = np.repeat(np.arange(S, dtype=np.int32), cn)
del sc, st, cost
torch.cuda.empty_cache()
t0 = time.perf_counter()
rowsol, colsol, (total_cpu, u, v) = oracle.lapjv_i32(cost_np, row_map)
t_cpu = time.perf_counter() - t0
spots_cpu = colsol if row_map is None else row_map[colsol]
out = {"config": name, "n_cells": N, "n_spots": S, "n_genes": G, "cells_per_spot": cps,
       "total_gpu": int(res.total), "total_cpu_oracle": int(total_cpu), "total_equal": bool(int(total_cpu) == int(res.total)),
       "same_cell_to_spot_map": bool(np.array_equal(spots_cpu, spot.cpu().numpy())),
       "certificate": cert, "gpu_lap_ms": lap_ms, "cpu_oracle_lap_s": t_cpu, "cpu_threads": 1,
       "lap_speedup_vs_cpu_oracle": t_cpu * 1e3 / lap_ms,
       "oracle": "restated Jonker-Volgenant (int32 costs, int64 duals) on the GPU-built integer matrix"}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_parity_{name}.json"), "w"), indent=1)
print(json.dumps(out))
