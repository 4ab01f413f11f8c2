"""GPU: LAP timing of the default solver on the bench workloads (cost matrix built on the device)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine
eng = AssignmentEngine()
CFG = {"4k": (4000, 4000, 5000, 1, 1002), "cfg2": (10000, 10000, 20000, 1, 1002), "cfg4": (30000, 5000, 30000, 6, 1004),
       "25k": (25000, 25000, 20000, 1, 1005), "cfg3": (50000, 50000, 20000, 1, 1003)}
names = sys.argv[1].split(",")
envs = [dict(kv.split("=") for kv in a.split(",")) if a else {} for a in (sys.argv[2:] or [""])]
for name in names:
    N, S, G, cps, seed = CFG[name]
    scd, std, cn = syn.structured_counts_torch(N, S, G, cps, seed=seed, device=eng.device)
    scd, std = syn.normalize_data_torch(scd), syn.normalize_data_torch(std)
    for env in envs:
        os.environ.update(env)
        eng.profile = True
        ms = []
        for rep in range(3):
            spot, res, costm = eng.assign(scd, std, cn)
            ms.append(eng.last_ms("lap"))
        s = res.stats
        keys = ("phases", "rounds", "bids", "tails", "list_hits", "tail_bids", "paths", "ns_tail", "ns_select", "ns_relax", "ns_augment", "ns_sel_pass", "ns_sel_scan", "ns_bid", "ns_barrier", "ns_resolve", "small_rounds", "ns_phase_start", "ns_auction", "ns_init", "ns_total")
        print(name, env, "lap ms", [round(x, 2) for x in ms], "total", res.total, {k: s.get(k) for k in keys}, flush=True)
        for k in env: os.environ.pop(k, None)
    del scd, std
