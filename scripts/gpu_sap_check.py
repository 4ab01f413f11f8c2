"""GPU: first checks of the SAP-finish LAP solver -- oracle totals, CPU-model assignments, timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine

eng = AssignmentEngine()

def to_dev(m):
    ld = (m.shape[1] + 31) // 32 * 32
    dev = torch.full((m.shape[0], ld), 2 ** 30 - 1, dtype=torch.int32, device=eng.device)
    dev[:, :m.shape[1]] = torch.from_numpy(np.ascontiguousarray(m)).to(eng.device)
    return dev

def run(m, cap=None, label="", model=True, jv=True):
    dev = to_dev(m)
    res = eng.lap_solve(dev, cap, n_persons=m.shape[0], n_objects=m.shape[1])
    po = res.person_obj.cpu().numpy()
    cert = eng.lap_check(dev, res)
    msg = f"{label}: total={res.total} cert={cert['max_violation']},{cert['invalid_rows']},{cert['capacity_mismatch']}"
    if jv:
        rm = None if cap is None else np.repeat(np.arange(m.shape[1], dtype=np.int32), cap)
        tot = oracle.lapjv_i32(np.ascontiguousarray(m.T) if cap is not None else m, rm)[2][0]
        msg += f" jv_ok={tot == res.total}"
    if model:
        pm, sm, tm, _, st, _ = oracle.sap_model(m, cap, theta=8, sap_t=64, K=296, multi=16, partial=64, warm=0 if (os.environ.get('CYB_LAP_SMEM_OWNER') == '0' or os.environ.get('CYB_LAP_SMEM_PRICES') == '0') else 1)      # the device defaults
        msg += f" model_total_ok={tm == res.total} model_assign_ok={np.array_equal(pm, po)} model(srounds={st[4]},srows={st[5]},searches={st[3]},paths={st[6]})"
    s = res.stats
    msg += f" dev(phases={s['phases']},rounds={s['rounds']},bids={s['bids']},searches={s['tails']},srounds={s['list_hits']},srows={s['tail_bids']},paths={s['paths']})"
    print(msg, flush=True)
    return res

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
rng = np.random.default_rng(3)
for n in ((5, 64, 300, 1000) if mode == "small" else ()):
    run(rng.integers(-1000, 1000, (n, n), dtype=np.int32), label=f"uniform{n}")
if mode != "small":
    def run(*a, **k): pass
run(rng.integers(0, 50, (200, 200), dtype=np.int32), label="ties200")
cap = rng.integers(0, 6, 70).astype(np.int32)
run(rng.integers(-1000, 1000, (int(cap.sum()), 70), dtype=np.int32), cap, label="cap70")
cap = rng.integers(0, 7, 400).astype(np.int32)
run(rng.integers(-300000, 300000, (int(cap.sum()), 400), dtype=np.int32), cap, label="cap400")
for env in ({"CYB_LAP_SMEM_OWNER": "0"}, {"CYB_LAP_SMEM_PRICES": "0"}):
    os.environ.update(env)
    run(rng.integers(-1000, 1000, (500, 500), dtype=np.int32), label=f"uniform500 {env}")
    cap = rng.integers(0, 6, 120).astype(np.int32)
    run(rng.integers(-1000, 1000, (int(cap.sum()), 120), dtype=np.int32), cap, label=f"cap120 {env}")
    for k in env: del os.environ[k]

# structured instances: correctness at 2k, timing at larger sizes
from oracle import cost_oracle as co
if mode != "small":
    sys.argv = sys.argv[:1]
sc, st, cn = syn.structured_counts(2000, 2000, 3000, 1, seed=1002)
cost = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))
run(cost, label="struct2k")
sc, st, cn = syn.structured_counts(3000, 500, 3000, 6, seed=1004)
cost = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))
run(np.ascontiguousarray(cost.T), cn.astype(np.int32), label="structcap3k")

def timed(N, S, G, cps, seed, label):
    scd, std, cn = syn.structured_counts_torch(N, S, G, cps, seed=seed, device=eng.device)
    scd, std = syn.normalize_data_torch(scd), syn.normalize_data_torch(std)
    for solver in ("sap", "auction"):
        if solver == "auction": os.environ["CYB_LAP_SOLVER"] = "auction"
        eng.profile = True
        for rep in range(2):
            spot, res, costm = eng.assign(scd, std, cn)
            ms = eng.last_ms("lap")
        cert = eng.lap_check(costm, res)
        s = res.stats
        print(f"{label} [{solver}]: lap {ms:.2f} ms total={res.total} cert={cert['max_violation']} stats={ {k: s[k] for k in s} }", flush=True)
        os.environ.pop("CYB_LAP_SOLVER", None)

if mode == "small":
    sys.exit(0)
timed(4000, 4000, 5000, 1, 1002, "4k")
timed(10000, 10000, 20000, 1, 1002, "cfg2")
timed(30000, 5000, 30000, 6, 1004, "cfg4")
timed(25000, 25000, 20000, 1, 1005, "25k")
