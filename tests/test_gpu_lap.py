"""GPU: the LAP kernel (through the C ABI) against the CPU oracle.  Bit-exact bar: identical total
cost; the assignment must be valid and optimal (certificate: eps-complementary slackness <= 1
scaled unit, checked on the device), and the run deterministic."""
import numpy as np
import pytest
import torch

import oracle
from cytospace_b200 import synthetic as syn
from conftest import LAP_NAMES

pytestmark = pytest.mark.gpu


def to_dev(engine, m):
    ld = (m.shape[1] + 31) // 32 * 32
    dev = torch.full((m.shape[0], ld), 2 ** 30 - 1, dtype=torch.int32, device=engine.device)
    dev[:, :m.shape[1]] = torch.from_numpy(np.ascontiguousarray(m)).to(engine.device)
    return dev


def solve_and_check(engine, m, cap=None, grid=0):
    """m: persons x objects.  Returns (LapResult, person_obj ndarray)."""
    n_p, n_o = m.shape
    dev = to_dev(engine, m)
    res = engine.lap_solve(dev, cap, n_persons=n_p, n_objects=n_o, grid=grid)
    po = res.person_obj.cpu().numpy(); so = res.slot_owner.cpu().numpy()
    capv = np.ones(n_o, np.int64) if cap is None else np.asarray(cap)
    assert po.min() >= 0 and po.max() < n_o
    assert np.array_equal(np.bincount(po, minlength=n_o), capv), "capacities violated"
    soff = np.concatenate([[0], np.cumsum(capv)])
    assert sorted(so.tolist()) == list(range(n_p)), "slot_owner is not a permutation of the persons"
    assert np.array_equal(np.repeat(np.arange(n_o), capv)[np.argsort(so)], po), "slot_owner inconsistent with person_obj"
    assert int(m[np.arange(n_p), po].astype(np.int64).sum()) == res.total
    cert = engine.lap_check(dev, res)
    assert cert["invalid_rows"] == 0 and cert["capacity_mismatch"] == 0 and cert["total"] == res.total
    assert cert["max_violation"] <= 1, cert
    return res, po


@pytest.mark.parametrize("name", LAP_NAMES)
def test_golden_instances(engine, lap_golden, name):
    cost = lap_golden[f"{name}_cost"]
    res, po = solve_and_check(engine, cost)
    assert res.total == int(lap_golden[f"{name}_opt"])
    res_t, _ = solve_and_check(engine, cost.T)               # either side may bid
    assert res_t.total == res.total


@pytest.mark.parametrize("n,high,seed", [(1, 10, 0), (2, 10, 1), (3, 5, 2), (31, 100, 3), (33, 2_000_000, 4),
                                          (127, 50, 5), (257, 2_000_000, 6), (1000, 2_000_000, 7),
                                          (1023, 1000, 8), (2050, 2_000_000, 9)])
def test_uniform_random_vs_oracle(engine, n, high, seed):
    cost = np.random.default_rng(seed).integers(-high, high, (n, n), dtype=np.int32)
    res, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]


@pytest.mark.parametrize("grid", [1, 2, 7, 148])
def test_result_independent_of_grid_size(engine, grid):
    cost = np.random.default_rng(11).integers(0, 10_000, (300, 300), dtype=np.int32)
    ref, po_ref = solve_and_check(engine, cost, grid=0)
    res, po = solve_and_check(engine, cost, grid=grid)
    assert res.total == ref.total and np.array_equal(po, po_ref)      # deterministic tie-breaks


def _with_env(env, fn):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


SAP_KNOBS = [dict(theta=8, sap_t=64, K=296, multi=16, partial=64),       # the defaults (lap_sap.cu): incomplete phases, searches only at eps = 1
             dict(theta=64, sap_t=64, K=296, multi=16, partial=0),       # every phase finished by searches (the schedule before incomplete phases)
             dict(theta=64, sap_t=148, K=296, multi=32, partial=0),
             dict(theta=4, sap_t=8, K=16, multi=1, partial=0),            # one path per search, tiny rounds (threshold histogram in use)
             dict(theta=256, sap_t=40, K=64, multi=8, partial=0),
             dict(theta=16, sap_t=256, K=100000, multi=32, partial=0),    # every dirty object each round (Bellman-Ford rounds)
             dict(theta=4, sap_t=16, K=296, multi=16, partial=16),        # incomplete phases with a smaller hand-over
             dict(theta=8, sap_t=32, K=296, multi=16, partial=12)]        # ... and phases that stop in the middle of their searches


def _sap_env(k):
    return {"CYB_LAP_THETA": k["theta"], "CYB_LAP_SAP_T": k["sap_t"], "CYB_LAP_SAP_K": k["K"], "CYB_LAP_SAP_MULTI": k["multi"],
            "CYB_LAP_PARTIAL": k["partial"]}


@pytest.mark.parametrize("knobs", SAP_KNOBS, ids=lambda k: "t{sap_t}_k{K}_m{multi}_th{theta}_p{partial}".format(**k))
def test_matches_cpu_model_of_the_device_algorithm(engine, knobs):
    """Same rules as oracle/sap_model.c (auction rounds + shortest-augmenting-path finish: frontier threshold,
    strict relaxations, (label, slot) tie-breaks, path claims): identical assignment, identical counters --
    not just an identical total."""
    rng = np.random.default_rng(3)
    cap = rng.integers(0, 6, 70).astype(np.int32)
    m = rng.integers(-1000, 1000, (int(cap.sum()), 70), dtype=np.int32)
    cases = [(m, cap), (rng.integers(0, 50, (90, 90), dtype=np.int32), None),          # unit capacities, many ties
             (rng.integers(-100_000, 100_000, (700, 700), dtype=np.int32), None)]
    # the default schedule and the complete-phases schedule on every memory variant, the others on the default variant
    variants = MEMORY_VARIANTS if knobs in SAP_KNOBS[:2] else MEMORY_VARIANTS[:1]
    for (mat, cp), (sp, so_) in [(c, v) for c in cases for v in variants]:
        env = dict(_sap_env(knobs), CYB_LAP_SMEM_PRICES=sp, CYB_LAP_SMEM_OWNER=so_, CYB_LAP_WARM=1)
        res, po = _with_env(env, lambda: solve_and_check(engine, mat, cp))
        # searches are warm-started from the previous one's forest where the predecessors live in shared memory
        po_model, so_model, tot_model, lam_model, st, _ = oracle.sap_model(mat, cp, warm=int(sp and so_), **knobs)
        assert res.total == tot_model and np.array_equal(po, po_model)
        assert np.array_equal(res.slot_owner.cpu().numpy(), so_model)
        assert np.array_equal(res.price.cpu().numpy(), lam_model)
        assert (res.stats["tails"], res.stats["list_hits"], res.stats["tail_bids"], res.stats["paths"]) == \
            (st[3], st[4], st[5], st[6]), "searches / search rounds / rows relaxed / paths differ from the model"


@pytest.mark.parametrize("knobs", SAP_KNOBS[1:], ids=lambda k: "t{sap_t}_k{K}_m{multi}_th{theta}_p{partial}".format(**k))
def test_search_knobs_reach_the_same_optimum(engine, knobs):
    sc, st, cn = syn.structured_counts(1200, 200, 800, 6, seed=1004)
    from oracle import cost_oracle as co
    compact = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))
    m = np.ascontiguousarray(compact.T)
    res, _ = _with_env(_sap_env(knobs), lambda: solve_and_check(engine, m, cn.astype(np.int32)))
    assert res.total == oracle.lapjv_i32(compact, np.repeat(np.arange(200, dtype=np.int32), 6))[2][0]
    sq = syn.uniform_cost_i32(700, seed=3)
    res, _ = _with_env(_sap_env(knobs), lambda: solve_and_check(engine, sq))
    assert res.total == oracle.lapjv_i32(sq)[2][0]


@pytest.mark.parametrize("n_obj,max_cap,seed", [(1, 7, 0), (5, 4, 1), (40, 6, 2), (333, 9, 3), (1000, 3, 4)])
def test_capacitated_vs_expanded_oracle(engine, n_obj, max_cap, seed):
    """Spots with cn cells each (incl. empty spots): optimum equals JV on the expanded matrix (LAS:63-66)."""
    rng = np.random.default_rng(seed)
    cap = rng.integers(0, max_cap + 1, n_obj).astype(np.int32)
    cap[rng.integers(0, n_obj)] += 1
    n = int(cap.sum())
    compact = rng.integers(-50_000, 50_000, (n_obj, n), dtype=np.int32)        # spots x cells
    res, po = solve_and_check(engine, np.ascontiguousarray(compact.T), cap)
    row_map = np.repeat(np.arange(n_obj, dtype=np.int32), cap)
    assert res.total == oracle.lapjv_i32(compact, row_map)[2][0]


def test_visium_like_structured(engine):
    """cfg4-shaped: S spots x 6 cells per spot on a correlation-built matrix."""
    from oracle import cost_oracle as co
    sc, st, cn = syn.structured_counts(1200, 200, 800, 6, seed=1004)
    compact = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))       # spots x cells
    res, _ = solve_and_check(engine, np.ascontiguousarray(compact.T), cn.astype(np.int32))
    row_map = np.repeat(np.arange(200, dtype=np.int32), 6)
    assert res.total == oracle.lapjv_i32(compact, row_map)[2][0]


def test_degenerate_inputs(engine):
    for cost in (np.zeros((64, 64), np.int32), np.full((5, 5), -7, np.int32),
                 np.tile(np.arange(40, dtype=np.int32), (40, 1)),            # identical rows
                 np.tile(np.arange(40, dtype=np.int32)[:, None], (1, 40)),   # identical columns
                 np.random.default_rng(1).integers(0, 2, (200, 200)).astype(np.int32)):
        res, _ = solve_and_check(engine, cost)
        assert res.total == oracle.lapjv_i32(cost)[2][0]
    # one spot takes everything / duplicated cells (up-sampled, cytospace.py:278-281)
    m = np.random.default_rng(2).integers(0, 100, (30, 1), dtype=np.int32)
    res, po = solve_and_check(engine, m, np.array([30], np.int32))
    assert res.total == int(m.sum()) and (po == 0).all()
    base = np.random.default_rng(3).integers(-1000, 1000, (20, 60), dtype=np.int32)
    dup = np.repeat(base, 3, axis=0)                                          # 60 persons, 3 copies each
    res, _ = solve_and_check(engine, dup)
    assert res.total == oracle.lapjv_i32(dup)[2][0]


def test_extreme_cost_range(engine):
    lim = 2 ** 30 - 1
    cost = np.random.default_rng(5).integers(-lim, lim, (96, 96), dtype=np.int64).astype(np.int32)
    res, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]


def test_structured_cfg1_vs_oracle(engine):
    """BASELINE config 1 (1k x 1k x 2k genes): LAP on the oracle-built integer matrix."""
    from oracle import cost_oracle as co
    sc, st, cn = syn.structured_counts(1000, 1000, 2000, 1, seed=1001)
    cost = co.cost_matrix_i32(co.normalize_data(sc), co.normalize_data(st))
    res, _ = solve_and_check(engine, np.ascontiguousarray(cost.T))
    assert res.total == oracle.lapjv_i32(cost)[2][0]


def test_large_lap_properties_4k(engine):
    """Size-independent properties at a size the oracle still finishes quickly."""
    cost = syn.uniform_cost_i32(4096, seed=21)
    res, _ = solve_and_check(engine, cost)
    assert res.total == oracle.lapjv_i32(cost)[2][0]
    # linearity: adding a row potential / column potential shifts the optimum by a known constant
    a = np.random.default_rng(1).integers(-1000, 1000, 4096).astype(np.int32)
    b = np.random.default_rng(2).integers(-1000, 1000, 4096).astype(np.int32)
    shifted = cost + a[:, None] + b[None, :]
    res2, _ = solve_and_check(engine, shifted)
    assert res2.total == res.total + int(a.sum()) + int(b.sum())


def test_lap_errors(engine):
    with pytest.raises(ValueError):
        engine.lap_solve(torch.zeros((4, 32), dtype=torch.float32, device=engine.device))
    with pytest.raises(ValueError):
        engine.lap_solve(torch.zeros((4, 32), dtype=torch.int32, device=engine.device), n_persons=40)
    with pytest.raises(ValueError, match="square"):
        engine.lap_solve(torch.zeros((4, 32), dtype=torch.int32, device=engine.device), np.array([1, 1]),
                         n_persons=4, n_objects=2)


def test_certificate_detects_wrong_assignments(engine):
    """The certificate is a checker, not a rubber stamp: a swapped pair, an invalid object and a broken
    capacity are each reported."""
    cost = np.random.default_rng(8).integers(0, 1_000_000, (500, 500), dtype=np.int32)
    dev = to_dev(engine, cost)
    res = engine.lap_solve(dev, None, n_persons=500, n_objects=500)
    good = engine.lap_check(dev, res)
    assert good == {"max_violation": good["max_violation"], "total": res.total, "invalid_rows": 0, "capacity_mismatch": 0}
    assert good["max_violation"] <= 1
    po = res.person_obj.clone()
    res.person_obj = po.clone(); res.person_obj[[3, 4]] = po[[4, 3]]              # still a permutation, not optimal
    bad = engine.lap_check(dev, res)
    assert bad["max_violation"] > 1 and bad["capacity_mismatch"] == 0 and bad["total"] != res.total
    res.person_obj = po.clone(); res.person_obj[7] = po[8]                         # object po[8] twice, po[7] never
    assert engine.lap_check(dev, res)["capacity_mismatch"] == 2
    res.person_obj = po.clone(); res.person_obj[9] = -1
    out = engine.lap_check(dev, res)
    assert out["invalid_rows"] == 1 and out["capacity_mismatch"] == 1


def test_certificate_tiled_variant_above_12288_objects(engine):
    """More than 12 288 objects: the price vector no longer fits in shared memory and the certificate
    runs tile by tile (the kernel behind the 25k / 50k row-scan numbers)."""
    n = 12800
    cost = syn.uniform_cost_i32(n, seed=5)
    dev = to_dev(engine, cost)
    res = engine.lap_solve(dev, None, n_persons=n, n_objects=n)
    po = res.person_obj.cpu().numpy()
    assert sorted(po.tolist()) == list(range(n))
    assert int(cost[np.arange(n), po].astype(np.int64).sum()) == res.total
    cert = engine.lap_check(dev, res)
    assert cert == {"max_violation": cert["max_violation"], "total": res.total, "invalid_rows": 0, "capacity_mismatch": 0}
    assert cert["max_violation"] <= 1
    keep = res.person_obj.clone()
    res.person_obj = keep.clone(); res.person_obj[[10, 11]] = keep[[11, 10]]
    bad = engine.lap_check(dev, res)
    assert bad["max_violation"] > 1 and bad["capacity_mismatch"] == 0
    res.person_obj = keep.clone(); res.person_obj[5] = keep[6]
    assert engine.lap_check(dev, res)["capacity_mismatch"] == 2


MEMORY_VARIANTS = [(1, 1), (1, 0), (0, 0)]      # (prices / search snapshot in shared memory, slot owners / predecessors in shared memory)


@pytest.mark.parametrize("smem_prices,smem_owner", MEMORY_VARIANTS)
def test_memory_variants_on_small_problems(engine, lap_golden, smem_prices, smem_owner):
    """The code paths of problems too large for shared memory, forced on small inputs: slot owners / tree
    predecessors read from global memory (25k and up: CYB_LAP_SMEM_OWNER=0), prices and the search snapshot
    streamed from L2 (50k: CYB_LAP_SMEM_PRICES=0).  Same totals AND the same assignment as the default path."""
    env = {"CYB_LAP_SMEM_PRICES": smem_prices, "CYB_LAP_SMEM_OWNER": smem_owner}
    for name in LAP_NAMES:
        cost = lap_golden[f"{name}_cost"]
        res, po = _with_env(env, lambda: solve_and_check(engine, cost))
        assert res.total == int(lap_golden[f"{name}_opt"]) and res.stats["smem_prices"] == smem_prices
        assert res.stats["tail_mode"] == 2 + smem_owner                                                  # the variant that ran
    rng = np.random.default_rng(12)
    cap = rng.integers(0, 5, 300).astype(np.int32)
    m = rng.integers(-200_000, 200_000, (int(cap.sum()), 300), dtype=np.int32)
    cold = {"CYB_LAP_WARM": 0}            # (warm-started searches exist only with shared-memory predecessors: compare like with like)
    ref, po_ref = _with_env(cold, lambda: solve_and_check(engine, m, cap))
    res, po = _with_env(dict(env, **cold), lambda: solve_and_check(engine, m, cap))
    assert res.total == ref.total and np.array_equal(po, po_ref)
    assert solve_and_check(engine, m, cap)[0].total == ref.total
    row_map = np.repeat(np.arange(300, dtype=np.int32), cap)
    assert res.total == oracle.lapjv_i32(np.ascontiguousarray(m.T), row_map)[2][0]
    sq = syn.uniform_cost_i32(1500, seed=9, high=3000)                    # many near-ties
    ref, po_ref = _with_env(cold, lambda: solve_and_check(engine, sq))
    res, po = _with_env(dict(env, **cold), lambda: solve_and_check(engine, sq))
    assert res.total == ref.total == oracle.lapjv_i32(sq)[2][0] and np.array_equal(po, po_ref)
    assert solve_and_check(engine, sq)[0].total == ref.total
    tie = np.zeros((200, 200), np.int32)                                   # every column ties
    res, _ = _with_env(env, lambda: solve_and_check(engine, tie))
    assert res.total == 0


def _unique_optimum(compact, row_map, spot_of_cell):
    """Is the optimal cell -> spot map unique?  Decided by the oracle itself: scale the costs by n + 1 and add 1 to
    every arc of the optimum found -- any other optimal map would now be strictly cheaper and JV would return it."""
    n = compact.shape[1]
    pert = compact.astype(np.int64) * (n + 1)
    pert[spot_of_cell, np.arange(n)] += 1
    assert np.abs(pert).max() < 2 ** 30
    _, colsol2, _ = oracle.lapjv_i32(pert.astype(np.int32), row_map)
    spots2 = colsol2 if row_map is None else row_map[colsol2]
    return np.array_equal(spots2, spot_of_cell)


def test_permutation_equals_jv_oracle_when_the_optimum_is_unique(engine):
    """North-star: "permutation identical up to documented tie-breaks".  Where the optimum is unique there is no
    tie to break: the device assignment must BE the JV oracle's.  (Uniqueness is decided by the oracle, see
    _unique_optimum; instances with ties -- small cost ranges -- are skipped and counted.)"""
    found = 0
    for seed in range(30):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(40, 400))
        high = int(rng.choice([50, 5000, 1_000_000]))
        cost = rng.integers(-high, high, (n, n), dtype=np.int32)               # rows = slots, columns = cells
        rowsol, colsol, (total, u, v) = oracle.lapjv_i32(cost)
        if not _unique_optimum(cost, None, colsol):
            continue
        found += 1
        res, po = solve_and_check(engine, cost)
        assert res.total == total
        assert np.array_equal(po, rowsol), "unique optimum, different permutation"
        assert np.array_equal(res.slot_owner.cpu().numpy(), colsol)
    assert found >= 10


def test_capacitated_permutation_equals_jv_oracle_when_unique(engine):
    """Same with spot capacities: the expanded problem has cn[s] identical rows per spot, so only the cell -> spot
    map can be unique; it must equal location_repeat[oracle assignment] (cytospace.py:331)."""
    found = 0
    for seed in range(30):
        rng = np.random.default_rng(2000 + seed)
        n_obj = int(rng.integers(5, 60))
        cap = rng.integers(0, 4, n_obj).astype(np.int32); cap[0] += 1
        n = int(cap.sum())
        compact = rng.integers(-1_000_000, 1_000_000, (n_obj, n), dtype=np.int32)          # spots x cells
        row_map = np.repeat(np.arange(n_obj, dtype=np.int32), cap)
        rowsol, colsol, (total, u, v) = oracle.lapjv_i32(compact, row_map)
        if not _unique_optimum(compact, row_map, row_map[colsol]):
            continue
        found += 1
        res, po = solve_and_check(engine, np.ascontiguousarray(compact.T), cap)
        assert res.total == total and np.array_equal(po, row_map[colsol])
    assert found >= 10


def test_price_overflow_is_reported_not_returned(engine):
    """|cost| up to 2^30 times (P + 1) exceeds the 46-bit bid / label field for P >~ 32k (e.g. Euclidean costs at
    the 2^30 clamp): the solve must fail loudly with the overflow status, never return a wrong 'optimum'."""
    n = 33000
    gen = torch.Generator(device=engine.device); gen.manual_seed(5)
    dev = torch.randint(0, 2 ** 30 - 1, (n, (n + 31) // 32 * 32), generator=gen, device=engine.device, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="overflow"):
        engine.lap_solve(dev, None, n_persons=n, n_objects=n)
    # a matrix wider than the documented |cost| < 2^30 contract is refused as well
    bad = torch.zeros((64, 64), dtype=torch.int32, device=engine.device)
    bad[0, 0] = 2 ** 30 + 5; bad[1, 1] = -(2 ** 30) - 5
    with pytest.raises(RuntimeError, match="overflow"):
        engine.lap_solve(bad, None, n_persons=64, n_objects=64)
