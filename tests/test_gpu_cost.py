"""GPU: the cost build (standardise + tcgen05 GEMM, through the C ABI) against the float64 oracle
and the committed outputs of the reference's own calculate_cost.

Tolerance (floating point, stated here as the prompt requires): the integer cost is
rint(-1e6 * r); with fp16 hi/lo-split operands (f16x3) and fp32 accumulation we require
|cost_gpu - cost_oracle| <= 4 units (4e-6 in r) everywhere and <= 1 unit for 90% of entries
(measured on B200: max 2, 95-100% within 1 unit for G = 200 .. 30000);
with single fp16 operands (f16) |delta| <= 400 units (4e-4 in r; measured rms ~40-50, max ~200)."""
import numpy as np
import pytest
import torch

from cytospace_b200 import synthetic as syn
from cytospace_b200.engine import AssignmentEngine
from oracle import cost_oracle as co

pytestmark = pytest.mark.gpu

TOL = {"f16x3": 4, "f16": 400}


def build(engine, sc, st, log_tpm=False, dtype=torch.float64):
    sc_d = engine.to_device(sc, dtype); st_d = engine.to_device(st, dtype)
    cost, cs_sc, cs_st = engine.cost_build(sc_d, st_d, log_tpm=log_tpm, return_colstats=True)
    # the device matrix is cells x spots; compare in the reference's orientation (spots x cells)
    return cost[:, :st.shape[1]].T.cpu().numpy(), cs_sc.cpu().numpy(), cs_st.cpu().numpy()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("precision", ["f16x3", "f16"])
def test_against_reference_golden(cost_golden, tag, precision):
    eng = AssignmentEngine(precision=precision)
    g = cost_golden
    got, cs_sc, cs_st = build(eng, g[f"{tag}_sc_norm"], g[f"{tag}_st_norm"])
    want = np.rint(-g[f"{tag}_corr"] * 1e6)                # reference's matrix_correlation_pearson
    d = np.abs(got - want)
    assert d.max() <= TOL[precision], d.max()
    if precision == "f16x3":
        assert (d <= 1).mean() >= 0.90
    np.testing.assert_allclose(cs_sc[0], g[f"{tag}_sc_norm"].mean(0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(cs_sc[1], g[f"{tag}_sc_norm"].std(0), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cs_st[1], g[f"{tag}_st_norm"].std(0), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_fused_normalize_data(engine, cost_golden, tag):
    """log_tpm=1 fuses normalize_data (common.py:142-147) in front: raw counts in."""
    g = cost_golden
    got, _, _ = build(engine, g[f"{tag}_sc"], g[f"{tag}_st"], log_tpm=True)
    want = np.rint(-g[f"{tag}_corr"] * 1e6)
    assert np.abs(got - want).max() <= TOL["f16x3"]


@pytest.mark.parametrize("n_cells,n_spots,n_genes", [(1000, 1000, 2000), (300, 130, 777), (257, 129, 64),
                                                      (1, 1, 70), (513, 2, 1000)])
def test_shapes_vs_oracle(engine, n_cells, n_spots, n_genes):
    sc, st, _ = syn.structured_counts(n_cells, n_spots, n_genes, 1, seed=n_cells + n_genes)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    keep_sc = sc_n.std(0) > 0; keep_st = st_n.std(0) > 0
    sc_n, st_n = sc_n[:, keep_sc], st_n[:, keep_st]
    got, _, _ = build(engine, sc_n, st_n)
    want = co.cost_matrix_i32(sc_n, st_n)
    d = np.abs(got.astype(np.int64) - want)
    assert got.shape == want.shape and d.max() <= TOL["f16x3"], d.max()


def test_float32_input(engine):
    sc, st, _ = syn.structured_counts(200, 200, 500, 1, seed=3)
    sc_n, st_n = co.normalize_data(sc).astype(np.float32), co.normalize_data(st).astype(np.float32)
    got, _, _ = build(engine, sc_n, st_n, dtype=torch.float32)
    want = co.cost_matrix_i32(sc_n.astype(np.float64), st_n.astype(np.float64))
    assert np.abs(got - want).max() <= TOL["f16x3"]


def test_unstructured_stress(engine):
    sc, st, _ = syn.unstructured_counts(400, 400, 1500)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    got, _, _ = build(engine, sc_n, st_n)
    assert np.abs(got - co.cost_matrix_i32(sc_n, st_n)).max() <= TOL["f16x3"]


def test_deterministic(engine):
    sc, st, _ = syn.structured_counts(300, 300, 900, 1, seed=9)
    a, _, _ = build(engine, sc, st, log_tpm=True)
    b, _, _ = build(engine, sc, st, log_tpm=True)
    assert np.array_equal(a, b)


def test_errors(engine):
    dev = engine.device
    with pytest.raises(ValueError, match="same genes"):
        engine.cost_build(torch.zeros((5, 3), dtype=torch.float64, device=dev),
                          torch.zeros((6, 3), dtype=torch.float64, device=dev))
    x = torch.rand((50, 4), dtype=torch.float64, device=dev)
    y = x.clone(); y[:, 2] = 3.0                                   # zero-variance spot
    with pytest.raises(ValueError, match="zero variance"):
        engine.cost_build(x, y)


@pytest.mark.parametrize("n_genes", [20000, 30000])
def test_tolerance_at_the_baseline_gene_counts(engine, n_genes):
    """BASELINE configs 2-5 use 20 000 / 30 000 genes: the chunked tcgen05 accumulation (K = 3 * G fp16 products
    per entry) against the float64 oracle on a block small enough for numpy.  Same tolerance as above: every
    entry within 4 units of 1e-6, 90 % within 1 unit."""
    sc, st, _ = syn.structured_counts(320, 256, n_genes, 1, seed=77)
    sc_n, st_n = co.normalize_data(sc), co.normalize_data(st)
    got, _, _ = build(engine, sc_n, st_n)
    want = co.cost_matrix_i32(sc_n, st_n)
    d = np.abs(got.astype(np.int64) - want)
    assert d.max() <= TOL["f16x3"], d.max()
    assert (d <= 1).mean() >= 0.90
    # and with the fused normalize_data on raw counts
    got2, _, _ = build(engine, sc, st, log_tpm=True)
    d2 = np.abs(got2.astype(np.int64) - want)
    assert d2.max() <= TOL["f16x3"] and (d2 <= 1).mean() >= 0.90
