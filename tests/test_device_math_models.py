"""CPU: numpy models of two pieces of device arithmetic whose correctness is an identity or a
combinatorial argument (the GPU tests then check the kernels against the oracle):
  * the all-ascending bitonic network with virtual padding (metrics.cu: bitonic_sort_smem),
  * the Euclidean epilogue (cost_build.cu: EPI 1)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cost_oracle as co


def bitonic_flip_sort(keys):
    """The network of bitonic_sort_smem: every compare-exchange moves the smaller key to the LOWER
    index, pairs whose upper index is >= n are skipped (virtual +inf padding)."""
    k = list(keys)
    n = len(k)
    l2 = 0
    while (1 << l2) < n:
        l2 += 1
    half = (1 << l2) >> 1
    for ls in range(1, l2 + 1):
        size, lh = 1 << ls, ls - 1
        hs = 1 << lh
        for i in range(half):
            base, j = (i >> lh) << ls, i & (hs - 1)
            lo, hi = base + j, base + (size - 1 - j)
            if hi < n and k[hi] < k[lo]:
                k[lo], k[hi] = k[hi], k[lo]
        for ld in range(lh - 1, -1, -1):
            d = 1 << ld
            for i in range(half):
                lo = ((i >> ld) << (ld + 1)) + (i & (d - 1))
                hi = lo + d
                if hi < n and k[hi] < k[lo]:
                    k[lo], k[hi] = k[hi], k[lo]
    return k


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 8, 13, 31, 32, 33, 100, 257, 1000])
def test_bitonic_flip_network_sorts_any_length(n):
    rng = np.random.default_rng(n)
    for keys in (rng.integers(0, 5, n), rng.integers(0, 2 ** 62, n), np.arange(n)[::-1]):
        assert bitonic_flip_sort(keys.tolist()) == sorted(keys.tolist())


def test_euclidean_identity_from_standardised_operands():
    """|a - b|^2 = G[(mu_a - mu_b)^2 + (sd_a - sd_b)^2 + 2 sd_a sd_b (1 - r)], including constant columns."""
    rng = np.random.default_rng(1)
    a = rng.gamma(0.3, 2.0, (300, 40)); b = rng.gamma(0.5, 1.0, (300, 25))
    a[:, 3] = 1.5; b[:, 7] = 0.0                                       # sigma = 0: z = 0, r := 0
    G = a.shape[0]
    mu_a, sd_a, mu_b, sd_b = a.mean(0), a.std(0), b.mean(0), b.std(0)
    za = np.divide(a - mu_a, sd_a, out=np.zeros_like(a), where=sd_a > 0)
    zb = np.divide(b - mu_b, sd_b, out=np.zeros_like(b), where=sd_b > 0)
    r = zb.T @ za / G                                                   # spots x cells, what the GEMM accumulates / G
    d2 = G * ((mu_b[:, None] - mu_a[None, :]) ** 2 + (sd_b[:, None] - sd_a[None, :]) ** 2
              + 2 * sd_b[:, None] * sd_a[None, :] * (1 - r))
    want = co.euclidean_distance(a, b)
    np.testing.assert_allclose(np.sqrt(np.maximum(d2, 0)), want, rtol=1e-10, atol=1e-9)
