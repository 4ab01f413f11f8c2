// Pearson-correlation cost build on sm_100a.
//
// Replaces matrix_correlation_pearson (cytospace/common/common.py:190-199) as
// called by calculate_cost (cytospace/linear_assignment_solvers/
// linear_assignment_solvers.py:53-55):
//     corr = (v2.T.dot(v1) - outer(sum2, sum1)/G) / outer(std2, std1) / G
// which equals (1/G) * sum_g z2[g,s] * z1[g,c] with z = (x - mean) / sigma
// (population sigma).  So the build is
//   1. standardise: per-column mean / sigma (float64 sums, fixed reduction
//      order), then z as fp16 written K-major [columns x genes] -- an HBM-bound
//      read-twice / write-once pre-pass (optionally with normalize_data,
//      common.py:142-147, fused in front);
//   2. one TMA-fed tcgen05 GEMM  acc[s, c] = sum_k zst[s,k] * zsc[c,k]  with the
//      fp32 accumulator in TMEM and a fused epilogue
//      cost[s, c] = rint(-(1e6/G) * acc)  stored as int32 -- the integer matrix
//      the LAP solves (scale precedent: cytospace/cytospace.py:337).
// The `cost[location_repeat, :]` row expansion of linear_assignment_solvers.py:63-66
// is never materialised (lap_sap.cu turns it into object capacities).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdint>

#include "common.h"

namespace {

// ------------------------------------------------------------------ standardise

constexpr int kStatSplit = 32;      // gene-range splits of the column sums
constexpr int kStatCols = 32;       // columns per stats CTA
constexpr int kStatRows = 8;        // row lanes per stats CTA

template <typename T>
__device__ __forceinline__ double load_clean(const T *p) {
    const double v = (double)__ldg(p);
    if (v != v) return 0.0;                                          // np.nan_to_num on the input (common.py:143):
    return isinf(v) ? copysign(1.7976931348623157e308, v) : v;       // nan -> 0, +-inf -> +-DBL_MAX
}

// normalize_data (common.py:142-147): x * (1e6 / colsum), log2(. + 1), nan -> 0.
__device__ __forceinline__ double log_tpm(double x, double fac) {
    return fac > 0.0 ? log2(x * fac + 1.0) : 0.0;
}

// MODE 0: partial sum of the raw values.  MODE 1: partial sum and sum of squares
// of y (y = raw, or log-TPM when fac != nullptr).
template <typename T, int MODE>
__global__ void __launch_bounds__(kStatCols *kStatRows)
colsum_kernel(const T *__restrict__ x, long long ld, int n_genes, int n_cols,
              const double *__restrict__ fac, double *__restrict__ partial) {
    __shared__ double s1[kStatRows][kStatCols + 1], s2[kStatRows][kStatCols + 1];
    const int tx = threadIdx.x % kStatCols, ty = threadIdx.x / kStatCols;
    const int c = blockIdx.x * kStatCols + tx;
    const int per = (n_genes + gridDim.y - 1) / gridDim.y;
    const int g0 = blockIdx.y * per, g1 = min(n_genes, g0 + per);
    double a1 = 0.0, a2 = 0.0;
    if (c < n_cols) {
        const double f = (MODE == 1 && fac) ? fac[c] : 0.0;
        const bool tr = (MODE == 1 && fac);
#pragma unroll 4
        for (int g = g0 + ty; g < g1; g += kStatRows) {
            double v = load_clean(x + (long long)g * ld + c);
            if (tr) v = log_tpm(v, f);
            a1 += v;
            if (MODE == 1) a2 += v * v;
        }
    }
    s1[ty][tx] = a1; s2[ty][tx] = a2;
    __syncthreads();
    if (ty == 0 && c < n_cols) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < kStatRows; ++k) { t1 += s1[k][tx]; t2 += s2[k][tx]; }
        partial[((long long)blockIdx.y * n_cols + c) * 2 + 0] = t1;
        partial[((long long)blockIdx.y * n_cols + c) * 2 + 1] = t2;
    }
}

// MODE 0: fac[c] = 1e6 / colsum (0 when the column sums to 0: y == 0 everywhere).
// MODE 1: mean / sigma / 1/sigma per column + zero-variance count.
template <int MODE>
__global__ void colstat_finalize_kernel(const double *__restrict__ partial, int n_cols, int n_genes,
                                        int splits, double *__restrict__ fac, double *__restrict__ mean_o,
                                        double *__restrict__ inv_o, double *__restrict__ colstat,
                                        int32_t *__restrict__ zero_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < splits; ++k) {
        t1 += partial[((long long)k * n_cols + c) * 2 + 0];
        t2 += partial[((long long)k * n_cols + c) * 2 + 1];
    }
    if (MODE == 0) {
        fac[c] = (t1 > 0.0) ? 1.0e6 / t1 : 0.0;
    } else {
        const double mean = t1 / n_genes, ex2 = t2 / n_genes;
        double var = ex2 - mean * mean;
        const bool degenerate = !(var > 1e-12 * ex2);          // also catches NaN
        if (degenerate) var = 0.0;
        const double sd = sqrt(var);
        mean_o[c] = mean;
        inv_o[c] = degenerate ? 0.0 : 1.0 / sd;
        if (colstat) { colstat[c] = mean; colstat[n_cols + c] = sd; }
        if (degenerate && zero_var) atomicAdd(zero_var, 1);
    }
}

constexpr int kTileC = 64;   // columns (cells / spots) per transpose tile
constexpr int kTileG = 64;   // genes per transpose tile

// z[c, g] = (y[g, c] - mean[c]) * inv[c] as fp16, K-major; genes >= n_genes are
// zero-filled up to kp.  X3: three K segments of width kp hold hi/hi/lo (A
// operand) or hi/lo/hi (B operand) so one GEMM over 3*kp sums hi*hi+hi*lo+lo*hi.
template <typename T>
__global__ void __launch_bounds__(256)
standardise_write_kernel(const T *__restrict__ x, long long ld, int n_genes, int n_cols,
                         const double *__restrict__ fac, const double *__restrict__ mean,
                         const double *__restrict__ inv, int x3, int operand_b, long long kp,
                         __half *__restrict__ z) {
    __shared__ float tile[kTileC][kTileG + 1];
    const int c0 = blockIdx.x * kTileC, g0 = blockIdx.y * kTileG;
    {
        const int tx = threadIdx.x % kTileC, ty = threadIdx.x / kTileC;   // 64 x 4
        const int c = c0 + tx;
        double m = 0.0, iv = 0.0, f = 0.0;
        if (c < n_cols) { m = mean[c]; iv = inv[c]; f = fac ? fac[c] : 0.0; }
#pragma unroll 4
        for (int gg = ty; gg < kTileG; gg += 4) {
            const int g = g0 + gg;
            float zf = 0.f;
            if (c < n_cols && g < n_genes) {
                double v = load_clean(x + (long long)g * ld + c);
                if (fac) v = log_tpm(v, f);
                zf = (float)((v - m) * iv);
            }
            tile[tx][gg] = zf;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long kop = x3 ? 3 * kp : kp;
    for (int cc = w; cc < kTileC; cc += 8) {
        const int c = c0 + cc;
        if (c >= n_cols) break;
        const float a = tile[cc][2 * lane], b = tile[cc][2 * lane + 1];
        const __half2 hi = __floats2half2_rn(a, b);
        __half *row = z + (long long)c * kop + g0 + 2 * lane;
        *reinterpret_cast<__half2 *>(row) = hi;
        if (x3) {
            const float2 hf = __half22float2(hi);
            const __half2 lo = __floats2half2_rn(a - hf.x, b - hf.y);
            *reinterpret_cast<__half2 *>(row + kp) = operand_b ? lo : hi;
            *reinterpret_cast<__half2 *>(row + 2 * kp) = operand_b ? hi : lo;
        }
    }
}

__global__ void quantise_f64_kernel(const double *__restrict__ in, long long n_rows, long long n_cols,
                                    long long ld_in, double scale, int32_t *__restrict__ out,
                                    long long ld_out, int32_t *__restrict__ bad) {
    int nbad = 0;
    for (long long i = blockIdx.y; i < n_rows; i += gridDim.y) {
        for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_cols;
             j += (long long)gridDim.x * blockDim.x) {
            const double v = in[i * ld_in + j] * scale;
            int q = 0;
            if (!(fabs(v) < 1073741824.0)) ++nbad;       // NaN / inf / out of range
            else q = __double2int_rn(v);
            out[i * ld_out + j] = q;
        }
    }
    if (nbad && bad) atomicAdd(bad, nbad);
}

// ------------------------------------------------------------------------- GEMM
// C[128 x 256] tile per CTA, K stepped by 64 fp16 (one 128-byte swizzle atom per
// row), 4-stage TMA -> smem ring, tcgen05.mma.cta_group::1.kind::f16 with M=128,
// N=256, K=16 issued by one thread.  Persistent: one CTA per SM walks tiles in a
// grouped raster so that the CTAs resident at one time share A / B panels in L2.
//
// Accumulation is CHUNKED: the tensor core aligns every product to the running
// accumulator and truncates, which on a dot product that grows to G*r costs a
// systematic ~5e-9 * K relative error (measured: 23 units of 1e-6 at G = 2000,
// ~200 at G = 20000).  So the K loop is cut into chunks of kChunkKB * 64; chunk c
// accumulates from zero into TMEM accumulator c & 1 while 16 epilogue warps drain
// chunk c-1 into an fp32 register running sum (round-to-nearest adds).  The
// truncation then scales with the chunk length instead of K, and the drain
// overlaps the MMAs of the next chunk.

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;              // 16 KB
constexpr int kBBytes = BN * BK * 2;              // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;    // 48 KB
constexpr int kEpiWarps = 16;                     // 4 lane quadrants x 4 column groups of 64
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-19 epilogue
constexpr int kChunkKB = 8;                       // k-blocks (of 64) per accumulation chunk
constexpr int kTmemCols = 512;
constexpr int kGroupM = 16;                       // raster: m-blocks per group
constexpr size_t kGemmSmem = 1024 /*align slack*/ + (size_t)kStages * kStageBytes + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows 128 bytes apart, 8-row groups
// 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = F32 (bit 4), A = B = F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TileCoord { int m, n; };
__device__ __forceinline__ TileCoord tile_coord(int t, int mblocks, int nblocks) {
    const int per_group = kGroupM * nblocks;
    const int group = t / per_group, r = t - group * per_group;
    const int first_m = group * kGroupM;
    const int gm = min(kGroupM, mblocks - first_m);
    TileCoord tc;
    tc.m = first_m + r % gm;
    tc.n = r / gm;
    return tc;
}

// EPI 0: cost = rint(neg_scale * acc)                                  (correlation metrics)
// EPI 1: Euclidean distance from the SAME standardised operands.  With a = mu_a + sd_a * z_a,
//        b = mu_b + sd_b * z_b and sum_g z = 0:  a.b = G (mu_a mu_b + sd_a sd_b r),  r = acc / G, so
//            |a - b|^2 = G [ (mu_a - mu_b)^2 + (sd_a - sd_b)^2 + 2 sd_a sd_b (1 - r) ]
//        -- three non-negative terms, no cancellation of the large norms, and the GEMM keeps its
//        mixed-sign products (an un-centred a.b accumulates only positive products, and the tensor
//        core's truncating fp32 accumulate then biases it: measured 1e-4 relative on the distance).
//        cost = rint(scale * sqrt(.)); the epilogue runs in float64; stat_* = [mean | sd] per column.
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
cost_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 int n_spots, int n_cells, int num_kb, float neg_scale, int32_t *__restrict__ cost,
                 long long ld_cost, const double *__restrict__ stat_a, const double *__restrict__ stat_b,
                 double n_genes_d, int32_t *__restrict__ bad) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kStages * kStageBytes);
    // bars[0..3] full, [4..7] empty, [8..9] tmem_full, [10..11] tmem_empty, then the TMEM base
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mblocks = (n_spots + BM - 1) / BM, nblocks = (n_cells + BN - 1) / BN;
    const int ntiles = mblocks * nblocks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(t, mblocks, nblocks);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), kStageBytes);
                    const uint32_t sa = smem_u32(smem + (size_t)stage * kStageBytes);
                    tma_load_2d(sa, &map_a, full_bar(stage), kb * BK, tc.m * BM);
                    tma_load_2d(sa + kABytes, &map_b, full_bar(stage), kb * BK, tc.n * BN);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                for (int kb0 = 0; kb0 < num_kb; kb0 += kChunkKB) {
                    const int kb1 = min(num_kb, kb0 + kChunkKB);
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1);       // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)stage * kStageBytes);
                        const uint64_t adesc = umma_desc_sw128(sa);
                        const uint64_t bdesc = umma_desc_sw128(sa + kABytes);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // +32 bytes per K=16 step inside the swizzle atom: +2 in the (addr >> 4) field
                            tc_mma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, kIdesc, (kb > kb0 || k) ? 1u : 0u);
                        }
                        tc_commit(empty_bar(stage));      // smem slot free once these MMAs retire
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull_bar(acc));            // chunk accumulator complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: drain chunk accumulators into registers, then int32 -> global =====
        const int quad = warp & 3;                         // TMEM lanes [32*quad, 32*quad+32)
        const int cg = (warp - 4) >> 2;                    // columns [64*cg, 64*cg+64) of the tile
        int acc = 0; uint32_t acc_phase = 0;
        const bool vec_ok = ((ld_cost & 3) == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const TileCoord tc = tile_coord(t, mblocks, nblocks);
            float sum[64];
#pragma unroll
            for (int q = 0; q < 64; ++q) sum[q] = 0.f;
            for (int kb0 = 0; kb0 < num_kb; kb0 += kChunkKB) {
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + cg * 64);
#pragma unroll
                for (int c = 0; c < 64; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) sum[c + q] += __uint_as_float(v[q]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            const int row = tc.m * BM + quad * 32 + lane;
            const int col0 = tc.n * BN + cg * 64;
            if (row < n_spots) {
                int32_t *out_row = cost + (long long)row * ld_cost + col0;
                const double mu_a = (EPI == 1) ? stat_a[row] : 0.0, sd_a = (EPI == 1) ? stat_a[n_spots + row] : 0.0;
                int nbad = 0;
                auto quant = [&](int q) -> int {
                    if (EPI == 0) return __float2int_rn(neg_scale * sum[q]);
                    const int col = min(col0 + q, n_cells - 1);
                    const double mu_b = __ldg(stat_b + col), sd_b = __ldg(stat_b + n_cells + col);
                    const double dm = mu_a - mu_b, ds = sd_a - sd_b;
                    const double one_minus_r = fmax(1.0 - (double)sum[q] / n_genes_d, 0.0);
                    const double d2 = n_genes_d * (dm * dm + ds * ds + 2.0 * sd_a * sd_b * one_minus_r);
                    double v = (double)neg_scale * sqrt(d2);
                    if (!(v < 1073741823.0)) { v = 1073741823.0; ++nbad; }
                    return __double2int_rn(v);
                };
                if (vec_ok && col0 + 64 <= n_cells) {
#pragma unroll
                    for (int q = 0; q < 64; q += 4) {
                        int4 o;
                        o.x = quant(q); o.y = quant(q + 1); o.z = quant(q + 2); o.w = quant(q + 3);
                        *reinterpret_cast<int4 *>(out_row + q) = o;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 64; ++q)
                        if (col0 + q < n_cells) out_row[q] = quant(q);
                }
                if (EPI == 1 && nbad && bad) atomicAdd(bad, nbad);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn *out) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CYB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn)
            return cyb::set_error(CYB_ERR_UNSUPPORTED, "driver lacks cuTensorMapEncodeTiled");
        cached = reinterpret_cast<EncodeTiledFn>(fn);
    }
    *out = cached;
    return CYB_OK;
}

// [rows x k] fp16 K-major -> 2-D map, box = box_rows x 64 elements, 128-byte swizzle.
int make_operand_map(EncodeTiledFn enc, CUtensorMap *map, const void *base, int64_t rows, int64_t k, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cyb::set_error(CYB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CYB_OK;
}

struct StdLayout { size_t partial, fac, mean, inv, total; };
StdLayout std_layout(int64_t n_cols) {
    StdLayout L; size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o = cyb::align_up(o + b, 256); return r; };
    L.partial = take((size_t)kStatSplit * n_cols * 16);
    L.fac = take((size_t)n_cols * 8);
    L.mean = take((size_t)n_cols * 8);
    L.inv = take((size_t)n_cols * 8);
    L.total = o;
    return L;
}

template <typename T>
int standardise_impl(const T *x, int64_t n_genes, int64_t n_cols, int64_t ld, int log_tpm_flag, int precision,
                     int operand_b, __half *z, double *colstat, int32_t *zero_var, char *ws, cudaStream_t stream) {
    const StdLayout L = std_layout(n_cols);
    double *partial = reinterpret_cast<double *>(ws + L.partial);
    double *fac = reinterpret_cast<double *>(ws + L.fac);
    double *mean = reinterpret_cast<double *>(ws + L.mean);
    double *inv = reinterpret_cast<double *>(ws + L.inv);
    const int splits = (int)std::min<int64_t>(kStatSplit, std::max<int64_t>(1, n_genes / 64));
    const dim3 sgrid((unsigned)((n_cols + kStatCols - 1) / kStatCols), (unsigned)splits);
    const int fin_blocks = (int)((n_cols + 255) / 256);
    if (log_tpm_flag) {
        colsum_kernel<T, 0><<<sgrid, kStatCols * kStatRows, 0, stream>>>(x, ld, (int)n_genes, (int)n_cols, nullptr, partial);
        colstat_finalize_kernel<0><<<fin_blocks, 256, 0, stream>>>(partial, (int)n_cols, (int)n_genes, splits, fac,
                                                                   nullptr, nullptr, nullptr, nullptr);
    }
    const double *facp = log_tpm_flag ? fac : nullptr;
    colsum_kernel<T, 1><<<sgrid, kStatCols * kStatRows, 0, stream>>>(x, ld, (int)n_genes, (int)n_cols, facp, partial);
    colstat_finalize_kernel<1><<<fin_blocks, 256, 0, stream>>>(partial, (int)n_cols, (int)n_genes, splits, nullptr, mean,
                                                               inv, colstat, zero_var);
    const int64_t kp = cyb::align_up((size_t)n_genes, 64);
    const dim3 wgrid((unsigned)((n_cols + kTileC - 1) / kTileC), (unsigned)(kp / kTileG));
    standardise_write_kernel<T><<<wgrid, 256, 0, stream>>>(x, ld, (int)n_genes, (int)n_cols, facp, mean, inv,
                                                           precision == CYB_PREC_F16X3, operand_b, kp, z);
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}

}  // namespace

extern "C" int64_t cyb_operand_k(int64_t n_genes, int precision) {
    if (n_genes <= 0) return 0;
    const int64_t kp = (int64_t)cyb::align_up((size_t)n_genes, 64);
    return precision == CYB_PREC_F16X3 ? 3 * kp : kp;
}

extern "C" size_t cyb_standardise_workspace_bytes(int64_t n_genes, int64_t n_cols) {
    (void)n_genes;
    if (n_cols <= 0) return 256;
    return std_layout(n_cols).total;
}

extern "C" int cyb_standardise(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols, int64_t ld_x,
                               int log_tpm_flag, int precision, int operand_b, void *z_dev, double *colstat_dev,
                               int32_t *zero_var_dev, void *workspace_dev, size_t workspace_bytes, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x_dev || !z_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_standardise: null pointer argument");
    if (n_genes <= 0 || n_cols <= 0 || ld_x < n_cols || n_genes >= (1ll << 31) || n_cols >= (1ll << 31))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_standardise: bad shape genes=%lld cols=%lld ld=%lld",
                              (long long)n_genes, (long long)n_cols, (long long)ld_x);
    if (precision != CYB_PREC_F16 && precision != CYB_PREC_F16X3)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_standardise: unknown precision %d", precision);
    if (workspace_bytes < std_layout(n_cols).total)
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_standardise: workspace %zu < required %zu", workspace_bytes,
                              std_layout(n_cols).total);
    if ((reinterpret_cast<uintptr_t>(workspace_dev) & 255) || (reinterpret_cast<uintptr_t>(z_dev) & 15))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_standardise: workspace must be 256-byte, z 16-byte aligned");
    char *ws = static_cast<char *>(workspace_dev);
    if (x_dtype == CYB_F64)
        return standardise_impl(static_cast<const double *>(x_dev), n_genes, n_cols, ld_x, log_tpm_flag, precision,
                                operand_b, static_cast<__half *>(z_dev), colstat_dev, zero_var_dev, ws, stream);
    if (x_dtype == CYB_F32)
        return standardise_impl(static_cast<const float *>(x_dev), n_genes, n_cols, ld_x, log_tpm_flag, precision,
                                operand_b, static_cast<__half *>(z_dev), colstat_dev, zero_var_dev, ws, stream);
    return cyb::set_error(CYB_ERR_INVALID, "cyb_standardise: unknown dtype %d", x_dtype);
}

// fac[c] = 1e6 / colsum(x[:, c]) -- the TPM factor of normalize_data (common.py:144), shared with metrics.cu
int cyb_internal_tpm_factors(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols, int64_t ld_x,
                             double *fac_dev, double *partial_dev, cudaStream_t stream) {
    const int splits = (int)std::min<int64_t>(kStatSplit, std::max<int64_t>(1, n_genes / 64));
    const dim3 sgrid((unsigned)((n_cols + kStatCols - 1) / kStatCols), (unsigned)splits);
    if (x_dtype == CYB_F64)
        colsum_kernel<double, 0><<<sgrid, kStatCols * kStatRows, 0, stream>>>(
            static_cast<const double *>(x_dev), ld_x, (int)n_genes, (int)n_cols, nullptr, partial_dev);
    else
        colsum_kernel<float, 0><<<sgrid, kStatCols * kStatRows, 0, stream>>>(
            static_cast<const float *>(x_dev), ld_x, (int)n_genes, (int)n_cols, nullptr, partial_dev);
    colstat_finalize_kernel<0><<<(int)((n_cols + 255) / 256), 256, 0, stream>>>(
        partial_dev, (int)n_cols, (int)n_genes, splits, fac_dev, nullptr, nullptr, nullptr, nullptr);
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}

namespace {
int gemm_launch(const void *zst_dev, const void *zsc_dev, int64_t n_spots, int64_t n_cells, int64_t k, float scale,
                int32_t *cost_dev, int64_t ld_cost, const double *stat_a, const double *stat_b, double n_genes,
                int32_t *bad_dev, void *stream_v) {
    // (local names: "spots" = rows of the output / operand A, "cells" = columns / operand B)
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!zst_dev || !zsc_dev || !cost_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_gemm_i32: null pointer argument");
    if (n_spots <= 0 || n_cells <= 0 || k <= 0 || (k % BK) != 0 || ld_cost < n_cells ||
        n_spots >= (1ll << 31) || n_cells >= (1ll << 31) || k >= (1ll << 31))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_gemm_i32: bad shape spots=%lld cells=%lld k=%lld ld=%lld",
                              (long long)n_spots, (long long)n_cells, (long long)k, (long long)ld_cost);
    if ((reinterpret_cast<uintptr_t>(zst_dev) & 15) || (reinterpret_cast<uintptr_t>(zsc_dev) & 15))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_gemm_i32: operands must be 16-byte aligned");
    int dev = 0, sms = 0, major = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return cyb::set_error(CYB_ERR_UNSUPPORTED, "cost GEMM needs sm_100 (got cc %d.x)", major);
    EncodeTiledFn enc = nullptr;
    if (int rc = get_encode_fn(&enc)) return rc;
    CUtensorMap map_a, map_b;
    if (int rc = make_operand_map(enc, &map_a, zst_dev, n_spots, k, BM)) return rc;
    if (int rc = make_operand_map(enc, &map_b, zsc_dev, n_cells, k, BN)) return rc;
    const int64_t mblocks = (n_spots + BM - 1) / BM, nblocks = (n_cells + BN - 1) / BN;
    const int grid = (int)std::min<int64_t>(sms, mblocks * nblocks);
    if (stat_a) {
        CYB_CUDA_CHECK(cudaFuncSetAttribute(cost_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        cost_gemm_kernel<1><<<grid, kGemmThreads, kGemmSmem, stream>>>(map_a, map_b, (int)n_spots, (int)n_cells,
                                                                       (int)(k / BK), scale, cost_dev, ld_cost,
                                                                       stat_a, stat_b, n_genes, bad_dev);
    } else {
        CYB_CUDA_CHECK(cudaFuncSetAttribute(cost_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        cost_gemm_kernel<0><<<grid, kGemmThreads, kGemmSmem, stream>>>(map_a, map_b, (int)n_spots, (int)n_cells,
                                                                       (int)(k / BK), -scale, cost_dev, ld_cost,
                                                                       nullptr, nullptr, 0.0, nullptr);
    }
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}
}  // namespace

extern "C" int cyb_cost_gemm_i32(const void *za_dev, const void *zb_dev, int64_t n_a, int64_t n_b, int64_t k,
                                 float scale, int32_t *cost_dev, int64_t ld_cost, void *stream) {
    return gemm_launch(za_dev, zb_dev, n_a, n_b, k, scale, cost_dev, ld_cost, nullptr, nullptr, 0.0, nullptr, stream);
}

extern "C" int cyb_cost_gemm_euclid_i32(const void *za_dev, const void *zb_dev, int64_t n_a, int64_t n_b, int64_t k,
                                        int64_t n_genes, float scale, const double *colstat_a_dev,
                                        const double *colstat_b_dev, int32_t *cost_dev, int64_t ld_cost,
                                        int32_t *bad_dev, void *stream) {
    if (!colstat_a_dev || !colstat_b_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_gemm_euclid_i32: null pointer argument");
    if (n_genes <= 0) return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_gemm_euclid_i32: n_genes must be positive");
    return gemm_launch(za_dev, zb_dev, n_a, n_b, k, scale, cost_dev, ld_cost, colstat_a_dev, colstat_b_dev,
                       (double)n_genes, bad_dev, stream);
}

extern "C" size_t cyb_rank_workspace_bytes(int64_t n_genes, int64_t n_cols);
extern "C" int cyb_rank_columns(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols, int64_t ld_x,
                                int log_tpm, float *rank_dev, int64_t ld_rank, void *workspace_dev,
                                size_t workspace_bytes, void *stream);

namespace {
struct BuildLayout { size_t za, zb, std_ws, norm_a, norm_b, rank_a, rank_b, rank_ws, total; };
BuildLayout build_layout(int64_t n_genes, int64_t n_a, int64_t n_b, int precision, int metric) {
    BuildLayout L; size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o = cyb::align_up(o + b, 1024); return r; };
    const size_t kop = (size_t)cyb_operand_k(n_genes, precision);
    L.za = take((size_t)n_a * kop * 2);
    L.zb = take((size_t)n_b * kop * 2);
    L.std_ws = take(std_layout(std::max(n_a, n_b)).total);
    L.norm_a = L.norm_b = L.rank_a = L.rank_b = L.rank_ws = 0;
    if (metric == CYB_METRIC_EUCLIDEAN) {
        L.norm_a = take((size_t)n_a * 16);      // [mean | sd] per column
        L.norm_b = take((size_t)n_b * 16);
    } else if (metric == CYB_METRIC_SPEARMAN) {
        L.rank_a = take((size_t)n_genes * cyb::align_up((size_t)n_a, 32) * 4);
        L.rank_b = take((size_t)n_genes * cyb::align_up((size_t)n_b, 32) * 4);
        L.rank_ws = take(cyb_rank_workspace_bytes(n_genes, std::max(n_a, n_b)));
    }
    L.total = o;
    return L;
}
}  // namespace

extern "C" size_t cyb_cost_build_workspace_bytes(int64_t n_genes, int64_t n_a, int64_t n_b, int precision) {
    if (n_genes <= 0 || n_a <= 0 || n_b <= 0) return 1024;
    return build_layout(n_genes, n_a, n_b, precision, CYB_METRIC_PEARSON).total;
}

extern "C" size_t cyb_cost_build_metric_workspace_bytes(int metric, int64_t n_genes, int64_t n_a, int64_t n_b,
                                                        int precision) {
    if (n_genes <= 0 || n_a <= 0 || n_b <= 0) return 1024;
    return build_layout(n_genes, n_a, n_b, precision, metric).total;
}

extern "C" int cyb_cost_build(int metric, const void *a_dev, const void *b_dev, int x_dtype, int64_t n_genes,
                              int64_t n_a, int64_t n_b, int64_t ld_a, int64_t ld_b, int log_tpm_flag, int precision,
                              double cost_scale, int32_t *cost_dev, int64_t ld_cost, int32_t *bad_dev,
                              void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (metric == CYB_METRIC_PEARSON)
        return cyb_cost_build_pearson(a_dev, b_dev, x_dtype, n_genes, n_a, n_b, ld_a, ld_b, log_tpm_flag, precision,
                                      cost_scale, cost_dev, ld_cost, nullptr, nullptr, bad_dev, workspace_dev,
                                      workspace_bytes, stream);
    if (metric != CYB_METRIC_SPEARMAN && metric != CYB_METRIC_EUCLIDEAN)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build: unknown metric %d", metric);
    if (!workspace_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build: null workspace");
    if (n_genes <= 0 || n_a <= 0 || n_b <= 0) return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build: empty problem");
    const BuildLayout L = build_layout(n_genes, n_a, n_b, precision, metric);
    if (workspace_bytes < L.total)
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_cost_build: workspace %zu < required %zu", workspace_bytes, L.total);
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 1023)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build: workspace must be 1024-byte aligned");
    char *ws = static_cast<char *>(workspace_dev);
    const size_t std_bytes = std_layout(std::max(n_a, n_b)).total;
    const int64_t kop = cyb_operand_k(n_genes, precision);
    if (metric == CYB_METRIC_EUCLIDEAN) {
        // cdist(.., 'euclidean') (linear_assignment_solvers.py:51,59) from the standardised operands of
        // the Pearson build plus the column means / sigmas (a constant column is fine here: z = 0)
        double *sa = reinterpret_cast<double *>(ws + L.norm_a), *sb = reinterpret_cast<double *>(ws + L.norm_b);
        if (int rc = cyb_standardise(a_dev, x_dtype, n_genes, n_a, ld_a, log_tpm_flag, precision, 0, ws + L.za, sa,
                                     nullptr, ws + L.std_ws, std_bytes, stream)) return rc;
        if (int rc = cyb_standardise(b_dev, x_dtype, n_genes, n_b, ld_b, log_tpm_flag, precision, 1, ws + L.zb, sb,
                                     nullptr, ws + L.std_ws, std_bytes, stream)) return rc;
        return cyb_cost_gemm_euclid_i32(ws + L.za, ws + L.zb, n_a, n_b, kop, n_genes, (float)cost_scale, sa, sb,
                                        cost_dev, ld_cost, bad_dev, stream);
    }
    // Spearman (common.py:202-215): Pearson of the per-column average ranks
    float *ra = reinterpret_cast<float *>(ws + L.rank_a), *rb = reinterpret_cast<float *>(ws + L.rank_b);
    const int64_t lda_r = (int64_t)cyb::align_up((size_t)n_a, 32), ldb_r = (int64_t)cyb::align_up((size_t)n_b, 32);
    const size_t rank_bytes = cyb_rank_workspace_bytes(n_genes, std::max(n_a, n_b));
    if (int rc = cyb_rank_columns(a_dev, x_dtype, n_genes, n_a, ld_a, log_tpm_flag, ra, lda_r, ws + L.rank_ws,
                                  rank_bytes, stream)) return rc;
    if (int rc = cyb_rank_columns(b_dev, x_dtype, n_genes, n_b, ld_b, log_tpm_flag, rb, ldb_r, ws + L.rank_ws,
                                  rank_bytes, stream)) return rc;
    if (int rc = cyb_standardise(ra, CYB_F32, n_genes, n_a, lda_r, 0, precision, 0, ws + L.za, nullptr, bad_dev,
                                 ws + L.std_ws, std_bytes, stream)) return rc;
    if (int rc = cyb_standardise(rb, CYB_F32, n_genes, n_b, ldb_r, 0, precision, 1, ws + L.zb, nullptr, bad_dev,
                                 ws + L.std_ws, std_bytes, stream)) return rc;
    return cyb_cost_gemm_i32(ws + L.za, ws + L.zb, n_a, n_b, kop, (float)(cost_scale / (double)n_genes), cost_dev,
                             ld_cost, stream);
}

extern "C" int cyb_cost_build_pearson(const void *a_dev, const void *b_dev, int x_dtype, int64_t n_genes,
                                      int64_t n_a, int64_t n_b, int64_t ld_a, int64_t ld_b, int log_tpm_flag,
                                      int precision, double cost_scale, int32_t *cost_dev, int64_t ld_cost,
                                      double *colstat_a_dev, double *colstat_b_dev, int32_t *zero_var_dev,
                                      void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!workspace_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build_pearson: null workspace");
    if (n_genes <= 0 || n_a <= 0 || n_b <= 0)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build_pearson: empty problem");
    const BuildLayout L = build_layout(n_genes, n_a, n_b, precision, CYB_METRIC_PEARSON);
    if (workspace_bytes < L.total)
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_cost_build_pearson: workspace %zu < required %zu",
                              workspace_bytes, L.total);
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 1023)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_cost_build_pearson: workspace must be 1024-byte aligned");
    char *ws = static_cast<char *>(workspace_dev);
    const size_t std_bytes = std_layout(std::max(n_a, n_b)).total;
    // matrix a supplies the rows of the output (GEMM operand A), matrix b the columns (operand B)
    if (int rc = cyb_standardise(a_dev, x_dtype, n_genes, n_a, ld_a, log_tpm_flag, precision, 0, ws + L.za,
                                 colstat_a_dev, zero_var_dev, ws + L.std_ws, std_bytes, stream)) return rc;
    if (int rc = cyb_standardise(b_dev, x_dtype, n_genes, n_b, ld_b, log_tpm_flag, precision, 1, ws + L.zb,
                                 colstat_b_dev, zero_var_dev, ws + L.std_ws, std_bytes, stream)) return rc;
    return cyb_cost_gemm_i32(ws + L.za, ws + L.zb, n_a, n_b, cyb_operand_k(n_genes, precision),
                             (float)(cost_scale / (double)n_genes), cost_dev, ld_cost, stream);
}

extern "C" int cyb_quantise_f64(const double *in_dev, int64_t n_rows, int64_t n_cols, int64_t ld_in, double scale,
                                int32_t *out_dev, int64_t ld_out, int32_t *bad_dev, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!in_dev || !out_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_quantise_f64: null pointer argument");
    if (n_rows <= 0 || n_cols <= 0 || ld_in < n_cols || ld_out < n_cols)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_quantise_f64: bad shape");
    const unsigned gx = (unsigned)std::min<int64_t>((n_cols + 255) / 256, 64);
    const unsigned gy = (unsigned)std::min<int64_t>(n_rows, 2048);
    quantise_f64_kernel<<<dim3(gx, gy), 256, 0, stream>>>(in_dev, n_rows, n_cols, ld_in, scale, out_dev, ld_out, bad_dev);
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}
