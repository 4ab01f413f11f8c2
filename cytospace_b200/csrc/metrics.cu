// The other two distance metrics and the integerised `lap_CSPR` matrix of CytoSPACE on sm_100a.
//
//  * cyb_rank_columns      <- `pd.DataFrame(v).rank().values` inside matrix_correlation_spearman
//                             (cytospace/common/common.py:202-215): per-column AVERAGE ranks.  The
//                             Spearman cost is then the Pearson pipeline (cost_build.cu) on the ranks.
//  * cyb_expand_rows_noise_i32
//                          <- the lap_CSPR matrix of cytospace/cytospace.py:334-340:
//                             int(1e6 * cost[location_repeat, :] + 10 * U(0,1) + 1); the expansion
//                             (linear_assignment_solvers.py:63-66) IS materialised here because the
//                             noise differs per (slot, cell).  U comes from a counter-based hash of
//                             (seed, slot, cell) instead of the reference's MT19937 stream
//                             (documented deviation: same distribution, different numbers).
//
// Ranking.  One CTA per column.  The tie group of the column MINIMUM (the zeros of an expression
// column, >= 90 % of it) is counted, not sorted.  The other values are mapped to order-preserving 64-bit keys
// (exact: no float32 shortcut, ties are ties of the doubles), sorted in shared memory by a bitonic
// network in its all-ascending "flip" form -- which sorts any length with VIRTUAL +inf padding:
// every compare-exchange moves the smaller key to the lower index, so slots >= n never have to
// exist -- and every element then finds  less = lower_bound, leq = upper_bound  in the sorted
// keys: average rank = (less + leq + 1) / 2  (pandas method="average", 1-based).  Columns longer
// than the shared-memory run (28 672 keys) are ranked run by run, the counts accumulated in a
// per-CTA scratch row.  HBM-bound in principle (read 8 B, write 4 B per element, both strided by
// the row pitch -- neighbouring CTAs share the 32-byte sectors through L2); the sort is on-chip.
#include <algorithm>
#include <cstdint>

#include "common.h"

namespace {

constexpr int kRankThreads = 1024;
constexpr int kRunCap = 28672;                 // keys per shared-memory run (224 KB)
constexpr int kScratchCtas = 296;              // CTAs of a multi-run launch (each owns a count row)

__device__ __forceinline__ unsigned long long order_key(double v) {
    v = (v != v) ? 0.0 : v + 0.0;             // nan -> 0 (np.nan_to_num upstream), -0.0 -> +0.0
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double transform(double x, double fac, bool tr) {
    x = (x != x) ? 0.0 : x;
    if (!tr) return x;
    return fac > 0.0 ? log2(x * fac + 1.0) : 0.0;     // normalize_data, common.py:142-147
}

__device__ __forceinline__ void cmpx(unsigned long long *k, int lo, int hi) {
    const unsigned long long a = k[lo], b = k[hi];
    if (b < a) { k[lo] = b; k[hi] = a; }
}

// Ascending sort of k[0..n) in shared memory (all threads of the CTA).
__device__ void bitonic_sort_smem(unsigned long long *k, int n) {
    int l2 = 0;
    while ((1 << l2) < n) ++l2;
    const int half = (1 << l2) >> 1;
    for (int ls = 1; ls <= l2; ++ls) {                 // blocks of size 2^ls
        const int size = 1 << ls, lh = ls - 1, hs = 1 << lh;
        // flip: element j of the lower half of each block against element (size-1-j) of the block
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            const int base = (i >> lh) << ls, j = i & (hs - 1);
            const int lo = base + j, hi = base + (size - 1 - j);
            if (hi < n) cmpx(k, lo, hi);
        }
        __syncthreads();
        for (int ld = lh - 1; ld >= 0; --ld) {         // distance 2^ld
            const int d = 1 << ld;
            for (int i = threadIdx.x; i < half; i += blockDim.x) {
                const int lo = ((i >> ld) << (ld + 1)) + (i & (d - 1)), hi = lo + d;
                if (hi < n) cmpx(k, lo, hi);
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int lower_bound(const unsigned long long *k, int n, unsigned long long x) {
    int lo = 0, hi = n;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (k[m] < x) lo = m + 1; else hi = m; }
    return lo;
}
__device__ __forceinline__ int upper_bound(const unsigned long long *k, int n, unsigned long long x) {
    int lo = 0, hi = n;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (k[m] <= x) lo = m + 1; else hi = m; }
    return lo;
}

template <typename T>
__global__ void __launch_bounds__(kRankThreads)
rank_columns_kernel(const T *__restrict__ x, long long ld, int n_genes, int n_cols, const double *__restrict__ fac,
                    float *__restrict__ r, long long ld_r, int run_cap, unsigned int *__restrict__ scratch) {
    extern __shared__ __align__(16) unsigned long long skeys[];
    __shared__ unsigned long long red_key[32];
    __shared__ int red_cnt[32], s_fill;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nruns = (n_genes + run_cap - 1) / run_cap;
    unsigned int *acc = scratch ? scratch + (size_t)blockIdx.x * n_genes : nullptr;
    for (int c = blockIdx.x; c < n_cols; c += gridDim.x) {
        const bool tr = fac != nullptr;
        const double f = tr ? fac[c] : 0.0;
        auto key_of = [&](int g) { return order_key(transform((double)__ldg(x + (long long)g * ld + c), f, tr)); };

        // ---- the column minimum and its multiplicity: that tie group (the zeros of an expression
        // column, >= 90 % of it) needs no sorting -- rank (n_min + 1) / 2, everything else above it
        unsigned long long kmin = ~0ull;
        int n_min = 0;
        for (int g = threadIdx.x; g < n_genes; g += blockDim.x) {
            const unsigned long long k = key_of(g);
            if (k < kmin) { kmin = k; n_min = 1; } else if (k == kmin) ++n_min;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, kmin, d);
            const int oc = __shfl_xor_sync(0xffffffffu, n_min, d);
            if (ok < kmin) { kmin = ok; n_min = oc; } else if (ok == kmin) n_min += oc;
        }
        if (lane == 0) { red_key[warp] = kmin; red_cnt[warp] = n_min; }
        if (threadIdx.x == 0) s_fill = 0;
        __syncthreads();
        kmin = red_key[0]; n_min = red_cnt[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            const unsigned long long ok = red_key[w];
            if (ok < kmin) { kmin = ok; n_min = red_cnt[w]; } else if (ok == kmin) n_min += red_cnt[w];
        }
        const int n_rest = n_genes - n_min;

        if (n_rest <= run_cap) {
            // ---- one run holds everything above the minimum: compact (any order), sort, look up
            for (int g0 = 0; g0 < n_genes; g0 += blockDim.x) {
                const int g = g0 + threadIdx.x;
                const unsigned long long k = g < n_genes ? key_of(g) : kmin;
                const bool keep = k != kmin;
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&s_fill, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) skeys[base + __popc(m & ((1u << lane) - 1u))] = k;
            }
            __syncthreads();
            bitonic_sort_smem(skeys, n_rest);
            for (int g = threadIdx.x; g < n_genes; g += blockDim.x) {
                const unsigned long long k = key_of(g);
                unsigned int s2 = (unsigned)n_min;                              // 2 * rank - 1
                if (k != kmin)
                    s2 = 2u * (unsigned)n_min + (unsigned)lower_bound(skeys, n_rest, k) + (unsigned)upper_bound(skeys, n_rest, k);
                r[(long long)g * ld_r + c] = 0.5f * (float)(s2 + 1u);
            }
            __syncthreads();
            continue;
        }
        // ---- long dense column: rank run by run over index ranges, counts accumulated per element
        for (int run = 0; run < nruns; ++run) {
            const int g0 = run * run_cap, n = min(run_cap, n_genes - g0);
            for (int g = threadIdx.x; g < n; g += blockDim.x) skeys[g] = key_of(g0 + g);
            __syncthreads();
            bitonic_sort_smem(skeys, n);
            for (int g = threadIdx.x; g < n_genes; g += blockDim.x) {
                const unsigned long long key = key_of(g);
                unsigned int s2 = (unsigned)lower_bound(skeys, n, key) + (unsigned)upper_bound(skeys, n, key);
                if (run > 0) s2 += acc[g];
                if (run + 1 < nruns) { acc[g] = s2; continue; }
                r[(long long)g * ld_r + c] = 0.5f * (float)(s2 + 1u);
            }
            __syncthreads();
        }
    }
}

// splitmix64 finaliser: the counter-based generator behind the lap_CSPR tie noise
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void expand_rows_noise_kernel(const int32_t *__restrict__ cost, long long ld, long long n_rows_out,
                                         long long n_cols, const int32_t *__restrict__ row_map,
                                         unsigned long long seed, int noise_lo, int noise_span,
                                         int32_t *__restrict__ out, long long ld_out) {
    for (long long i = blockIdx.y; i < n_rows_out; i += gridDim.y) {
        const long long src = row_map ? row_map[i] : i;
        const unsigned long long rowkey = mix64(seed ^ ((unsigned long long)i * 0xD1B54A32D192ED03ull));
        for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_cols;
             j += (long long)gridDim.x * blockDim.x) {
            int v = cost[src * ld + j];
            if (noise_span > 0) {
                const unsigned long long h = mix64(rowkey + (unsigned long long)j);
                v += noise_lo + (int)(((h >> 32) * (unsigned long long)noise_span) >> 32);   // floor(span * U)
            }
            out[i * ld_out + j] = v;
        }
    }
}

}  // namespace

extern "C" size_t cyb_rank_workspace_bytes(int64_t n_genes, int64_t n_cols) {
    if (n_genes <= 0 || n_cols <= 0) return 256;
    size_t b = cyb::align_up((size_t)n_cols * 8, 256) * 2;                    // log-TPM factors + partial sums
    b += cyb::align_up((size_t)32 * n_cols * 16, 256);
    if (n_genes > kRunCap) b += cyb::align_up((size_t)4 * kScratchCtas * (size_t)n_genes, 256);   // count scratch per CTA
    return b;
}

// implemented in cost_build.cu: fac[c] = 1e6 / colsum(x[:, c]) (0 for an all-zero column)
int cyb_internal_tpm_factors(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols, int64_t ld_x,
                             double *fac_dev, double *partial_dev, cudaStream_t stream);

extern "C" int cyb_rank_columns(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols, int64_t ld_x,
                                int log_tpm, float *rank_dev, int64_t ld_rank, void *workspace_dev,
                                size_t workspace_bytes, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x_dev || !rank_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_rank_columns: null pointer argument");
    if (n_genes <= 0 || n_cols <= 0 || ld_x < n_cols || ld_rank < n_cols || n_genes >= (1ll << 24) || n_cols >= (1ll << 31))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_rank_columns: bad shape genes=%lld cols=%lld (genes < 2^24)",
                              (long long)n_genes, (long long)n_cols);
    if (x_dtype != CYB_F64 && x_dtype != CYB_F32)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_rank_columns: unknown dtype %d", x_dtype);
    if (workspace_bytes < cyb_rank_workspace_bytes(n_genes, n_cols))
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_rank_columns: workspace %zu < required %zu", workspace_bytes,
                              cyb_rank_workspace_bytes(n_genes, n_cols));
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 255)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_rank_columns: workspace must be 256-byte aligned");
    char *ws = static_cast<char *>(workspace_dev);
    double *fac = reinterpret_cast<double *>(ws);
    size_t off = cyb::align_up((size_t)n_cols * 8, 256) * 2;
    double *partial = reinterpret_cast<double *>(ws + off);
    off += cyb::align_up((size_t)32 * n_cols * 16, 256);
    unsigned int *scratch = n_genes > kRunCap ? reinterpret_cast<unsigned int *>(ws + off) : nullptr;
    if (log_tpm)
        if (int rc = cyb_internal_tpm_factors(x_dev, x_dtype, n_genes, n_cols, ld_x, fac, partial, stream)) return rc;

    int dev = 0, sms = 0, max_smem = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int run_cap = kRunCap;
    if ((size_t)run_cap * 8 > (size_t)max_smem) run_cap = max_smem / 8 / 1024 * 1024;
    const int nruns = (int)((n_genes + run_cap - 1) / run_cap);
    const int run_len = (int)((n_genes + nruns - 1) / nruns);                // balanced runs
    const size_t smem = (size_t)std::min<int64_t>(n_genes, run_len) * 8;
    int grid = (int)n_cols;
    if (scratch) grid = (int)std::min<int64_t>(n_cols, kScratchCtas);
    const double *facp = log_tpm ? fac : nullptr;
    if (x_dtype == CYB_F64) {
        CYB_CUDA_CHECK(cudaFuncSetAttribute(rank_columns_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rank_columns_kernel<double><<<grid, kRankThreads, smem, stream>>>(
            static_cast<const double *>(x_dev), ld_x, (int)n_genes, (int)n_cols, facp, rank_dev, ld_rank, run_len, scratch);
    } else {
        CYB_CUDA_CHECK(cudaFuncSetAttribute(rank_columns_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rank_columns_kernel<float><<<grid, kRankThreads, smem, stream>>>(
            static_cast<const float *>(x_dev), ld_x, (int)n_genes, (int)n_cols, facp, rank_dev, ld_rank, run_len, scratch);
    }
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}

extern "C" int cyb_expand_rows_noise_i32(const int32_t *cost_dev, int64_t ld, int64_t n_rows_out, int64_t n_cols,
                                         const int32_t *row_map_dev, uint64_t seed, int noise_lo, int noise_span,
                                         int32_t *out_dev, int64_t ld_out, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!cost_dev || !out_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_expand_rows_noise_i32: null pointer argument");
    if (n_rows_out <= 0 || n_cols <= 0 || ld < n_cols || ld_out < n_cols || noise_span < 0)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_expand_rows_noise_i32: bad shape");
    const unsigned gx = (unsigned)std::min<int64_t>((n_cols + 255) / 256, 64);
    const unsigned gy = (unsigned)std::min<int64_t>(n_rows_out, 4096);
    expand_rows_noise_kernel<<<dim3(gx, gy), 256, 0, stream>>>(cost_dev, ld, n_rows_out, n_cols, row_map_dev, seed,
                                                               noise_lo, noise_span, out_dev, ld_out);
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}
