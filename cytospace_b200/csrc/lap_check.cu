// On-device optimality certificate of the LAP solve, and the HBM roofline probe of its row scan.
//
// No reference counterpart (the reference trusts lapjv, linear_assignment_solvers.py:38): given the
// assignment and object prices cyb_lap_solve_i32 returned, one coalesced pass over the cost matrix checks
// eps-complementary slackness with eps = 1 in units of 1/(P+1) cost -- which proves the assignment optimal
// for the integer matrix -- the capacities and the total.

#include <algorithm>
#include <climits>
#include <cstdint>

#include "common.h"

namespace {

constexpr int kPersonBits = 18;

// Row data is streamed: no L1 allocation, L2 lines marked evict-first.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int4 ld_stream(const int4 *p, unsigned long long pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}

// ---------------------------------------------------------------------------------
// Certificate / row-scan pass.  Grid = (object tiles, person groups); a CTA stages a
// tile of object prices in shared memory once and streams `rows_per_cta` rows
// against it, one warp per row, so the cost matrix is read exactly once from HBM and
// the price vector once per person group from L2.
constexpr int kChkThreads = 512;
constexpr int kChkTileCols = 4096;

__global__ void __launch_bounds__(kChkThreads) lap_rowmin_kernel(
    const int32_t *__restrict__ cost, long long ld, int np, int no,
    const long long *__restrict__ price, long long S, int rows_per_cta, long long *__restrict__ rowmin) {
    __shared__ __align__(16) long long sp[kChkTileCols];
    const int c0 = blockIdx.x * kChkTileCols;
    const int nc = min(kChkTileCols, no - c0);
    for (int j = threadIdx.x; j < nc; j += kChkThreads) sp[j] = price[c0 + j];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long pol = l2_policy_evict_first();
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(np, r0 + rows_per_cta);
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
    for (int i = r0 + w; i < r1; i += kChkThreads / 32) {
        const int32_t *r = cost + (long long)i * ld + c0;
        long long m = LLONG_MAX;
        int jt = 0;
        if (vec_ok) {
            const int4 *r4 = reinterpret_cast<const int4 *>(r);
            const int n4 = nc >> 2;
#pragma unroll 4
            for (int q = lane; q < n4; q += 32) {
                const int4 c = ld_stream(r4 + q, pol);
                const longlong2 a = *reinterpret_cast<const longlong2 *>(sp + 4 * q);
                const longlong2 bb = *reinterpret_cast<const longlong2 *>(sp + 4 * q + 2);
                m = min(m, (long long)c.x * S + a.x);
                m = min(m, (long long)c.y * S + a.y);
                m = min(m, (long long)c.z * S + bb.x);
                m = min(m, (long long)c.w * S + bb.y);
            }
            jt = n4 << 2;
        }
        for (int j = jt + lane; j < nc; j += 32) m = min(m, (long long)__ldg(r + j) * S + sp[j]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) atomicMin(rowmin + i, m);
    }
}

// Whole-row variant for price vectors that fit in shared memory (objects <= kChkWholeMax): every CTA
// stages ALL prices once, a TEAM of four warps streams one row at a time (rows are dealt to ~8 teams
// per SM, so a 10k-row matrix still balances to within one row in nine), and the certificate terms of
// that row (violation, cost, capacity count) are taken in the same pass -- no row-minimum buffer, no
// atomics on it, no second kernel.  acc = {max violation, total, invalid rows} (zero-initialised).
constexpr int kChkWholeMax = 12288;              // 96 KB of prices -> two CTAs per SM
constexpr int kTeam = 128;                       // threads per row team (measured: 256-thread teams are slower, 148 vs 102 us at 10k)
constexpr int kTeams = kChkThreads / kTeam;
constexpr int kBatch = 4;                        // 16-byte row loads in flight per thread

__global__ void __launch_bounds__(kChkThreads, 2) lap_rowcheck_whole_kernel(
    const int32_t *__restrict__ cost, long long ld, int np, int no, const int32_t *__restrict__ person_obj,
    const long long *__restrict__ price, long long S, int32_t *__restrict__ count, long long *__restrict__ acc,
    const int32_t *__restrict__ soff, long long *__restrict__ out, unsigned int *__restrict__ done,
    unsigned int *__restrict__ rowctr) {
    extern __shared__ __align__(16) long long spw[];
    __shared__ long long part[2][kTeams][kTeam / 32];
    __shared__ int s_next[2][kTeams];
    __shared__ long long r_viol[kTeams], r_tot[kTeams];
    __shared__ int r_bad[kTeams];
    for (int j = threadIdx.x; j < no; j += kChkThreads) spw[j] = price[j];
    __syncthreads();
    const int team = threadIdx.x / kTeam, tt = threadIdx.x % kTeam, lane = tt & 31, wt = tt >> 5;
    const unsigned long long pol = l2_policy_evict_first();
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
    const int n4 = vec_ok ? (no >> 2) : 0;
    long long viol = 0, tot = 0;
    int bad = 0, it = 0;
    // Rows are handed out dynamically after the first one per team (10k rows over 2 368 teams is 4.2 rows each: a static
    // deal leaves most teams idle during the fifth); the leader draws its ticket one row ahead.
    const int i0 = blockIdx.x * kTeams + team, n_teams = gridDim.x * kTeams;
    // the row's certificate terms ride along with the stream: obj(i) is fetched one row ahead and
    // cost[i, obj(i)] with the row itself, so the team leader adds no dependent round trip per row
    int o_cur = (tt == 0 && i0 < np) ? __ldg(person_obj + i0) : -1;
    int i_next = np;
    if (tt == 0 && i0 < np) i_next = (int)min((unsigned)np, atomicAdd(rowctr, 1u) + (unsigned)n_teams);
    for (int i = i0; i < np; ++it) {
        const int32_t *r = cost + (long long)i * ld;
        const int o_next = (tt == 0 && i_next < np) ? __ldg(person_obj + i_next) : -1;
        int i_next2 = np;
        if (tt == 0 && i_next < np) i_next2 = (int)min((unsigned)np, atomicAdd(rowctr, 1u) + (unsigned)n_teams);
        const bool o_ok = o_cur >= 0 && o_cur < no;
        const int c_o = (tt == 0 && o_ok) ? __ldg(r + o_cur) : 0;
        long long m = LLONG_MAX;
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
        // four 16-byte loads per thread are requested before the first one is consumed (ncu, round 1: the loop was
        // stalled on its own loads -- 9.5 long-scoreboard stalls per issue, 54.8 % DRAM throughput)
        for (int q0 = tt; q0 < n4; q0 += kBatch * kTeam) {
            int4 c[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int q = q0 + u * kTeam;
                c[u] = q < n4 ? ld_stream(r4 + q, pol) : make_int4(INT_MAX, INT_MAX, INT_MAX, INT_MAX);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int q = min(q0 + u * kTeam, n4 - 1);            // (a padded lane re-reads a valid price: its cost is INT_MAX)
                const longlong2 a = *reinterpret_cast<const longlong2 *>(spw + 4 * q);
                const longlong2 bb = *reinterpret_cast<const longlong2 *>(spw + 4 * q + 2);
                m = min(m, (long long)c[u].x * S + a.x);
                m = min(m, (long long)c[u].y * S + a.y);
                m = min(m, (long long)c[u].z * S + bb.x);
                m = min(m, (long long)c[u].w * S + bb.y);
            }
        }
        for (int j = (n4 << 2) + tt; j < no; j += kTeam) m = min(m, (long long)__ldg(r + j) * S + spw[j]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) part[it & 1][team][wt] = m;
        if (tt == 0) s_next[it & 1][team] = i_next;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(kTeam) : "memory");     // the team only
        if (tt == 0) {
#pragma unroll
            for (int k = 0; k < kTeam / 32; ++k) m = min(m, part[it & 1][team][k]);
            if (!o_ok) ++bad;
            else {
                tot += c_o;
                atomicAdd(count + o_cur, 1);
                viol = max(viol, (long long)c_o * S + spw[o_cur] - m);
            }
        }
        i = s_next[it & 1][team];
        o_cur = o_next;
        i_next = i_next2;
    }
    if (tt == 0) { r_viol[team] = viol; r_tot[team] = tot; r_bad[team] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kTeams; ++k) { viol = max(viol, r_viol[k]); tot += r_tot[k]; bad += r_bad[k]; }
        if (viol > 0) atomicMax(acc + 0, viol);
        if (tot) atomicAdd(reinterpret_cast<unsigned long long *>(acc + 1), (unsigned long long)tot);
        if (bad) atomicAdd(reinterpret_cast<unsigned long long *>(acc + 2), (unsigned long long)bad);
        __threadfence();
        r_bad[0] = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1 : 0;         // the last CTA finishes the certificate
    }
    __syncthreads();
    if (r_bad[0]) {
        __threadfence();
        long long capbad = 0;
        for (int o = threadIdx.x; o < no; o += kChkThreads) {
            const int cap = soff ? __ldg(soff + o + 1) - __ldg(soff + o) : 1;
            if (__ldcg(count + o) != cap) ++capbad;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) capbad += __shfl_xor_sync(0xffffffffu, capbad, d);
        if (threadIdx.x == 0) r_tot[0] = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && capbad) atomicAdd(reinterpret_cast<unsigned long long *>(&r_tot[0]), (unsigned long long)capbad);
        __syncthreads();
        if (threadIdx.x == 0) {
            out[0] = __ldcg(acc + 0); out[1] = __ldcg(acc + 1); out[2] = __ldcg(acc + 2); out[3] = r_tot[0];
        }
    }
}

__global__ void lap_check_finish_kernel(const int32_t *__restrict__ cost, long long ld, int np, int no,
                                        const int32_t *__restrict__ person_obj,
                                        const long long *__restrict__ price, long long S,
                                        const long long *__restrict__ rowmin, int32_t *__restrict__ count,
                                        long long *out) {
    long long viol = LLONG_MIN, tot = 0, bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const int o = person_obj[i];
        if (o < 0 || o >= no) { ++bad; continue; }
        const int c = cost[(long long)i * ld + o];
        tot += c;
        atomicAdd(count + o, 1);
        viol = max(viol, (long long)c * S + price[o] - rowmin[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        viol = max(viol, __shfl_xor_sync(0xffffffffu, viol, d));
        tot += __shfl_xor_sync(0xffffffffu, tot, d);
        bad += __shfl_xor_sync(0xffffffffu, bad, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out + 0, viol);
        atomicAdd(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)tot);
        atomicAdd(reinterpret_cast<unsigned long long *>(out + 2), (unsigned long long)bad);
    }
}

__global__ void lap_check_capacity_kernel(const int32_t *__restrict__ soff, int no,
                                          const int32_t *__restrict__ count, long long *out) {
    long long bad = 0;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < no; o += gridDim.x * blockDim.x) {
        const int cap = soff ? soff[o + 1] - soff[o] : 1;
        if (count[o] != cap) ++bad;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, d);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(reinterpret_cast<unsigned long long *>(out + 3), (unsigned long long)bad);
}

struct ChkLayout {
    size_t rowmin, count, total;
};

ChkLayout chk_layout(int64_t np, int64_t no) {
    ChkLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = cyb::align_up(o + bytes, 256); return r; };
    L.rowmin = take((size_t)np * 8);
    L.count = take(cyb::align_up((size_t)no * 4, 64) + 64);      // counts, then {4 x int64 accumulators, ticket}
    L.total = o;
    return L;
}

}  // namespace

namespace cyb {
size_t lap_check_workspace_bytes(int64_t n_persons, int64_t n_objects) {
    if (n_persons <= 0 || n_objects <= 0) return 256;
    return chk_layout(n_persons, n_objects).total;
}
}  // namespace cyb

extern "C" int cyb_lap_check_i32(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                                 const int32_t *slot_offset_dev, const int32_t *person_obj_dev,
                                 const int64_t *price_dev, int64_t *out_dev, void *workspace_dev,
                                 size_t workspace_bytes, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int64_t np = n_persons, no = n_objects;
    if (np <= 0 || np >= (1ll << kPersonBits) || no <= 0 || no >= (1ll << kPersonBits))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_check_i32: persons=%lld objects=%lld out of range",
                              (long long)np, (long long)no);
    if (!cost_dev || !person_obj_dev || !price_dev || !out_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_check_i32: null pointer argument");
    const ChkLayout L = chk_layout(np, no);
    if (workspace_bytes < L.total)
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_lap_check_i32: workspace %zu < required %zu", workspace_bytes, L.total);
    char *ws = static_cast<char *>(workspace_dev);
    long long *rowmin = reinterpret_cast<long long *>(ws + L.rowmin);
    int32_t *count = reinterpret_cast<int32_t *>(ws + L.count);
    int dev0 = 0, sms0 = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev0));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0));
    if (no <= kChkWholeMax) {
        // one memset (counts + accumulators + ticket), one pass over the matrix, one tiny publish kernel
        const size_t acc_off = cyb::align_up((size_t)no * 4, 64);
        long long *acc = reinterpret_cast<long long *>(ws + L.count + acc_off);            // 4 x int64, then the ticket
        unsigned int *done = reinterpret_cast<unsigned int *>(ws + L.count + acc_off + 32);
        CYB_CUDA_CHECK(cudaMemsetAsync(count, 0, acc_off + 64, stream));
        const size_t smem = (size_t)no * 8;
        CYB_CUDA_CHECK(cudaFuncSetAttribute(lap_rowcheck_whole_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CYB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lap_rowcheck_whole_kernel, kChkThreads, smem));
        if (per_sm < 1) per_sm = 1;
        int grid = (int)std::min<long long>((long long)sms0 * per_sm, (np + kTeams - 1) / kTeams);
        lap_rowcheck_whole_kernel<<<grid, kChkThreads, smem, stream>>>(
            cost_dev, ld, (int)np, (int)no, person_obj_dev, reinterpret_cast<const long long *>(price_dev), np + 1, count, acc,
            slot_offset_dev, reinterpret_cast<long long *>(out_dev), done, done + 1);
        CYB_CUDA_CHECK(cudaGetLastError());
        return CYB_OK;
    }
    CYB_CUDA_CHECK(cudaMemsetAsync(rowmin, 0x7F, (size_t)np * 8, stream));     // large positive sentinel
    CYB_CUDA_CHECK(cudaMemsetAsync(count, 0, (size_t)no * 4, stream));
    const long long out_init[4] = {LLONG_MIN, 0, 0, 0};
    CYB_CUDA_CHECK(cudaMemcpyAsync(out_dev, out_init, sizeof(out_init), cudaMemcpyHostToDevice, stream));
    int dev = 0, sms = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long S = np + 1;
    const int tiles = (int)((no + kChkTileCols - 1) / kChkTileCols);
    // enough person groups for ~8 CTAs per SM, at least 16 rows each
    int rows_per_cta = (int)((np * (long long)tiles + (long long)sms * 8 - 1) / ((long long)sms * 8));
    if (rows_per_cta < 16) rows_per_cta = 16;
    const int groups = (int)((np + rows_per_cta - 1) / rows_per_cta);
    lap_rowmin_kernel<<<dim3(tiles, groups), kChkThreads, 0, stream>>>(
        cost_dev, ld, (int)np, (int)no, reinterpret_cast<const long long *>(price_dev), S, rows_per_cta, rowmin);
    CYB_CUDA_CHECK(cudaGetLastError());
    lap_check_finish_kernel<<<sms, 256, 0, stream>>>(cost_dev, ld, (int)np, (int)no, person_obj_dev,
                                                     reinterpret_cast<const long long *>(price_dev), S, rowmin, count,
                                                     reinterpret_cast<long long *>(out_dev));
    CYB_CUDA_CHECK(cudaGetLastError());
    lap_check_capacity_kernel<<<sms, 256, 0, stream>>>(slot_offset_dev, (int)no, count,
                                                       reinterpret_cast<long long *>(out_dev));
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}
