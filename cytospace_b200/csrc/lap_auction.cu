// Exact dense linear assignment on sm_100a: synchronous (Jacobi) eps-scaling
// auction in ONE persistent cooperative kernel, one CTA per SM.
//
// Replaces the third-party `lapjv.lapjv(cost)` call CytoSPACE makes at
// cytospace/linear_assignment_solvers/linear_assignment_solvers.py:38 (from
// cytospace/cytospace.py:329).  Rows = spot slots, columns = cells; the
// `cost[location_repeat, :]` expansion of linear_assignment_solvers.py:63-66 is
// never materialised -- LAP row i reads compact row row_map[i].
//
// Algorithm (min-cost form).  C = (cost - cmin) * (n+1) >= 0, prices p >= 0,
// h(i,j) = C[i,j] + p[j].  A free row i bids for j1 = argmin_j h (lowest j on
// ties) at price p[j1] + (second-min h - min h) + eps; per column the highest
// bid wins (lowest row on ties), the previous owner becomes free.  eps is
// divided by 8 per phase down to 1; because costs carry the factor n+1, the
// eps = 1 phase ends with an assignment whose total is < n+1 scaled units from
// optimal, i.e. optimal for the integer matrix.  At each phase start the pairs
// that already satisfy eps-CS for the new eps are kept.
//
// Execution model.  Every round costs exactly one grid barrier:
//   [bid r]      CTA b scans the rows at positions k = b (mod G) of the free
//                list: one coalesced pass over the row in HBM against the
//                price vector held in shared memory (or L2 when n is too big),
//                warp-shuffle + shared-memory reduction of (min, 2nd min,
//                argmin), then one 64-bit atomicMax of (bid | ~row) on the
//                column's bid slot and a record (column, previous owner).
//   barrier
//   [resolve r]  EVERY CTA replays all F records (they are tiny and L2
//                resident): winners update the CTA's private shared-memory
//                price replica, the column owner (identical values written by
//                all CTAs, so each CTA's own view is complete without a second
//                barrier) and the next free list (a block-wide prefix sum gives
//                every CTA the same positions; CTA b keeps positions = b mod G
//                as its next work queue).
// Buffers touched by atomics / records / lists alternate by round parity, so a
// fast CTA bidding in round r+1 never disturbs a slow CTA resolving round r.
// Bid slots are never reset: a stale slot value is <= the column's current
// price and every new bid is strictly greater.
//
// HBM traffic: each bid reads one row (n*4 bytes) once; prices, owners, lists
// and slots live in shared memory / L2.

#include <climits>
#include <cstdint>
#include <cstdlib>

#include "common.h"

namespace {

constexpr int kThreads = 1024;
constexpr int kRowBits = 18;                                   // n < 2^18
constexpr unsigned long long kRowMask = (1ull << kRowBits) - 1;
constexpr long long kInf = 0x3FFFFFFFFFFFFFFFll;
constexpr long long kBidLimit = 1ll << 45;                     // 46-bit bid field
constexpr int kTheta = 8;
constexpr int kEps0Div = 4;
constexpr int kTailMax = 64;                                   // capacity of the tail FIFO

struct LapParams {
    const int32_t *cost;
    long long ld;
    int n;
    const int32_t *row_map;
    int32_t *rowsol;
    int32_t *owner;          // == colsol output
    long long *price;
    long long *total;
    long long *stats;
    int32_t *list[2];
    int2 *rec[2];
    unsigned long long *slot[2];
    int32_t *flag;
    unsigned int *bar;
    int *gmm;                // [0] cmin, [1] cmax, [2] status
    int qcap;
    long long max_rounds;
    int tail_t;              // rounds with <= tail_t bidders are finished by CTA 0 alone (Gauss-Seidel tail)
};

struct Best {
    long long b1, b2;
    int j1;
};

__device__ __forceinline__ int4 ld_stream(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void upd(Best &s, long long h, int j) {
    if (h < s.b2) {
        if (h < s.b1) { s.b2 = s.b1; s.b1 = h; s.j1 = j; }
        else s.b2 = h;
    }
}

// Merge the summaries of two disjoint column sets.
__device__ __forceinline__ Best combine(const Best &a, const Best &b) {
    const bool bwins = (b.b1 < a.b1) || (b.b1 == a.b1 && (unsigned)b.j1 < (unsigned)a.j1);
    Best r;
    if (bwins) { r.b1 = b.b1; r.j1 = b.j1; r.b2 = a.b1 < b.b2 ? a.b1 : b.b2; }
    else       { r.b1 = a.b1; r.j1 = a.j1; r.b2 = b.b1 < a.b2 ? b.b1 : a.b2; }
    return r;
}

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &target, unsigned int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += G;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

// Block-wide exclusive prefix count of `valid`; `total` = number of valid threads.
__device__ __forceinline__ int block_excl_count(bool valid, int *wcnt, int &total) {
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int within = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[w] = __popc(m);
    __syncthreads();
    const int c = wcnt[lane];
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    const int woff = __shfl_sync(0xffffffffu, inc - c, w);
    total = __shfl_sync(0xffffffffu, inc, 31);
    __syncthreads();
    return woff + within;
}

// CTA-wide scan of one LAP row: min / second-min / argmin of (c-cmin)*S + p.
// The result is valid in thread 0.
template <bool SMEMP>
__device__ __forceinline__ Best scan_row(const int32_t *__restrict__ r, int n, int cmin, long long S,
                                         const long long *__restrict__ price, bool vec_ok,
                                         long long *red_b1, long long *red_b2, int *red_j) {
    Best s{kInf, kInf, -1};
    const int t = threadIdx.x;
    int jtail = 0;
    if (vec_ok) {
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
        const int n4 = n >> 2;
#pragma unroll 4
        for (int q = t; q < n4; q += kThreads) {
            const int4 c = ld_stream(r4 + q);
            const int j = q << 2;
            long long p0, p1, p2, p3;
            if (SMEMP) {
                const longlong2 a = *reinterpret_cast<const longlong2 *>(price + j);
                const longlong2 b = *reinterpret_cast<const longlong2 *>(price + j + 2);
                p0 = a.x; p1 = a.y; p2 = b.x; p3 = b.y;
            } else {
                const longlong2 a = __ldcg(reinterpret_cast<const longlong2 *>(price + j));
                const longlong2 b = __ldcg(reinterpret_cast<const longlong2 *>(price + j + 2));
                p0 = a.x; p1 = a.y; p2 = b.x; p3 = b.y;
            }
            upd(s, (long long)(c.x - cmin) * S + p0, j);
            upd(s, (long long)(c.y - cmin) * S + p1, j + 1);
            upd(s, (long long)(c.z - cmin) * S + p2, j + 2);
            upd(s, (long long)(c.w - cmin) * S + p3, j + 3);
        }
        jtail = n4 << 2;
    }
    for (int j = jtail + t; j < n; j += kThreads) {
        const long long p = SMEMP ? price[j] : __ldcg(price + j);
        upd(s, (long long)(__ldg(r + j) - cmin) * S + p, j);
    }
    // NOTE: in the vectorised loop a thread's columns are not globally
    // increasing against the tail loop, but tail columns are all larger than
    // vector columns, and `upd` keeps the earlier (lower) column on ties.
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Best o;
        o.b1 = __shfl_xor_sync(0xffffffffu, s.b1, d);
        o.b2 = __shfl_xor_sync(0xffffffffu, s.b2, d);
        o.j1 = __shfl_xor_sync(0xffffffffu, s.j1, d);
        s = combine(s, o);
    }
    const int lane = t & 31, w = t >> 5;
    if (lane == 0) { red_b1[w] = s.b1; red_b2[w] = s.b2; red_j[w] = s.j1; }
    __syncthreads();
    if (w == 0) {
        s.b1 = red_b1[lane]; s.b2 = red_b2[lane]; s.j1 = red_j[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            Best o;
            o.b1 = __shfl_xor_sync(0xffffffffu, s.b1, d);
            o.b2 = __shfl_xor_sync(0xffffffffu, s.b2, d);
            o.j1 = __shfl_xor_sync(0xffffffffu, s.j1, d);
            s = combine(s, o);
        }
    }
    __syncthreads();   // red_* may be reused by the next scan
    return s;
}

template <bool SMEMP>
__global__ void __launch_bounds__(kThreads, 1) lap_auction_kernel(const LapParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n;
    long long *sprice = reinterpret_cast<long long *>(smem_raw);
    size_t off = SMEMP ? ((size_t)n * 8 + 15) / 16 * 16 : 0;
    int *myq = reinterpret_cast<int *>(smem_raw + off);
    __shared__ long long red_b1[32], red_b2[32];
    __shared__ int red_j[32], wcnt[32];
    __shared__ int tq[kTailMax], tq_head, tq_cnt, tq_status;

    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    const long long S = (long long)n + 1;
    const bool vec_ok = ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.cost) & 15) == 0);
    unsigned int bar_target = 0;
    const long long *price_rd = SMEMP ? sprice : P.price;

    auto rowptr = [&](int i) -> const int32_t * {
        const long long r = P.row_map ? (long long)__ldg(P.row_map + i) : (long long)i;
        return P.cost + r * P.ld;
    };

    // ---- pass 0: state init and the cost range ------------------------------
    {
        int lmin = INT_MAX, lmax = INT_MIN;
        for (int i = b; i < n; i += G) {
            const int32_t *r = rowptr(i);
            for (int j = t; j < n; j += kThreads) {
                const int c = __ldg(r + j);
                lmin = min(lmin, c); lmax = max(lmax, c);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, d));
            lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
        }
        if ((t & 31) == 0 && lmin <= lmax) { atomicMin(P.gmm + 0, lmin); atomicMax(P.gmm + 1, lmax); }
        for (int j = b * kThreads + t; j < n; j += G * kThreads) {
            P.price[j] = 0; P.owner[j] = -1; P.rowsol[j] = -1;
            P.slot[0][j] = 0ull; P.slot[1][j] = 0ull;
        }
        if (SMEMP) for (int j = t; j < n; j += kThreads) sprice[j] = 0;
    }
    grid_barrier(P.bar, bar_target, G);
    const int cmin = __ldcg(P.gmm + 0), cmax = __ldcg(P.gmm + 1);
    long long eps = ((long long)cmax - (long long)cmin) * S / kEps0Div;
    if (eps < 1) eps = 1;

    long long rounds = 0, bids = 0, passes = 0, phases = 0, rounds1 = 0, maxF = 0, tail_bids = 0, tails = 0;
    int status = 0, par = 0;
    const int tail_t = min(P.tail_t, kTailMax);

    for (;;) {
        ++phases;
        // ---- phase start: which pairs survive eps-CS at the new eps? --------
        for (int i = b; i < n; i += G) {
            const int j0 = __ldcg(P.rowsol + i);
            int f = 1;
            if (j0 >= 0) {
                const int32_t *r = rowptr(i);
                const Best s = scan_row<SMEMP>(r, n, cmin, S, price_rd, vec_ok, red_b1, red_b2, red_j);
                if (t == 0) {
                    const long long pj = SMEMP ? sprice[j0] : __ldcg(P.price + j0);
                    const long long h0 = (long long)(__ldg(r + j0) - cmin) * S + pj;
                    f = (h0 > s.b1 + eps) ? j0 + 2 : 0;
                }
            }
            if (t == 0) P.flag[i] = f;
        }
        ++passes;
        grid_barrier(P.bar, bar_target, G);
        int F = 0;
        for (int i0 = 0; i0 < n; i0 += kThreads) {
            const int i = i0 + t;
            const int f = i < n ? __ldcg(P.flag + i) : 0;
            if (f >= 2) {
                P.owner[f - 2] = -1;                       // identical write from every CTA
                if (i % G == b) P.rowsol[i] = -1;
            }
            int tot;
            const int pos = F + block_excl_count(f != 0, wcnt, tot);
            if (f != 0 && pos % G == b) { P.list[par][pos] = i; myq[pos / G] = i; }
            F += tot;
        }
        __syncthreads();

        // ---- bidding rounds --------------------------------------------------
        bool ran_tail = false;
        while (F > 0) {
            if (F <= tail_t) {
                // ---- Gauss-Seidel tail: few bidders left, a grid barrier per round would cost more
                // than the bids.  CTA 0 alone drains a FIFO of free rows; every bid sees the prices
                // the previous one left and the lone bidder always wins, so nothing is exchanged
                // until the phase ends.  (Same auction, sequential order: still eps-CS, still exact.)
                grid_barrier(P.bar, bar_target, G);          // every CTA has finished its resolve writes
                ran_tail = true; ++tails;
                if (b == 0) {
                    if (t < F) tq[t] = __ldcg(P.list[par] + t);
                    if (t == 0) { tq_head = 0; tq_cnt = F; tq_status = 0; }
                    __syncthreads();
                    while (tq_cnt > 0 && tq_status == 0) {
                        const int i = tq[tq_head];
                        const int32_t *r = rowptr(i);
                        const Best s = scan_row<SMEMP>(r, n, cmin, S, price_rd, vec_ok, red_b1, red_b2, red_j);
                        ++tail_bids;
                        if (t == 0) {
                            const long long pj = SMEMP ? sprice[s.j1] : __ldcg(P.price + s.j1);
                            const long long bid = pj + (n > 1 ? s.b2 - s.b1 : 0) + eps;
                            if (bid >= kBidLimit) tq_status = CYB_ERR_OVERFLOW;
                            if (tail_bids + rounds > P.max_rounds) tq_status = CYB_ERR_NOT_CONVERGED;
                            const int prev = __ldcg(P.owner + s.j1);
                            P.owner[s.j1] = i; P.rowsol[i] = s.j1; P.price[s.j1] = bid;
                            if (SMEMP) sprice[s.j1] = bid;
                            int head = tq_head + 1; if (head == kTailMax) head = 0;
                            int cnt = tq_cnt - 1;
                            if (prev >= 0) {
                                P.rowsol[prev] = -1;
                                int tail = head + cnt; if (tail >= kTailMax) tail -= kTailMax;
                                tq[tail] = prev; ++cnt;
                            }
                            tq_head = head; tq_cnt = cnt;
                        }
                        __syncthreads();
                    }
                    if (t == 0 && tq_status) atomicExch(P.gmm + 2, tq_status);
                }
                F = 0;
                break;
            }
            if (++rounds > P.max_rounds) { status = CYB_ERR_NOT_CONVERGED; break; }
            bids += F;
            if (F <= 1) ++rounds1;
            if (F > maxF) maxF = F;
            const int myn = F > b ? (F - b - 1) / G + 1 : 0;
            for (int q = 0; q < myn; ++q) {
                const int i = myq[q];
                const Best s = scan_row<SMEMP>(rowptr(i), n, cmin, S, price_rd, vec_ok, red_b1, red_b2, red_j);
                if (t == 0) {
                    const long long gamma = (n > 1 ? s.b2 - s.b1 : 0) + eps;
                    const long long pj = SMEMP ? sprice[s.j1] : __ldcg(P.price + s.j1);
                    const long long bid = pj + gamma;
                    if (bid >= kBidLimit) atomicExch(P.gmm + 2, CYB_ERR_OVERFLOW);
                    const int prev = __ldcg(P.owner + s.j1);
                    P.rec[par][q * G + b] = make_int2(s.j1, prev);
                    atomicMax(P.slot[par] + s.j1,
                              ((unsigned long long)bid << kRowBits) | (kRowMask - (unsigned long long)i));
                }
            }
            grid_barrier(P.bar, bar_target, G);
            status = __ldcg(P.gmm + 2);
            if (status) break;
            // ---- resolve: every CTA replays every record ----------------------
            int Fn = 0;
            for (int k0 = 0; k0 < F; k0 += kThreads) {
                const int k = k0 + t;
                int entry = -1;
                if (k < F) {
                    const int i = __ldcg(P.list[par] + k);
                    const int2 rc = __ldcg(P.rec[par] + k);
                    const unsigned long long key = __ldcg(P.slot[par] + rc.x);
                    const int wrow = (int)(kRowMask - (key & kRowMask));
                    if (wrow == i) {
                        const long long bid = (long long)(key >> kRowBits);
                        const bool mine = (k % G == b);
                        if (SMEMP) { sprice[rc.x] = bid; if (mine) P.price[rc.x] = bid; }
                        else P.price[rc.x] = bid;              // identical write from every CTA
                        P.owner[rc.x] = i;                     // identical write from every CTA
                        if (mine) { P.rowsol[i] = rc.x; if (rc.y >= 0) P.rowsol[rc.y] = -1; }
                        entry = rc.y;
                    } else {
                        entry = i;
                    }
                }
                int tot;
                const int pos = Fn + block_excl_count(entry >= 0, wcnt, tot);
                if (entry >= 0 && pos % G == b) { P.list[par ^ 1][pos] = entry; myq[pos / G] = entry; }
                Fn += tot;
            }
            F = Fn;
            par ^= 1;
            __syncthreads();
        }
        if (status) break;
        grid_barrier(P.bar, bar_target, G);        // rowsol of the last resolve / the tail becomes visible
        if (ran_tail) {
            status = __ldcg(P.gmm + 2);
            if (status) break;
            if (SMEMP) {                            // pick up the prices CTA 0 moved during the tail
                for (int j = t; j < n; j += kThreads) sprice[j] = __ldcg(P.price + j);
                __syncthreads();
            }
        }
        if (eps == 1) break;
        eps /= kTheta;
        if (eps < 1) eps = 1;
    }

    // ---- total cost of the assignment -------------------------------------------
    if (!status) {
        long long sum = 0;
        for (int i = b * kThreads + t; i < n; i += G * kThreads) {
            const int j = __ldcg(P.rowsol + i);
            if (j >= 0) sum += (long long)__ldg(rowptr(i) + j);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if ((t & 31) == 0 && sum != 0) atomicAdd(reinterpret_cast<unsigned long long *>(P.total), (unsigned long long)sum);
    }
    if (b == 0 && t == 0) {
        P.stats[0] = status; P.stats[1] = phases; P.stats[2] = rounds; P.stats[3] = bids;
        P.stats[4] = passes; P.stats[5] = cmin; P.stats[6] = cmax; P.stats[7] = S;
        P.stats[8] = G; P.stats[9] = SMEMP ? 1 : 0; P.stats[10] = rounds1; P.stats[11] = maxF;
        P.stats[12] = (phases - 1) * (long long)n; P.stats[13] = tail_bids; P.stats[14] = tails; P.stats[15] = 0;
    }
}

// ---------------------------------------------------------------------------------
// Certificate / row-scan pass.  Grid = (column tiles, row groups); a CTA stages a
// tile of prices in shared memory once and streams `rows_per_cta` rows against
// it, one warp per row, so the cost matrix is read exactly once from HBM and the
// price vector once per row group from L2.
constexpr int kChkThreads = 512;
constexpr int kChkTileCols = 4096;

__global__ void __launch_bounds__(kChkThreads) lap_rowmin_kernel(
    const int32_t *__restrict__ cost, long long ld, int n, const int32_t *__restrict__ row_map,
    const long long *__restrict__ price, long long S, int rows_per_cta,
    long long *__restrict__ rowmin) {
    __shared__ __align__(16) long long sp[kChkTileCols];
    const int c0 = blockIdx.x * kChkTileCols;
    const int nc = min(kChkTileCols, n - c0);
    for (int j = threadIdx.x; j < nc; j += kChkThreads) sp[j] = price[c0 + j];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(n, r0 + rows_per_cta);
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
    for (int i = r0 + w; i < r1; i += kChkThreads / 32) {
        const long long rr = row_map ? (long long)row_map[i] : (long long)i;
        const int32_t *r = cost + rr * ld + c0;
        long long m = kInf;
        int jt = 0;
        if (vec_ok) {
            const int4 *r4 = reinterpret_cast<const int4 *>(r);
            const int n4 = nc >> 2;
#pragma unroll 4
            for (int q = lane; q < n4; q += 32) {
                const int4 c = ld_stream(r4 + q);
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

__global__ void lap_check_finish_kernel(const int32_t *__restrict__ cost, long long ld, int n,
                                        const int32_t *__restrict__ row_map,
                                        const int32_t *__restrict__ rowsol,
                                        const long long *__restrict__ price, long long S,
                                        const long long *__restrict__ rowmin, long long *out) {
    long long viol = LLONG_MIN, tot = 0, bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = rowsol[i];
        if (j < 0 || j >= n) { ++bad; continue; }
        const long long rr = row_map ? (long long)row_map[i] : (long long)i;
        const int c = cost[rr * ld + j];
        tot += c;
        const long long h = (long long)c * S + price[j];
        viol = max(viol, h - rowmin[i]);
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

struct WsLayout {
    size_t list0, list1, rec0, rec1, slot0, slot1, flag, rowmin, small, total;
};

WsLayout ws_layout(int64_t n) {
    WsLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = cyb::align_up(o + bytes, 256); return r; };
    L.list0 = take((size_t)n * 4); L.list1 = take((size_t)n * 4);
    L.rec0 = take((size_t)n * 8);  L.rec1 = take((size_t)n * 8);
    L.slot0 = take((size_t)n * 8); L.slot1 = take((size_t)n * 8);
    L.flag = take((size_t)n * 4);
    L.rowmin = take((size_t)n * 8);
    L.small = take(256);
    L.total = o;
    return L;
}

}  // namespace

extern "C" size_t cyb_lap_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return ws_layout(n).total;
}

extern "C" int cyb_lap_solve_i32(const int32_t *cost_dev, int64_t ld, int64_t n,
                                 const int32_t *row_map_dev, int32_t *rowsol_dev,
                                 int32_t *colsol_dev, int64_t *price_dev, int64_t *total_dev,
                                 int64_t *stats_dev, void *workspace_dev, size_t workspace_bytes,
                                 int grid_hint, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (n <= 0 || n >= (1ll << kRowBits))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: n=%lld outside [1, 2^18)", (long long)n);
    if (!cost_dev || !rowsol_dev || !colsol_dev || !price_dev || !total_dev || !stats_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: null pointer argument");
    if (ld < n) return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: ld=%lld < n=%lld", (long long)ld, (long long)n);
    const WsLayout L = ws_layout(n);
    if (workspace_bytes < L.total)
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_lap_solve_i32: workspace %zu < required %zu", workspace_bytes, L.total);
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 255)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: workspace must be 256-byte aligned");

    int dev = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    int sms = 0, coop = 0, max_smem = 0;
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (!coop) return cyb::set_error(CYB_ERR_UNSUPPORTED, "device lacks cooperative launch");

    int G = grid_hint > 0 ? grid_hint : sms;
    if (G > sms) G = sms;
    if (G > n) G = (int)n;
    if (G < 1) G = 1;

    char *ws = static_cast<char *>(workspace_dev);
    LapParams P;
    P.cost = cost_dev; P.ld = ld; P.n = (int)n; P.row_map = row_map_dev;
    P.rowsol = rowsol_dev; P.owner = colsol_dev; P.price = reinterpret_cast<long long *>(price_dev);
    P.total = reinterpret_cast<long long *>(total_dev); P.stats = reinterpret_cast<long long *>(stats_dev);
    P.list[0] = reinterpret_cast<int32_t *>(ws + L.list0); P.list[1] = reinterpret_cast<int32_t *>(ws + L.list1);
    P.rec[0] = reinterpret_cast<int2 *>(ws + L.rec0); P.rec[1] = reinterpret_cast<int2 *>(ws + L.rec1);
    P.slot[0] = reinterpret_cast<unsigned long long *>(ws + L.slot0);
    P.slot[1] = reinterpret_cast<unsigned long long *>(ws + L.slot1);
    P.flag = reinterpret_cast<int32_t *>(ws + L.flag);
    P.bar = reinterpret_cast<unsigned int *>(ws + L.small);
    P.gmm = reinterpret_cast<int *>(ws + L.small + 16);
    P.qcap = (int)((n + G - 1) / G);
    P.max_rounds = 2000ll * n + 100000;
    P.tail_t = 2;
    if (const char *e = getenv("CYB_LAP_TAIL")) P.tail_t = atoi(e);

    // small block: barrier counter, cmin = INT_MAX, cmax = INT_MIN, status = 0
    const int init[8] = {0, 0, 0, 0, INT_MAX, INT_MIN, 0, 0};
    CYB_CUDA_CHECK(cudaMemcpyAsync(ws + L.small, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    CYB_CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), stream));

    const size_t q_bytes = cyb::align_up((size_t)P.qcap * 4, 16);
    const size_t smem_with_price = cyb::align_up((size_t)n * 8, 16) + q_bytes;
    const size_t static_smem = 32 * 8 * 2 + 32 * 4 * 2 + 64;
    const bool smemp = smem_with_price + static_smem <= (size_t)max_smem;
    const size_t dyn = smemp ? smem_with_price : q_bytes;
    const void *fn = smemp ? (const void *)lap_auction_kernel<true> : (const void *)lap_auction_kernel<false>;
    CYB_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int occ = 0;
    CYB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, dyn));
    if (occ < 1) return cyb::set_error(CYB_ERR_UNSUPPORTED, "lap kernel does not fit on an SM (smem %zu)", dyn);
    if (G > occ * sms) G = occ * sms;
    void *args[] = {(void *)&P};
    CYB_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kThreads), args, dyn, stream));
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}

extern "C" int cyb_lap_check_i32(const int32_t *cost_dev, int64_t ld, int64_t n,
                                 const int32_t *row_map_dev, const int32_t *rowsol_dev,
                                 const int64_t *price_dev, int64_t *out_dev, void *workspace_dev,
                                 size_t workspace_bytes, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (n <= 0 || n >= (1ll << kRowBits))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_check_i32: n=%lld outside [1, 2^18)", (long long)n);
    if (!cost_dev || !rowsol_dev || !price_dev || !out_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_check_i32: null pointer argument");
    const WsLayout L = ws_layout(n);
    if (workspace_bytes < L.total)
        return cyb::set_error(CYB_ERR_WORKSPACE, "cyb_lap_check_i32: workspace %zu < required %zu", workspace_bytes, L.total);
    char *ws = static_cast<char *>(workspace_dev);
    long long *rowmin = reinterpret_cast<long long *>(ws + L.rowmin);
    CYB_CUDA_CHECK(cudaMemsetAsync(rowmin, 0x7F, (size_t)n * 8, stream));     // large positive sentinel
    const long long out_init[3] = {LLONG_MIN, 0, 0};
    CYB_CUDA_CHECK(cudaMemcpyAsync(out_dev, out_init, sizeof(out_init), cudaMemcpyHostToDevice, stream));
    int dev = 0, sms = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long S = n + 1;
    const int tiles = (int)((n + kChkTileCols - 1) / kChkTileCols);
    // enough row groups for >= 4 waves of CTAs, at least 16 rows each
    int rows_per_cta = (int)((n * (long long)tiles + (long long)sms * 8 - 1) / ((long long)sms * 8));
    if (rows_per_cta < 16) rows_per_cta = 16;
    const int groups = (int)((n + rows_per_cta - 1) / rows_per_cta);
    lap_rowmin_kernel<<<dim3(tiles, groups), kChkThreads, 0, stream>>>(
        cost_dev, ld, (int)n, row_map_dev, reinterpret_cast<const long long *>(price_dev), S,
        rows_per_cta, rowmin);
    CYB_CUDA_CHECK(cudaGetLastError());
    lap_check_finish_kernel<<<sms, 256, 0, stream>>>(cost_dev, ld, (int)n, row_map_dev, rowsol_dev,
                                                     reinterpret_cast<const long long *>(price_dev), S,
                                                     rowmin, reinterpret_cast<long long *>(out_dev));
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}
