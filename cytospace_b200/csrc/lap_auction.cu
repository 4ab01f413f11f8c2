// Exact dense linear assignment (transportation form) on sm_100a: synchronous
// (Jacobi) eps-scaling auction with capacitated objects in ONE persistent
// cooperative kernel, one CTA per SM.
//
// Replaces the third-party `lapjv.lapjv(cost)` call CytoSPACE makes at
// cytospace/linear_assignment_solvers/linear_assignment_solvers.py:38 (from
// cytospace/cytospace.py:329) on the expanded matrix `cost[location_repeat, :]`
// (linear_assignment_solvers.py:63-66).  The expansion is never materialised:
// PERSONS are the cells (rows of M = cost^T, cells x spots), OBJECTS are the
// spots, object o has cap[o] = cell_number_to_node_assignment[o] SLOTS.
//     min sum_i M[i, obj(i)]   s.t.  object o holds exactly cap[o] persons
// is the same optimisation problem as the reference's square LAP (identical
// optimal total); duplicated spot rows become capacity instead of price wars.
//
// Algorithm.  C = (M - cmin) * (P+1) >= 0.  Every slot has a price (its last
// accepted bid) and a holder; the object's price lambda[o] is its cheapest
// slot.  A free person i scans its row: v1 = min_o (C[i,o] + lambda[o]) at o*
// (lowest o on ties), w = the minimum over o != o*; it bids
// b = lambda[o*] + (w - v1) + eps for the cheapest slot of o*.  Per object the
// highest bid of a round wins (lowest person on ties) and evicts the slot's
// holder.  Invariant (eps-CS): an assigned person's C[i,o] + (own slot price) is
// within eps of its best alternative object, and lambda[o] <= own slot price.
// eps is divided by 4 per phase down to 1; at a phase start a pair is kept iff
// C[i,o] + lambda[o] <= alt + eps, and its slot price is clamped down to
// alt + eps - C[i,o] (never below lambda[o], so object prices never decrease).
// With the factor P+1 on the costs, eps = 1 leaves the total < P+1 scaled units
// from optimal, i.e. optimal for the integer matrix (DESIGN.md has the proof).
//
// Execution model.  Every bidding round costs exactly one grid barrier:
//   [bid r]      CTA b scans the rows at positions k = b (mod G) of the free
//                list: one coalesced pass over the row in HBM against the
//                object prices held in shared memory (L2 when O is too big),
//                warp-shuffle + shared-memory reduction of (min, 2nd min,
//                argmin), one 64-bit atomicMax of (bid | ~person) on the
//                object's bid word and a record (object, slot, slot holder).
//   barrier
//   [resolve r]  EVERY CTA replays all F records (tiny, L2 resident): winners
//                update the slot, the object's cheapest slot / price (also in
//                the CTA's shared-memory replica) and the next free list (a
//                block-wide prefix sum gives every CTA the same positions;
//                CTA b keeps positions = b mod G as its next work queue).  Every
//                CTA ends the replay with the same view of the state -- in its
//                shared-memory replicas (prices, slot owners, cheapest slots)
//                when they fit, otherwise by writing the identical values to
//                global memory itself -- so no second barrier is needed; state
//                that is only read behind the next barrier has ONE writer (148
//                CTAs storing to one line serialise at the L2 slice).
// Lists, records and bid words rotate over three buffers by round: a fast CTA
// bidding in round r+1 never disturbs a slow CTA resolving round r, and the bid
// words of round r-1 are cleared during resolve r (their next use is round r+2).
// Tail.  A round costs ~7 us whatever its size and most rounds have one or two
// bidders (long eviction chains), so when few bidders are left CTA 0 finishes
// the phase alone while the other CTAs ("sweepers") keep per-person candidate
// lists fresh -- a list certifies the exact result of a row scan from a handful
// of entries.  Two tails, a compile-time choice (template parameter TM) made on
// the host from what fits in shared memory:
//   TM = 0  Gauss-Seidel FIFO (<= 8 bidders): one bid at a time by warp 0, each
//           seeing the prices the previous one left; used when prices AND slot
//           owners are shared-memory resident (a step = one L2 round trip).
//   TM = 1  Jacobi rounds inside CTA 0 (<= 32 bidders): every warp owns one
//           eviction chain, all bids of a round see the same prices, winners are
//           resolved from the posted bids; same round semantics as the grid, so
//           the result does not depend on where the switch happens.  Used when
//           prices or owners live in L2 (25k, 50k): parallel chains hide the
//           two or three chained L2 round trips of a step.
// The other CTAs pick up prices / owners at the phase end.
//
// HBM traffic: each bid reads one row (O*4 bytes) once; prices, holders, lists
// and bid words live in shared memory / L2.
//
// Tuning knobs (environment, read at launch; defaults are the measured best):
// CYB_LAP_TAIL_MODE, CYB_LAP_TAIL, CYB_LAP_THETA, CYB_LAP_EPS0, CYB_LAP_LISTS,
// CYB_LAP_SWEEPERS, CYB_LAP_PACKED, CYB_LAP_PREFETCH, CYB_LAP_SMEM_OWNER, and
// the experiments that did not pay (off): CYB_LAP_EARLY, CYB_LAP_APPROX;
// CYB_LAP_SMEM_PRICES=0 forces the L2-price code paths (tests).

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>

#include "common.h"

namespace cyb {
int lap_solve_auction(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                      const int32_t *slot_offset_dev, int32_t *person_obj_dev, int32_t *slot_owner_dev,
                      int64_t *price_dev, int64_t *total_dev, int64_t *stats_dev, void *workspace_dev,
                      size_t workspace_bytes, int grid_hint, void *stream_v);
size_t lap_auction_workspace_bytes(int64_t n_persons, int64_t n_objects);
}  // namespace cyb

namespace {

constexpr int kThreads = 1024;
constexpr int kPersonBits = 18;                                // P < 2^18
constexpr unsigned long long kPersonMask = (1ull << kPersonBits) - 1;
constexpr long long kInf = 1ll << 60;                          // price of an object without capacity
constexpr long long kBidLimit = 1ll << 45;                     // 46-bit bid field
constexpr int kTheta = 4;
constexpr int kEps0Div = 4;
constexpr int kTailMax = 64;                                   // capacity of the tail FIFO
constexpr int kListK = 128;                                    // candidate-list capacity per person
constexpr unsigned kGenMask = 0x3FFF;                          // 14-bit generation tag above the 18-bit object index

struct LapParams {
    const int32_t *cost;     // M: persons x objects
    long long ld;
    int P, O;
    const int32_t *soff;     // slot offsets [O+1] (nullptr: every capacity is 1)
    int32_t *person_obj;     // out: object of person i
    int32_t *slot_owner;     // out: person holding slot t (slots ordered by object)
    long long *lambda;       // out: object prices (scaled by P+1)
    long long *total;
    long long *stats;
    long long *slot_price;
    int32_t *person_slot;
    int32_t *minslot;
    int32_t *list[3];
    int4 *rec[3];            // (object, slot, holder of that slot, -)
    unsigned long long *bidw[3];
    int32_t *flag;
    unsigned int *bar;
    int *gmm;                // [0] cmin, [1] cmax, [2] status
    int qcap;
    long long max_rounds;
    int tail_t;              // rounds with <= tail_t bidders are finished by CTA 0 alone
    // candidate lists (tail accelerator): per person a header {bound, gen<<32 | n} and kListK entries
    // (object | gen << 18, cost).  Every object NOT in the list had value >= bound when the list was
    // built; prices never decrease, so that stays true and a list evaluation whose best two values are
    // < bound returns exactly what a full row scan would.
    longlong2 *lst_hdr;
    int2 *lst_ent;
    int use_lists;
    int sweepers;            // CTAs that rebuild lists during a tail (<= G-1)
    int theta, eps0_div;     // eps schedule: eps0 = range*(P+1)/eps0_div, eps /= theta per phase
    int prefetch;            // 1: bulk L2 prefetch of the next row to scan
    int approx;              // 1 (only without shared-memory prices): 32-bit price prefixes in shared memory, scan_row_approx
    int packed_reduce;       // 1: scan_row reduces packed (value, column) keys with REDUX when the range allows
    int tail_mode;           // 0: Gauss-Seidel FIFO tail, 1: Jacobi rounds inside CTA 0 (one warp per bidder)
    int early_stop;          // a phase with eps > 1 ends once <= early_stop persons are free (they bid again next phase)
    int smem_owner;          // 1: every CTA keeps a replica of slot_owner (and minslot) in shared memory
};

struct Best {
    long long b1, b2;
    int j1;
};

// Row data is streamed: no L1 allocation, and L2 lines marked evict-first so that the matrix
// stream does not push the (hot, small) prices / holders / candidate lists out of L2.
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

// Ask the memory system for a whole row ahead of its scan: bulk L2 prefetches issued by a few threads
// (no registers, no shared memory).  The scan's own loads then hit L2 instead of waiting on HBM three
// dependent batches in a row (a 1024-thread CTA keeps only 64 KB in flight; a 50k row is 200 KB).
__device__ __forceinline__ void prefetch_row_l2(const int32_t *r, int n) {
    const unsigned bytes = ((unsigned)n * 4u) & ~15u;
    const unsigned chunk = 8192u;
    const unsigned off = threadIdx.x * chunk;
    if (off < bytes && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
        const unsigned len = min(chunk, bytes - off);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char *>(r) + off), "r"(len) : "memory");
    }
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

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
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

// Warp minimum of a 64-bit key with two 32-bit REDUX steps (high word, then low word among the lanes
// that hold the minimal high word).
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    return ((unsigned long long)mhi << 32) | mlo;
}

// Packed keys: every finite value is < 2^46 - 1 (scaled cost range < 2^45, prices < 2^45), so
// (value << 18 | column) orders like (value, column) and the CTA-wide (min, argmin, second min) is four
// REDUX per level instead of ten 64-bit shuffle rounds; priced-out objects clamp to the all-ones value.
constexpr long long kPackMax = (1ll << 46) - 1;
__device__ __forceinline__ unsigned long long pack_key(long long v, unsigned j) {
    return ((unsigned long long)min(v, kPackMax) << kPersonBits) | j;
}
// k1 = the thread's best key, k2 = its second best (only the value part is used); result valid in warp 0.
__device__ __forceinline__ Best packed_reduce(unsigned long long k1, unsigned long long k2, long long *red_b1,
                                              long long *red_b2) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const unsigned long long w1 = warp_min64(k1);
    const unsigned long long w2 = warp_min64(k1 == w1 ? k2 : k1);
    unsigned long long *rk1 = reinterpret_cast<unsigned long long *>(red_b1);
    unsigned long long *rk2 = reinterpret_cast<unsigned long long *>(red_b2);
    if (lane == 0) { rk1[w] = w1; rk2[w] = w2; }
    __syncthreads();
    Best s{LLONG_MAX, LLONG_MAX, -1};
    if (w == 0) {
        const unsigned long long q1 = rk1[lane], q2 = rk2[lane];
        const unsigned long long W1 = warp_min64(q1);
        const unsigned long long W2 = warp_min64(q1 == W1 ? q2 : q1);
        const long long v1 = (long long)(W1 >> kPersonBits), v2 = (long long)(W2 >> kPersonBits);
        s.j1 = W1 == ~0ull ? -1 : (int)(W1 & kPersonMask);
        s.b1 = (W1 == ~0ull) ? LLONG_MAX : (v1 == kPackMax ? kInf : v1);
        s.b2 = (W2 == ~0ull) ? LLONG_MAX : (v2 == kPackMax ? kInf : v2);
    }
    __syncthreads();
    return s;
}

// ---- approximate-price scan (objects too many for 8-byte prices in shared memory) ----------------
// Shared memory holds hi[o] = price[o] >> 15 as 32 bits (50k objects: 200 KB) instead of streaming the
// 8-byte prices from L2 with every row.  With x = (c - cmin) * S:  a = (x >> 15) + hi[o]  satisfies
// a <= (x + price) >> 15 <= a + 1.  Let m2 be the second-smallest a of the row (as a multiset): a column
// with a > m2 + 1 is strictly worse than the two columns attaining the two smallest a, so the exact
// (min, argmin, second min) lies among the columns with a <= m2 + 1 -- a handful, re-evaluated exactly
// against the 64-bit prices in L2.  Every thread tracks its two best columns and the VALUE of its third;
// if some thread's third also passes the threshold the row falls back to the exact scan (returns false).
constexpr int kApproxShift = 15;
__device__ __forceinline__ unsigned price_hi(long long p) {
    return p >= (1ll << 46) ? 0x7FFFFFFFu : (unsigned)(p >> kApproxShift);
}
struct Top3 {
    unsigned a1, a2, a3;
    int j1, j2, c1, c2;
};
__device__ __forceinline__ void ins3(Top3 &s, unsigned a, int j, int c) {
    if (a < s.a3) {
        if (a < s.a2) {
            s.a3 = s.a2;
            if (a < s.a1) { s.a2 = s.a1; s.j2 = s.j1; s.c2 = s.c1; s.a1 = a; s.j1 = j; s.c1 = c; }
            else { s.a2 = a; s.j2 = j; s.c2 = c; }
        } else s.a3 = a;
    }
}
__device__ __forceinline__ bool scan_row_approx(const int32_t *__restrict__ r, int n, int cmin, int S,
                                                const unsigned *__restrict__ sp32, const long long *__restrict__ price,
                                                bool vec_ok, long long *red_b1, long long *red_b2, int *red_j, Best &out) {
    Top3 s{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, -1, -1, 0, 0};
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const unsigned long long pol = l2_policy_evict_first();
    int jtail = 0;
    if (vec_ok) {
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
        const int n4 = n >> 2;
#pragma unroll 4
        for (int q = t; q < n4; q += kThreads) {
            const int4 c = ld_stream(r4 + q, pol);
            const uint4 h = *reinterpret_cast<const uint4 *>(sp32 + 4 * q);
            const int j = q << 2;
            ins3(s, (unsigned)(((unsigned long long)(unsigned)(c.x - cmin) * (unsigned)S) >> kApproxShift) + h.x, j, c.x);
            ins3(s, (unsigned)(((unsigned long long)(unsigned)(c.y - cmin) * (unsigned)S) >> kApproxShift) + h.y, j + 1, c.y);
            ins3(s, (unsigned)(((unsigned long long)(unsigned)(c.z - cmin) * (unsigned)S) >> kApproxShift) + h.z, j + 2, c.z);
            ins3(s, (unsigned)(((unsigned long long)(unsigned)(c.w - cmin) * (unsigned)S) >> kApproxShift) + h.w, j + 3, c.w);
        }
        jtail = n4 << 2;
    }
    for (int j = jtail + t; j < n; j += kThreads) {
        const int c = __ldg(r + j);
        ins3(s, (unsigned)(((unsigned long long)(unsigned)(c - cmin) * (unsigned)S) >> kApproxShift) + sp32[j], j, c);
    }
    // threshold = (second-smallest a of the row) + 1
    unsigned m1 = __reduce_min_sync(0xffffffffu, s.a1);
    unsigned m2 = (__popc(__ballot_sync(0xffffffffu, s.a1 == m1)) >= 2)
                      ? m1 : __reduce_min_sync(0xffffffffu, s.a1 == m1 ? s.a2 : s.a1);
    unsigned *ru = reinterpret_cast<unsigned *>(red_j);          // 32 ints: m1 per warp; red_b1 reused for m2
    unsigned *ru2 = reinterpret_cast<unsigned *>(red_b1);
    if (lane == 0) { ru[w] = m1; ru2[w] = m2; }
    __syncthreads();
    m1 = ru[lane]; m2 = ru2[lane];
    const unsigned M1 = __reduce_min_sync(0xffffffffu, m1);
    const unsigned M2 = (__popc(__ballot_sync(0xffffffffu, m1 == M1)) >= 2)
                            ? M1 : __reduce_min_sync(0xffffffffu, m1 == M1 ? m2 : m1);
    const unsigned T = M2 == 0xFFFFFFFFu ? M2 : M2 + 1u;
    // (every warp computed the same T from the same shared values: no broadcast needed)
    if (__syncthreads_or(s.a3 <= T && s.a3 != 0xFFFFFFFFu)) return false;      // a third candidate in one thread
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    if (s.j1 >= 0 && s.a1 <= T) k1 = pack_key((long long)(unsigned)(s.c1 - cmin) * S + __ldcg(price + s.j1), (unsigned)s.j1);
    if (s.j2 >= 0 && s.a2 <= T) k2 = pack_key((long long)(unsigned)(s.c2 - cmin) * S + __ldcg(price + s.j2), (unsigned)s.j2);
    if (k2 < k1) { const unsigned long long x = k1; k1 = k2; k2 = x; }
    out = packed_reduce(k1, k2, red_b1, red_b2);
    return true;
}

// CTA-wide scan of one person's row: min / second-min / argmin of (c-cmin)*S + lambda.
// The result is valid in thread 0.
template <bool SMEMP>
__device__ __forceinline__ Best scan_row(const int32_t *__restrict__ r, int n, int cmin, int S,
                                         const long long *__restrict__ price, bool vec_ok,
                                         long long *red_b1, long long *red_b2, int *red_j, bool packed) {
    Best s{LLONG_MAX, LLONG_MAX, -1};
    const int t = threadIdx.x;
    const unsigned long long pol = l2_policy_evict_first();
    int jtail = 0;
    if (vec_ok) {
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
        const int n4 = n >> 2;
#pragma unroll 4
        for (int q = t; q < n4; q += kThreads) {
            const int4 c = ld_stream(r4 + q, pol);
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
    // A thread's columns increase over its iterations and tail columns are larger than vector
    // columns; `upd` keeps the earlier (lower) column on ties, `combine` the lower index.
    if (packed) {
        const unsigned long long k1 = s.j1 >= 0 ? pack_key(s.b1, (unsigned)s.j1) : ~0ull;
        const unsigned long long k2 = s.b2 != LLONG_MAX ? pack_key(s.b2, (unsigned)kPersonMask) : ~0ull;
        return packed_reduce(k1, k2, red_b1, red_b2);
    }
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

// Cheapest slot of object o (lowest slot index on ties) when slot `t_new` holds `p_new` and
// every other slot its stored price.
__device__ __forceinline__ void cheapest_slot(const LapParams &P, const int *__restrict__ ssoff, int o, int t_new,
                                              long long p_new, int &ms, long long &mp) {
    const int s0 = P.soff ? (ssoff ? ssoff[o] : __ldg(P.soff + o)) : o;
    const int s1 = P.soff ? (ssoff ? ssoff[o + 1] : __ldg(P.soff + o + 1)) : o + 1;
    ms = s0; mp = (s0 == t_new) ? p_new : __ldcg(P.slot_price + s0);
#pragma unroll 4
    for (int t = s0 + 1; t < s1; ++t) {
        const long long p = (t == t_new) ? p_new : __ldcg(P.slot_price + t);
        if (p < mp) { mp = p; ms = t; }
    }
}

// Sweeper side: rebuild person i's candidate list from one CTA-wide pass over its row.  Each
// thread keeps the best / second-best value of ITS columns; bound = the smallest per-thread
// second-best; the list = the per-thread bests below bound.  Any other object is either some
// thread's non-best (value >= that thread's second-best >= bound) or a best >= bound.  Prices may be
// stale (lower than current): the bound is then merely weaker, never wrong.
template <bool SMEMP>
__device__ __forceinline__ void build_list(const LapParams &P, int i, const int32_t *__restrict__ r, int n, int cmin,
                                           int S, const long long *__restrict__ price, bool vec_ok,
                                           long long *red_b2, int *wcnt) {
    Best s{LLONG_MAX, LLONG_MAX, -1};
    const int t = threadIdx.x;
    const unsigned long long pol = l2_policy_evict_first();
    int jtail = 0;
    if (vec_ok) {
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
        const int n4 = n >> 2;
#pragma unroll 4
        for (int q = t; q < n4; q += kThreads) {
            const int4 c = ld_stream(r4 + q, pol);
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
    long long m = s.b2;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((t & 31) == 0) red_b2[t >> 5] = m;
    __syncthreads();
    m = red_b2[t & 31];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
    const long long bound = m;
    const bool q = s.j1 >= 0 && s.b1 < bound;
    int total;
    const int pos = block_excl_count(q, wcnt, total);              // two __syncthreads inside
    const unsigned gen = ((unsigned)((unsigned long long)__ldcg(&P.lst_hdr[i].y) >> 32) + 1u) & kGenMask;
    if (total <= kListK) {
        if (q) P.lst_ent[(long long)i * kListK + pos] = make_int2(s.j1 | (int)(gen << kPersonBits), __ldg(r + s.j1));
        __threadfence();
    }
    __syncthreads();
    if (t == 0) {
        longlong2 h;
        h.x = bound;
        h.y = (long long)(((unsigned long long)gen << 32) | (unsigned long long)(total <= kListK ? total : 0));
        P.lst_hdr[i] = h;
    }
}

// Leader side (one warp).  A list is FETCHED (header + first 64 entries: one L2 round trip, issued
// as early as the next bidder is known so that it overlaps the current bid's bookkeeping) and later
// EVALUATED against the current prices.
struct ListRegs {
    longlong2 h;
    int2 v0, v1;
};

__device__ __forceinline__ ListRegs list_fetch(const LapParams &P, int i) {
    const int lane = threadIdx.x & 31;
    const int2 *e = P.lst_ent + (long long)i * kListK;
    ListRegs r;
    r.v0 = __ldcg(e + lane);
    r.v1 = __ldcg(e + lane + 32);
    r.h = __ldcg(&P.lst_hdr[i]);
    return r;
}

// Returns true with the exact (b1, j1, b2) of a full scan in every lane when the result is certified,
// false otherwise (no list, torn list, or the second-best candidate is not below the bound).
template <bool SMEMP>
__device__ __forceinline__ bool list_eval(const LapParams &P, int i, const ListRegs &r, int cmin, int S,
                                          const long long *__restrict__ price, Best &out) {
    const int lane = threadIdx.x & 31;
    const longlong2 h = r.h;
    const int n = (int)((unsigned long long)h.y & 0xFFFFFFFFull);
    const unsigned gen = (unsigned)((unsigned long long)h.y >> 32);
    int2 v[kListK / 32];
    v[0] = r.v0; v[1] = r.v1;
    const int2 *e = P.lst_ent + (long long)i * kListK;
#pragma unroll
    for (int k = 2; k < kListK / 32; ++k) v[k] = (n > 64) ? __ldcg(e + lane + 32 * k) : make_int2(0, 0);   // rare second trip
    if (n < 2) return false;
    bool ok = true;
    // Values are compared as 64-bit keys (value << 18 | object): lexicographic (value, object) order,
    // valid while every value is < 2^46 (guaranteed when the scaled cost range is < 2^45).
    unsigned long long k1 = ~0ull, k2 = ~0ull;
#pragma unroll
    for (int k = 0; k < kListK / 32; ++k) {
        const int q = lane + 32 * k;
        if (q < n) {
            const int o = v[k].x & (int)kPersonMask;
            ok = ok && ((unsigned)v[k].x >> kPersonBits) == gen && o < P.O;
            const long long p = SMEMP ? price[ok ? o : 0] : __ldcg(price + (ok ? o : 0));
            const long long hv = (long long)(v[k].y - cmin) * S + p;
            ok = ok && hv >= 0 && hv < (1ll << 46);
            const unsigned long long key = ((unsigned long long)hv << kPersonBits) | (unsigned)o;
            if (key < k1) { k2 = k1; k1 = key; }
            else if (key < k2) k2 = key;
        }
    }
    if (!__all_sync(0xffffffffu, ok)) return false;
    const unsigned long long w1 = warp_min64(k1);
    const unsigned long long w2 = warp_min64(k1 == w1 ? k2 : k1);      // keys are unique (distinct objects)
    if (w2 == ~0ull) return false;
    Best s;
    s.b1 = (long long)(w1 >> kPersonBits); s.j1 = (int)(w1 & kPersonMask);
    s.b2 = (long long)(w2 >> kPersonBits);
    if (!(s.b2 < h.x)) return false;
    out = s;
    return true;
}

// TM = tail mode (0 Gauss-Seidel FIFO, 1 in-CTA Jacobi rounds): a compile-time choice so that each
// instantiation carries one tail's code and register pressure only.
template <bool SMEMP, int TM>
__global__ void __launch_bounds__(kThreads, 1) lap_auction_kernel(const LapParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int np = P.P, no = P.O;
    long long *sprice = reinterpret_cast<long long *>(smem_raw);
    unsigned *sp32 = reinterpret_cast<unsigned *>(smem_raw);          // !SMEMP && P.approx: price >> 15 per object
    size_t off = SMEMP ? ((size_t)no * 8 + 15) / 16 * 16 : (P.approx ? ((size_t)no * 4 + 15) / 16 * 16 : 0);
    int *myq = reinterpret_cast<int *>(smem_raw + off);
    int *sowner = myq + ((P.qcap + 3) & ~3);          // [P] replica of slot_owner, only when P.smem_owner
    int *sminslot = sowner + ((np + 3) & ~3);         // [O] replica of minslot, only when P.smem_owner && P.soff
    // (a shared-memory copy of the slot offsets was measured and rejected: at 30k x 5k it pushes the carve-out
    // from 196 to 228 KB, the L1 from 60 to 28 KB, and the record replay from 7.4 to 9.0 us)
    const int *ssoff = nullptr;
    __shared__ long long red_b1[32], red_b2[32];
    __shared__ int red_j[32], wcnt[32];
    __shared__ int tq[kTailMax], tq_head, tq_cnt, tq_status, sw_stop;
    __shared__ int lj_obj[32], lj_need[32], lj_cur[32], ch_person[2][32], lj_stat[2], lj_stop;
    __shared__ long long lj_bid[32];

    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    const int S = np + 1;                            // < 2^18: cost * S is a 32 x 32 -> 64 bit multiply-add
    const bool vec_ok = ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.cost) & 15) == 0);
    unsigned int bar_target = 0;
    const long long *price_rd = SMEMP ? sprice : P.lambda;

    auto rowptr = [&](int i) -> const int32_t * { return P.cost + (long long)i * P.ld; };
    auto capacity = [&](int o) -> int { return P.soff ? __ldg(P.soff + o + 1) - __ldg(P.soff + o) : 1; };

    // ---- pass 0: state init and the cost range ------------------------------
    {
        int lmin = INT_MAX, lmax = INT_MIN;
        for (int i = b; i < np; i += G) {
            const int32_t *r = rowptr(i);
            for (int j = t; j < no; j += kThreads) {
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
        for (int i = b * kThreads + t; i < np; i += G * kThreads) {
            P.slot_price[i] = 0; P.slot_owner[i] = -1; P.person_obj[i] = -1; P.person_slot[i] = -1;
            P.lst_hdr[i] = make_longlong2(LLONG_MIN, 0);       // no list yet
        }
        for (int o = b * kThreads + t; o < no; o += G * kThreads) {
            P.lambda[o] = capacity(o) > 0 ? 0 : kInf;          // a spot that takes no cell is priced out
            P.minslot[o] = P.soff ? __ldg(P.soff + o) : o;
            P.bidw[0][o] = 0ull; P.bidw[1][o] = 0ull; P.bidw[2][o] = 0ull;
        }
        if (SMEMP) for (int o = t; o < no; o += kThreads) sprice[o] = capacity(o) > 0 ? 0 : kInf;
        if (!SMEMP && P.approx) for (int o = t; o < no; o += kThreads) sp32[o] = price_hi(capacity(o) > 0 ? 0 : kInf);
        if (P.smem_owner) {
            for (int k = t; k < np; k += kThreads) sowner[k] = -1;
            if (P.soff) for (int o = t; o < no; o += kThreads) sminslot[o] = __ldg(P.soff + o);
        }
    }
    grid_barrier(P.bar, bar_target, G);
    const int cmin = __ldcg(P.gmm + 0), cmax = __ldcg(P.gmm + 1);
    long long eps = ((long long)cmax - (long long)cmin) * S / P.eps0_div;
    if (eps < 1) eps = 1;

    // control counters stay in (32-bit) registers; pure statistics live in shared memory (thread 0 of the CTA
    // updates them) so that they cost no registers in the scan loops
    int rounds = 0, phases = 0, tails = 0, tail_bids = 0;
    __shared__ long long st_acc[8];   // bids, max bidders, list hits, small rounds, ns bid / barrier / replay / tail
    if (t < 8) st_acc[t] = 0;
    __syncthreads();
    int status = 0;
    int cur = 0;             // buffer of the current round; the previous round used (cur + 2) % 3
    int prevF = 0;           // records of the previous round (their bid words are cleared in this resolve)
    const int tail_t = min(P.tail_t, TM == 1 ? 32 : kTailMax);
    // candidate-list keys pack (value << 18 | object): needs every value < 2^46, i.e. scaled costs < 2^45
    // (prices are bounded by kBidLimit = 2^45 already)
    const bool packed = ((long long)cmax - (long long)cmin + 1) * S < (1ll << 45) && P.packed_reduce;
    const bool approx_ok = !SMEMP && P.approx && packed;
    // one person's row against the current prices: (min, argmin, second min), valid in thread 0
    auto scan = [&](const int32_t *r) -> Best {
        if (!SMEMP && approx_ok) {
            Best o;
            if (scan_row_approx(r, no, cmin, S, sp32, P.lambda, vec_ok, red_b1, red_b2, red_j, o)) return o;
        }
        return scan_row<SMEMP>(r, no, cmin, S, price_rd, vec_ok, red_b1, red_b2, red_j, packed);
    };
    const bool use_lists = P.use_lists && G > 1 && ((long long)cmax - (long long)cmin + 1) * S < (1ll << 45);

    for (;;) {
        ++phases;
        // ---- phase start: which pairs survive eps-CS at the new eps? --------
        if (phases > 1) {
            for (int i = b; i < np; i += G) {
                const int o = __ldcg(P.person_obj + i);
                int f = 1;
                if (i + G < np && P.prefetch) prefetch_row_l2(rowptr(i + G), no);
                if (o >= 0) {
                    const int32_t *r = rowptr(i);
                    const Best s = scan(r);
                    if (t == 0) {
                        const long long alt = (s.j1 == o) ? s.b2 : s.b1;
                        const long long base = (long long)(__ldg(r + o) - cmin) * S;
                        const long long lam = SMEMP ? sprice[o] : __ldcg(P.lambda + o);
                        const int ps = __ldcg(P.person_slot + i);
                        f = 0;
                        if (alt < kInf / 2) {
                            if (base + lam > alt + eps) f = ps + 2;                       // drop: vacate slot ps
                            else if (base + __ldcg(P.slot_price + ps) > alt + eps)
                                P.slot_price[ps] = alt + eps - base;                      // clamp (>= lambda[o])
                        }
                    }
                }
                if (t == 0) P.flag[i] = f;
            }
            grid_barrier(P.bar, bar_target, G);
        }
        int F = 0;
        for (int i0 = 0; i0 < np; i0 += kThreads) {
            const int i = i0 + t;
            const int f = (i < np) ? (phases == 1 ? 1 : __ldcg(P.flag + i)) : 0;
            if (f >= 2) {
                if (i % G == b || !P.smem_owner) P.slot_owner[f - 2] = -1;      // the slot keeps its price
                if (P.smem_owner) sowner[f - 2] = -1;
                if (i % G == b) { P.person_obj[i] = -1; P.person_slot[i] = -1; }
            }
            int tot;
            const int pos = F + block_excl_count(f != 0, wcnt, tot);
            if (f != 0 && pos % G == b) { P.list[cur][pos] = i; myq[pos / G] = i; }
            F += tot;
        }
        if (phases > 1 && P.soff) {
            // clamps may have created an equally cheap slot with a lower index: refresh the argmin
            for (int o = t; o < no; o += kThreads) {
                if (capacity(o) > 1) {
                    int ms; long long mp;
                    cheapest_slot(P, ssoff, o, -1, 0, ms, mp);
                    if (o % G == b || !P.smem_owner) P.minslot[o] = ms;
                    if (P.smem_owner) sminslot[o] = ms;
                }
            }
        }
        __syncthreads();

        // ---- bidding rounds --------------------------------------------------
        bool ran_tail = false;
        while (F > 0) {
            // Incomplete phases (Bertsekas): eps-CS holds for every assigned pair whether or not the
            // phase assigns everybody, so a phase with eps > 1 stops as soon as only a handful of
            // persons are still fighting a price war; they re-enter the next phase's free list
            // (person_obj == -1).  Only the eps == 1 phase has to run to completion.
            if (eps > 1 && F <= P.early_stop) { F = 0; break; }
            if (F <= tail_t) {
                // ---- Gauss-Seidel tail: few bidders left, a grid barrier per round would cost more
                // than the bids.  CTA 0 alone drains a FIFO of free persons; every bid sees the
                // prices the previous one left and the lone bidder always wins, so nothing is
                // exchanged until the phase ends.  (Same auction, sequential order: still exact.)
                grid_barrier(P.bar, bar_target, G);          // every CTA has finished its resolve writes
                ran_tail = true; ++tails;
                const long long tt0 = (b == 0 && t == 0) ? global_ns() : 0;
                if (b == 0) {
                    if (t < F) tq[t] = __ldcg(P.list[cur] + t);
                    if (t == 0) { tq_head = 0; tq_cnt = F; tq_status = 0; }
                    __syncthreads();
                    // ---- tail mode 1: the SAME synchronous (Jacobi) rounds as the grid runs, but inside
                    // CTA 0: warp w serves bidder w from its candidate list (a full-row scan by the whole CTA
                    // when the list cannot certify the result), warp 0 resolves the <= 32 bids with
                    // match.any / redux and commits the winners.  No grid barrier, and because the round
                    // semantics are unchanged the assignment does not depend on where the switch happens.
                    // one thread: person i takes the cheapest slot of object o at price `bid`; returns the
                    // person it evicts (-1: the slot was free)
                    // who holds the slot a winning bid for object o takes (known before any slot price is read)
                    auto peek_prev = [&](int o, int &slot) -> int {
                        slot = P.soff ? (P.smem_owner ? sminslot[o] : __ldcg(P.minslot + o)) : o;
                        return P.smem_owner ? sowner[slot] : __ldcg(P.slot_owner + slot);
                    };
                    auto assign_one = [&](int i, int o, long long bid, int slot, int prev) -> int {
                        if (bid >= kBidLimit) tq_status = CYB_ERR_OVERFLOW;
                        long long mp = bid; int ms = slot;
                        if (P.soff) cheapest_slot(P, ssoff, o, slot, bid, ms, mp);
                        if (P.smem_owner) { sowner[slot] = i; if (P.soff) sminslot[o] = ms; }
                        P.slot_owner[slot] = i; P.slot_price[slot] = bid;
                        P.person_obj[i] = o; P.person_slot[i] = slot;
                        if (P.soff) P.minslot[o] = ms;
                        P.lambda[o] = mp;
                        if (SMEMP) sprice[o] = mp;
                        else if (P.approx) sp32[o] = price_hi(mp);
                        if (prev >= 0) { P.person_obj[prev] = -1; P.person_slot[prev] = -1; }
                        return prev;
                    };
                    if (TM == 1) {
                        // Every warp owns one chain: it keeps bidding for the person it currently holds
                        // (first a free person of the list, afterwards whoever its last win evicted), so
                        // the next candidate list is fetched by the warp that needs it the moment the
                        // eviction is known, across the barrier that publishes the round's commits.
                        const int w = t >> 5, lane = t & 31;
                        if (t < 32) { ch_person[0][t] = t < F ? tq[t] : -1; ch_person[1][t] = -1; }
                        if (t == 0) { lj_stat[0] = 0; lj_stat[1] = 0; lj_stop = 0; }
                        __syncthreads();
                        int par = 0;
                        int me = ch_person[0][w];
                        ListRegs lr;
                        bool have_list = false;
                        for (;;) {
                            const int pl = ch_person[par][lane];
                            const unsigned amask = __ballot_sync(0xffffffffu, pl >= 0);
                            if (amask == 0u || lj_stop != 0) break;       // lj_stop only changes between #1 and #2
                            const bool alone = (amask & (amask - 1u)) == 0u;
                            bool ok = false;
                            Best s{0, 0, 0};
                            if (me >= 0) {
                                for (;;) {
                                    if (use_lists) {
                                        if (!have_list) { lr = list_fetch(P, me); have_list = true; }
                                        ok = list_eval<SMEMP>(P, me, lr, cmin, S, price_rd, s);
                                    }
                                    if (!(ok && alone)) break;
                                    // a single chain in flight: this warp follows it alone (a round with one
                                    // bidder needs no resolution and no barrier)
                                    const long long lam = SMEMP ? sprice[s.j1] : __ldcg(P.lambda + s.j1);
                                    const long long bid = lam + (s.b2 < kInf / 2 ? s.b2 - s.b1 : 0) + eps;
                                    const int old = me;
                                    int prev = 0, slot = 0;
                                    if (lane == 0) prev = peek_prev(s.j1, slot);
                                    prev = __shfl_sync(0xffffffffu, prev, 0);
                                    me = prev; have_list = false; ok = false;
                                    if (me >= 0 && use_lists) { lr = list_fetch(P, me); have_list = true; }   // before the commit's loads
                                    if (lane == 0) {
                                        assign_one(old, s.j1, bid, slot, prev);
                                        const int nb = atomicAdd(&lj_stat[0], 1) + 1; atomicAdd(&lj_stat[1], 1);
                                        if (nb + rounds > P.max_rounds) tq_status = CYB_ERR_NOT_CONVERGED;
                                    }
                                    __syncwarp();
                                    if (me < 0 || bid >= kBidLimit || tq_status != 0) break;
                                }
                                if (lane == 0) {
                                    lj_need[w] = (me >= 0 && !ok) ? 1 : 0;
                                    if (me >= 0 && ok) {
                                        const long long lam = SMEMP ? sprice[s.j1] : __ldcg(P.lambda + s.j1);
                                        lj_obj[w] = s.j1;
                                        lj_bid[w] = lam + (s.b2 < kInf / 2 ? s.b2 - s.b1 : 0) + eps;
                                    }
                                    lj_cur[w] = me;                       // person bidding for this chain in this round
                                }
                            } else if (lane == 0) {
                                lj_need[w] = 0; lj_cur[w] = -1;
                            }
                            __syncthreads();                                  // #1: every bid (or scan request) is posted
                            const int cl = lj_cur[lane];
                            const unsigned miss = __ballot_sync(0xffffffffu, cl >= 0 && lj_need[lane] != 0);
                            if (miss) {
                                for (int k = 0; k < 32; ++k) {
                                    if (!((miss >> k) & 1u)) continue;
                                    const Best r = scan(rowptr(lj_cur[k]));
                                    if (t == 0) {
                                        const long long lam = SMEMP ? sprice[r.j1] : __ldcg(P.lambda + r.j1);
                                        lj_obj[k] = r.j1;
                                        lj_bid[k] = lam + (r.b2 < kInf / 2 ? r.b2 - r.b1 : 0) + eps;
                                    }
                                }
                                __syncthreads();
                            }
                            int next = me;
                            if (me >= 0) {
                                const int o = lj_obj[w];
                                const long long bid = lj_bid[w];
                                // highest bid on the object wins, lowest person on equal bids
                                const bool beats = cl >= 0 && lane != w && lj_obj[lane] == o &&
                                                   (lj_bid[lane] > bid || (lj_bid[lane] == bid && cl < me));
                                const bool win = !__any_sync(0xffffffffu, beats);
                                if (win) {
                                    int prev = 0, slot = 0;
                                    if (lane == 0) prev = peek_prev(o, slot);
                                    next = __shfl_sync(0xffffffffu, prev, 0);
                                    have_list = false;
                                    if (next >= 0 && use_lists) { lr = list_fetch(P, next); have_list = true; }   // in flight across #2
                                    if (lane == 0) assign_one(me, o, bid, slot, prev);
                                }
                                if (lane == 0) {
                                    const int nb = atomicAdd(&lj_stat[0], 1) + 1;
                                    if (!((miss >> w) & 1u)) atomicAdd(&lj_stat[1], 1);
                                    if (nb + rounds > P.max_rounds) tq_status = CYB_ERR_NOT_CONVERGED;
                                }
                            }
                            if (lane == 0) ch_person[par ^ 1][w] = next;
                            if (t == 0) lj_stop = tq_status;
                            me = next;
                            par ^= 1;
                            __syncthreads();                                  // #2: commits and next persons visible
                        }
                        tail_bids += lj_stat[0];
                        if (t == 0) st_acc[2] += lj_stat[1];
                        if (t == 0) tq_cnt = 0;
                    }
                    while (TM == 0 && tq_cnt > 0 && tq_status == 0) {
                        // one bid: thread 0 books the result `s` of person i's scan (list or full row)
                        // commit one bid (one thread): person i takes `slot` of object o at price `bid`,
                        // evicting `prev`; (ms, mp) = the object's cheapest slot / price afterwards
                        auto commit = [&](int i, int o, long long bid, int slot, int prev, int ms, long long mp) {
                            if (bid >= kBidLimit) tq_status = CYB_ERR_OVERFLOW;
                            if (tail_bids + rounds > P.max_rounds) tq_status = CYB_ERR_NOT_CONVERGED;
                            if (P.smem_owner) { sowner[slot] = i; if (P.soff) sminslot[o] = ms; }
                            P.slot_owner[slot] = i; P.slot_price[slot] = bid;
                            P.person_obj[i] = o; P.person_slot[i] = slot;
                            if (P.soff) P.minslot[o] = ms;
                            P.lambda[o] = mp;
                            if (SMEMP) sprice[o] = mp;
                            else if (P.approx) sp32[o] = price_hi(mp);
                            int head = tq_head + 1; if (head == kTailMax) head = 0;
                            int cnt = tq_cnt - 1;
                            if (prev >= 0) {
                                P.person_obj[prev] = -1; P.person_slot[prev] = -1;
                                int tail = head + cnt; if (tail >= kTailMax) tail -= kTailMax;
                                tq[tail] = prev; ++cnt;
                            }
                            tq_head = head; tq_cnt = cnt;
                        };
                        // thread-0 version (after a full row scan)
                        auto book = [&](int i, const Best &s) {
                            const int o = s.j1;
                            const long long lam = SMEMP ? sprice[o] : __ldcg(P.lambda + o);
                            const long long bid = lam + (s.b2 < kInf / 2 ? s.b2 - s.b1 : 0) + eps;
                            const int slot = P.soff ? (P.smem_owner ? sminslot[o] : __ldcg(P.minslot + o)) : o;
                            const int prev = P.smem_owner ? sowner[slot] : __ldcg(P.slot_owner + slot);
                            long long mp = bid; int ms = slot;
                            if (P.soff) cheapest_slot(P, ssoff, o, slot, bid, ms, mp);
                            commit(i, o, bid, slot, prev, ms, mp);
                        };
                        // warp version (after a list hit, every lane holds `s`): the slots of a capacitated
                        // object are inspected by the lanes in parallel -- one L2 round trip instead of four
                        auto book_warp = [&](int i, const Best &s) -> int {
                            const int o = s.j1, lane = t & 31;
                            const long long lam = SMEMP ? sprice[o] : __ldcg(P.lambda + o);
                            const long long bid = lam + (s.b2 < kInf / 2 ? s.b2 - s.b1 : 0) + eps;
                            if (!P.soff) {
                                const int prev = P.smem_owner ? sowner[o] : __ldcg(P.slot_owner + o);   // uniform address
                                __syncwarp();
                                if (lane == 0) commit(i, o, bid, o, prev, o, bid);
                                return prev;
                            }
                            const int s0 = ssoff ? ssoff[o] : __ldg(P.soff + o), s1 = ssoff ? ssoff[o + 1] : __ldg(P.soff + o + 1);
                            if (s1 - s0 > 32) {
                                const int slot = __ldcg(P.minslot + o);
                                const int prev = __ldcg(P.slot_owner + slot);
                                __syncwarp();
                                if (lane == 0) book(i, s);
                                return prev;
                            }
                            const int tl = s0 + lane;
                            long long p = 0; int w = -1;
                            if (tl < s1) { p = __ldcg(P.slot_price + tl); w = __ldcg(P.slot_owner + tl); }
                            // cheapest slot, lowest index on ties (slot prices < 2^45)
                            const unsigned long long k1 = warp_min64(tl < s1 ? (((unsigned long long)p << 5) | lane) : ~0ull);
                            const int tlane = (int)(k1 & 31);
                            const int prev = __shfl_sync(0xffffffffu, w, tlane);
                            const unsigned long long k2 =
                                warp_min64(tl < s1 ? (((unsigned long long)(lane == tlane ? bid : p) << 5) | lane) : ~0ull);
                            if (lane == 0) commit(i, o, bid, s0 + tlane, prev, s0 + (int)(k2 & 31), (long long)(k2 >> 5));
                            return prev;
                        };
                        if (use_lists) {
                            // warp 0 alone serves bids from the candidate lists the sweeper CTAs keep
                            // fresh, until one cannot be certified; the other warps wait at the barrier
                            if (t < 32) {
                                int i = tq[tq_head];
                                ListRegs lr = list_fetch(P, i);
                                while (tq_cnt > 0 && tq_status == 0) {
                                    // the next bidder is the second FIFO entry when there is one, otherwise the
                                    // person this bid evicts: fetch its list as early as it is known
                                    int nxt = -1;
                                    ListRegs ln;
                                    if (tq_cnt > 1) {
                                        int h2 = tq_head + 1; if (h2 == kTailMax) h2 = 0;
                                        nxt = tq[h2];
                                        ln = list_fetch(P, nxt);
                                    }
                                    Best s;
                                    if (!list_eval<SMEMP>(P, i, lr, cmin, S, price_rd, s)) break;
                                    if (t == 0) ++st_acc[2];
                                    ++tail_bids;
                                    if (nxt < 0 && P.smem_owner && P.soff) {
                                        // capacitated object: who gets evicted is known from the shared-memory
                                        // replicas alone, so the next list is requested BEFORE the slot prices
                                        // of the object are read from L2 (one round trip per step, not two)
                                        const int pe = sowner[sminslot[s.j1]];
                                        if (pe >= 0) { nxt = pe; ln = list_fetch(P, nxt); }
                                    }
                                    const int prev = book_warp(i, s);
                                    __syncwarp();
                                    if (nxt < 0) {
                                        if (prev < 0) break;                  // queue is empty now
                                        nxt = prev;
                                        ln = list_fetch(P, nxt);
                                    }
                                    i = nxt; lr = ln;
                                }
                            }
                            __syncthreads();
                            if (!(tq_cnt > 0 && tq_status == 0)) break;
                        }
                        const int i = tq[tq_head];
                        const Best s = scan(rowptr(i));
                        ++tail_bids;
                        if (t == 0) book(i, s);
                        __syncthreads();
                    }
                    if (t == 0) {
                        st_acc[7] += global_ns() - tt0;
                        if (tq_status) atomicExch(P.gmm + 2, tq_status);
                        __threadfence();
                        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(P.gmm + 3), "r"((int)tails) : "memory");
                    }
                } else if (use_lists && b <= P.sweepers) {
                    // ---- sweepers: while CTA 0 runs the tail, the idle CTAs keep re-deriving candidate
                    // lists (person i belongs to CTA 1 + i mod (G-1)) against the prices CTA 0 publishes.
                    // Purely speculative: a list only ever short-cuts a scan whose result it reproduces
                    // exactly, so neither timing nor staleness can change the assignment.
                    bool stop = false;
                    while (!stop) {
                        if (SMEMP) {
                            for (int o = t; o < no; o += kThreads) sprice[o] = __ldcg(P.lambda + o);
                            __syncthreads();
                        }
                        for (int i = b - 1; i < np; i += P.sweepers) {
                            if (t == 0) {
                                int v;
                                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(P.gmm + 3) : "memory");
                                sw_stop = (v >= (int)tails);
                            }
                            __syncthreads();
                            if (sw_stop) { stop = true; break; }
                            build_list<SMEMP>(P, i, rowptr(i), no, cmin, S, price_rd, vec_ok, red_b2, wcnt);
                        }
                    }
                }
                F = 0;
                break;
            }
            if (++rounds > P.max_rounds) { status = CYB_ERR_NOT_CONVERGED; break; }
            if (t == 0) { st_acc[0] += F; if (F > st_acc[1]) st_acc[1] = F; }
            const bool timed = (b == 0 && t == 0 && F <= G);
            const long long tm0 = timed ? global_ns() : 0;
            const int myn = F > b ? (F - b - 1) / G + 1 : 0;
            if (myn > 0 && P.prefetch) prefetch_row_l2(rowptr(myq[0]), no);
            for (int q = 0; q < myn; ++q) {
                const int i = myq[q];
                if (q + 1 < myn && P.prefetch) prefetch_row_l2(rowptr(myq[q + 1]), no);     // next row rides under this scan
                const Best s = scan(rowptr(i));
                if (t == 0) {
                    const int o = s.j1;
                    const long long lam = SMEMP ? sprice[o] : __ldcg(P.lambda + o);
                    const long long bid = lam + (s.b2 < kInf / 2 ? s.b2 - s.b1 : 0) + eps;
                    if (bid >= kBidLimit) atomicExch(P.gmm + 2, CYB_ERR_OVERFLOW);
                    const int slot = P.soff ? (P.smem_owner ? sminslot[o] : __ldcg(P.minslot + o)) : o;
                    const int prev = P.smem_owner ? sowner[slot] : __ldcg(P.slot_owner + slot);
                    P.rec[cur][q * G + b] = make_int4(o, slot, prev, 0);
                    atomicMax(P.bidw[cur] + o,
                              ((unsigned long long)bid << kPersonBits) | (kPersonMask - (unsigned long long)i));
                }
            }
            const long long tm1 = timed ? global_ns() : 0;
            grid_barrier(P.bar, bar_target, G);
            const long long tm2 = timed ? global_ns() : 0;
            // ---- resolve: every CTA replays every record ----------------------
            // (the error flag is checked after the replay so that its load overlaps the record loads)
            status = __ldcg(P.gmm + 2);
            const int nxt = cur == 2 ? 0 : cur + 1, prv = cur == 0 ? 2 : cur - 1;
            // clear the bid words of the previous round (their next use is two rounds ahead)
            for (int k = b * kThreads + t; k < prevF; k += G * kThreads)
                P.bidw[prv][__ldcg(&P.rec[prv][k].x)] = 0ull;
            int Fn = 0;
            for (int k0 = 0; k0 < F; k0 += kThreads) {
                const int k = k0 + t;
                int entry = -1;
                if (k < F) {
                    const int i = __ldcg(P.list[cur] + k);
                    const int4 rc = __ldcg(P.rec[cur] + k);
                    const unsigned long long key = __ldcg(P.bidw[cur] + rc.x);
                    const int wperson = (int)(kPersonMask - (key & kPersonMask));
                    if (wperson == i) {
                        const long long bid = (long long)(key >> kPersonBits);
                        const bool mine = (k % G == b);
                        int ms; long long mp;
                        // sibling slot prices FIRST: with the stores in front, 148 CTAs storing to one line and
                        // then loading from it serialised at the L2 slice (replay 8.3 us -> 3.0 us at 30k x 5k)
                        cheapest_slot(P, ssoff, rc.x, rc.y, bid, ms, mp);
                        // global copies: every CTA writes them (identical values) only when some CTA reads them
                        // back before the next barrier, i.e. without the shared-memory replicas
                        // (slot prices are only read behind a barrier: one writer; the cheapest-slot index of a
                        // unit-capacity object never changes)
                        if (mine || !P.smem_owner) {
                            P.slot_owner[rc.y] = i;
                            if (P.soff) P.minslot[rc.x] = ms;
                        }
                        if (mine) P.slot_price[rc.y] = bid;
                        if (P.smem_owner) { sowner[rc.y] = i; if (P.soff) sminslot[rc.x] = ms; }
                        if (SMEMP) { sprice[rc.x] = mp; if (mine) P.lambda[rc.x] = mp; }
                        else { P.lambda[rc.x] = mp; if (P.approx) sp32[rc.x] = price_hi(mp); }
                        if (mine) {
                            P.person_obj[i] = rc.x; P.person_slot[i] = rc.y;
                            if (rc.z >= 0) { P.person_obj[rc.z] = -1; P.person_slot[rc.z] = -1; }
                        }
                        entry = rc.z;
                    } else {
                        entry = i;
                    }
                }
                int tot;
                const int pos = Fn + block_excl_count(entry >= 0, wcnt, tot);
                if (entry >= 0 && pos % G == b) { P.list[nxt][pos] = entry; myq[pos / G] = entry; }
                Fn += tot;
            }
            prevF = F;
            F = Fn;
            cur = nxt;
            __syncthreads();
            if (timed) { st_acc[4] += tm1 - tm0; st_acc[5] += tm2 - tm1; st_acc[6] += global_ns() - tm2; ++st_acc[3]; }
            if (status) break;
        }
        if (status) break;
        grid_barrier(P.bar, bar_target, G);        // state of the last resolve / the tail becomes visible
        {
            // the bid words of the last round back to zero before the next phase
            const int prv = cur == 0 ? 2 : cur - 1;
            for (int k = b * kThreads + t; k < prevF; k += G * kThreads)
                P.bidw[prv][__ldcg(&P.rec[prv][k].x)] = 0ull;
            prevF = 0;
        }
        if (ran_tail) {
            status = __ldcg(P.gmm + 2);
            if (status) break;
            if (eps > 1 && b != 0) {                // pick up what CTA 0 moved during the tail
                if (SMEMP) for (int o = t; o < no; o += kThreads) sprice[o] = __ldcg(P.lambda + o);
                else if (P.approx) for (int o = t; o < no; o += kThreads) sp32[o] = price_hi(__ldcg(P.lambda + o));
                if (P.smem_owner) {
                    for (int k = t; k < np; k += kThreads) sowner[k] = __ldcg(P.slot_owner + k);
                    if (P.soff) for (int o = t; o < no; o += kThreads) sminslot[o] = __ldcg(P.minslot + o);
                }
            }
            __syncthreads();
        }
        if (eps == 1) break;
        eps /= P.theta;
        if (eps < 1) eps = 1;
    }

    // ---- total cost of the assignment -------------------------------------------
    if (!status) {
        long long sum = 0;
        for (int i = b * kThreads + t; i < np; i += G * kThreads) {
            const int o = __ldcg(P.person_obj + i);
            if (o >= 0) sum += (long long)__ldg(rowptr(i) + o);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if ((t & 31) == 0 && sum != 0) atomicAdd(reinterpret_cast<unsigned long long *>(P.total), (unsigned long long)sum);
    }
    if (b == 0 && t == 0) {
        P.stats[0] = status; P.stats[1] = phases; P.stats[2] = rounds; P.stats[3] = st_acc[0];
        P.stats[4] = phases - 1; P.stats[5] = cmin; P.stats[6] = cmax; P.stats[7] = S;
        P.stats[8] = G; P.stats[9] = SMEMP ? 1 : 0; P.stats[10] = TM; P.stats[11] = st_acc[1];
        P.stats[12] = (phases - 1) * (long long)np; P.stats[13] = tail_bids; P.stats[14] = tails; P.stats[15] = st_acc[2];
        P.stats[16] = st_acc[3]; P.stats[17] = st_acc[4]; P.stats[18] = st_acc[5]; P.stats[19] = st_acc[6]; P.stats[20] = st_acc[7];
    }
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

__global__ void __launch_bounds__(kChkThreads) lap_rowcheck_whole_kernel(
    const int32_t *__restrict__ cost, long long ld, int np, int no, const int32_t *__restrict__ person_obj,
    const long long *__restrict__ price, long long S, int32_t *__restrict__ count, long long *__restrict__ acc,
    const int32_t *__restrict__ soff, long long *__restrict__ out, unsigned int *__restrict__ done) {
    extern __shared__ __align__(16) long long spw[];
    __shared__ long long part[2][kTeams][kTeam / 32];
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
    const int i0 = blockIdx.x * kTeams + team, istep = gridDim.x * kTeams;
    // the row's certificate terms ride along with the stream: obj(i) is fetched one row ahead and
    // cost[i, obj(i)] with the row itself, so the team leader adds no dependent round trip per row
    int o_cur = (tt == 0 && i0 < np) ? __ldg(person_obj + i0) : -1;
    for (int i = i0; i < np; i += istep, ++it) {
        const int32_t *r = cost + (long long)i * ld;
        const int o_next = (tt == 0 && i + istep < np) ? __ldg(person_obj + i + istep) : -1;
        const bool o_ok = o_cur >= 0 && o_cur < no;
        const int c_o = (tt == 0 && o_ok) ? __ldg(r + o_cur) : 0;
        long long m = LLONG_MAX;
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
#pragma unroll 8
        for (int q = tt; q < n4; q += kTeam) {
            const int4 c = ld_stream(r4 + q, pol);
            const longlong2 a = *reinterpret_cast<const longlong2 *>(spw + 4 * q);
            const longlong2 bb = *reinterpret_cast<const longlong2 *>(spw + 4 * q + 2);
            m = min(m, (long long)c.x * S + a.x);
            m = min(m, (long long)c.y * S + a.y);
            m = min(m, (long long)c.z * S + bb.x);
            m = min(m, (long long)c.w * S + bb.y);
        }
        for (int j = (n4 << 2) + tt; j < no; j += kTeam) m = min(m, (long long)__ldg(r + j) * S + spw[j]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) part[it & 1][team][wt] = m;
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
        o_cur = o_next;
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

struct WsLayout {
    size_t list[3], rec[3], bidw[3], flag, slot_price, person_slot, minslot, rowmin, count, lst_hdr, lst_ent, small, total;
};

WsLayout ws_layout(int64_t np, int64_t no) {
    WsLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = cyb::align_up(o + bytes, 256); return r; };
    for (int k = 0; k < 3; ++k) L.list[k] = take((size_t)np * 4);
    for (int k = 0; k < 3; ++k) L.rec[k] = take((size_t)np * 16);
    for (int k = 0; k < 3; ++k) L.bidw[k] = take((size_t)no * 8);
    L.flag = take((size_t)np * 4);
    L.slot_price = take((size_t)np * 8);
    L.person_slot = take((size_t)np * 4);
    L.minslot = take((size_t)no * 4);
    L.rowmin = take((size_t)np * 8);
    L.count = take(cyb::align_up((size_t)no * 4, 64) + 64);      // counts, then {4 x int64 accumulators, ticket}
    L.lst_hdr = take((size_t)np * 16);
    L.lst_ent = take((size_t)np * kListK * 8);
    L.small = take(256);
    L.total = o;
    return L;
}

}  // namespace

// (the exported cyb_lap_workspace_bytes / cyb_lap_solve_i32 live in lap_sap.cu; this is the round-1 solver,
// reachable with CYB_LAP_SOLVER=auction for A/B measurements)
size_t cyb::lap_auction_workspace_bytes(int64_t n_persons, int64_t n_objects) {
    if (n_persons <= 0 || n_objects <= 0) return 256;
    return ws_layout(n_persons, n_objects).total;
}

int cyb::lap_solve_auction(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                           const int32_t *slot_offset_dev, int32_t *person_obj_dev,
                           int32_t *slot_owner_dev, int64_t *price_dev, int64_t *total_dev,
                           int64_t *stats_dev, void *workspace_dev, size_t workspace_bytes,
                           int grid_hint, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int64_t np = n_persons, no = n_objects;
    if (np <= 0 || np >= (1ll << kPersonBits) || no <= 0 || no > np)
        return cyb::set_error(CYB_ERR_INVALID,
                              "cyb_lap_solve_i32: persons=%lld objects=%lld outside 1 <= objects <= persons < 2^18",
                              (long long)np, (long long)no);
    if (!slot_offset_dev && no != np)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: without capacities the problem must be square");
    if (!cost_dev || !person_obj_dev || !slot_owner_dev || !price_dev || !total_dev || !stats_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: null pointer argument");
    if (ld < no)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: ld=%lld < objects=%lld", (long long)ld, (long long)no);
    const WsLayout L = ws_layout(np, no);
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
    if (G > np) G = (int)np;
    if (G < 1) G = 1;

    char *ws = static_cast<char *>(workspace_dev);
    LapParams P;
    P.cost = cost_dev; P.ld = ld; P.P = (int)np; P.O = (int)no; P.soff = slot_offset_dev;
    P.person_obj = person_obj_dev; P.slot_owner = slot_owner_dev;
    P.lambda = reinterpret_cast<long long *>(price_dev);
    P.total = reinterpret_cast<long long *>(total_dev); P.stats = reinterpret_cast<long long *>(stats_dev);
    for (int k = 0; k < 3; ++k) {
        P.list[k] = reinterpret_cast<int32_t *>(ws + L.list[k]);
        P.rec[k] = reinterpret_cast<int4 *>(ws + L.rec[k]);
        P.bidw[k] = reinterpret_cast<unsigned long long *>(ws + L.bidw[k]);
    }
    P.flag = reinterpret_cast<int32_t *>(ws + L.flag);
    P.slot_price = reinterpret_cast<long long *>(ws + L.slot_price);
    P.person_slot = reinterpret_cast<int32_t *>(ws + L.person_slot);
    P.minslot = reinterpret_cast<int32_t *>(ws + L.minslot);
    P.bar = reinterpret_cast<unsigned int *>(ws + L.small);
    P.gmm = reinterpret_cast<int *>(ws + L.small + 16);
    P.qcap = (int)((np + G - 1) / G);
    P.max_rounds = 2000ll * np + 100000;
    P.lst_hdr = reinterpret_cast<longlong2 *>(ws + L.lst_hdr);
    P.lst_ent = reinterpret_cast<int2 *>(ws + L.lst_ent);
    P.tail_t = 8;
    // rows much longer than the 64 KB a CTA keeps in flight (measured: 50k 680 -> 620 ms; neutral at 40 KB rows,
    // within noise at 100 KB; prefetching for the sweepers as well slowed the tail they run beside)
    P.prefetch = no * 4 > 131072 ? 1 : 0;
    if (const char *e = getenv("CYB_LAP_PREFETCH")) P.prefetch = atoi(e) ? 1 : 0;
    P.packed_reduce = 1;
    if (const char *e = getenv("CYB_LAP_PACKED")) P.packed_reduce = atoi(e) ? 1 : 0;
    P.tail_mode = -1;        // chosen below, once the shared-memory residency is known
    P.early_stop = 0;      // measured: postponed price wars get longer at smaller eps
    P.use_lists = 1;
    P.theta = kTheta; P.eps0_div = kEps0Div;
    if (const char *e = getenv("CYB_LAP_THETA")) P.theta = std::max(2, atoi(e));
    if (const char *e = getenv("CYB_LAP_EPS0")) P.eps0_div = std::max(1, atoi(e));
    if (const char *e = getenv("CYB_LAP_TAIL_MODE")) P.tail_mode = atoi(e) ? 1 : 0;
    if (const char *e = getenv("CYB_LAP_EARLY")) P.early_stop = std::max(0, atoi(e));
    if (const char *e = getenv("CYB_LAP_LISTS")) P.use_lists = atoi(e);

    // small block: barrier counter, cmin = INT_MAX, cmax = INT_MIN, status = 0
    const int init[8] = {0, 0, 0, 0, INT_MAX, INT_MIN, 0, 0};      // gmm[3] = tails finished
    CYB_CUDA_CHECK(cudaMemcpyAsync(ws + L.small, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    CYB_CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), stream));

    const size_t q_bytes = cyb::align_up((size_t)P.qcap * 4, 16);
    const size_t smem_with_price = cyb::align_up((size_t)no * 8, 16) + q_bytes;
    const size_t static_smem = 2048;              // bound on the kernel's static shared memory (ptxas: 1552 B)
    bool smemp = smem_with_price + static_smem <= (size_t)max_smem;
    if (const char *e = getenv("CYB_LAP_SMEM_PRICES")) smemp = smemp && atoi(e);      // test hook: force the L2-price paths
    size_t dyn = smemp ? smem_with_price : q_bytes;
    // without room for the 8-byte prices, their 32-bit prefixes may still fit (50k objects: 200 KB)
    P.approx = 0;
    // (off by default: measured 50k 704 -> 696 ms, 32k 322 -> 379 ms -- the scan is bound by loads in flight,
    // not by the price stream, and the prefixes displace the slot-owner replica; CYB_LAP_APPROX=1 enables it)
    if (const char *e = getenv("CYB_LAP_APPROX"))
        if (atoi(e) && !smemp && cyb::align_up((size_t)no * 4, 16) + q_bytes + static_smem <= (size_t)max_smem) P.approx = 1;
    if (P.approx) dyn += cyb::align_up((size_t)no * 4, 16);
    const size_t owner_bytes = cyb::align_up((size_t)np * 4, 16) + (slot_offset_dev ? cyb::align_up((size_t)no * 4, 16) : 0);
    P.smem_owner = (dyn + owner_bytes + static_smem <= (size_t)max_smem) ? 1 : 0;
    if (const char *e = getenv("CYB_LAP_SMEM_OWNER")) P.smem_owner = P.smem_owner && atoi(e);
    if (P.smem_owner) dyn += owner_bytes;
    // Tail: with prices AND slot owners in shared memory a Gauss-Seidel step is one L2 round trip
    // (~0.95 us) and the FIFO tail wins or ties (10k, 30k x 5k); when either lives in L2 every step
    // chains two or three round trips and the multi-chain Jacobi tail hides them (measured on B200:
    // 25k 288 -> 244 ms, 50k 975 -> 712 ms).
    if (P.tail_mode < 0) P.tail_mode = (smemp && P.smem_owner) ? 0 : 1;
    P.tail_t = P.tail_mode == 1 ? 32 : 8;      // (6 for capacitated problems helped one 30k x 5k instance, hurt another)
    if (const char *e = getenv("CYB_LAP_TAIL")) P.tail_t = atoi(e);

    const void *fn = smemp ? (P.tail_mode ? (const void *)lap_auction_kernel<true, 1> : (const void *)lap_auction_kernel<true, 0>)
                           : (P.tail_mode ? (const void *)lap_auction_kernel<false, 1> : (const void *)lap_auction_kernel<false, 0>);
    CYB_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int occ = 0;
    CYB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, dyn));
    if (occ < 1) return cyb::set_error(CYB_ERR_UNSUPPORTED, "lap kernel does not fit on an SM (smem %zu)", dyn);
    if (G > occ * sms) G = occ * sms;
    P.sweepers = G - 1;
    if (const char *e = getenv("CYB_LAP_SWEEPERS")) P.sweepers = std::max(1, std::min(G - 1, atoi(e)));
    if (G < 2) P.use_lists = 0;
    void *args[] = {(void *)&P};
    CYB_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kThreads), args, dyn, stream));
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}

extern "C" int cyb_lap_check_i32(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                                 const int32_t *slot_offset_dev, const int32_t *person_obj_dev,
                                 const int64_t *price_dev, int64_t *out_dev, void *workspace_dev,
                                 size_t workspace_bytes, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int64_t np = n_persons, no = n_objects;
    if (np <= 0 || np >= (1ll << kPersonBits) || no <= 0 || no > np)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_check_i32: persons=%lld objects=%lld out of range",
                              (long long)np, (long long)no);
    if (!cost_dev || !person_obj_dev || !price_dev || !out_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_check_i32: null pointer argument");
    const WsLayout L = ws_layout(np, no);
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
            slot_offset_dev, reinterpret_cast<long long *>(out_dev), done);
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
