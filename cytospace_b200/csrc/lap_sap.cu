// Exact dense linear assignment (transportation form) on sm_100a -- the solver behind
// cyb_lap_solve_i32: eps-scaling auction rounds, phases with eps > 1 left INCOMPLETE (their last free persons
// are handed to the next phase), and a grid-wide multi-source SHORTEST-AUGMENTING-PATH finish for the last free
// persons of the eps = 1 phase, all in ONE persistent cooperative kernel (one CTA per SM).
//
// Replaces the third-party `lapjv.lapjv(cost)` call CytoSPACE makes at
// cytospace/linear_assignment_solvers/linear_assignment_solvers.py:38 (from
// cytospace/cytospace.py:329) on the expanded matrix `cost[location_repeat, :]`
// (linear_assignment_solvers.py:63-66).  The expansion is never materialised: PERSONS are the
// rows of the matrix, OBJECTS the columns, object o has cap[o] SLOTS:
//     min sum_i M[i, obj(i)]   s.t.  object o holds exactly cap[o] persons.
//
// Auction part (synchronous Jacobi rounds over the grid, one grid barrier per round: [bid] every CTA scans the
// rows of its share of the free list and posts a 64-bit atomicMax (bid | ~person) per object plus a record;
// [resolve] every CTA replays all records, so each CTA's view of the state is complete without a second barrier).
// C = (M - cmin) * (P+1) >= 0; every slot has a price and a holder, the object's price lambda[o] is
// its cheapest slot; a free person bids lambda[o*] + (w - v1) + eps for the cheapest slot of its
// best object, the highest bid per object wins.  Invariant (eps-CS): for every assigned (i, o)
//     C[i,o] + lambda[o] <= min_k (C[i,k] + lambda[k]) + eps,
// lambda never decreases, eps is divided by theta per phase down to 1; with the factor P+1 on the
// costs eps = 1 makes the integer total optimal (DESIGN.md has the proof).
//
// Why a different finish.  The last few free persons of a phase cost the auction thousands of
// DEPENDENT single bids (an augmenting path traced one eviction at a time: measured 45.7k dependent
// steps of ~0.9 us at 10k x 10k, 58-74 % of every solve).  A shortest-path search finds the same
// augmenting paths breadth-first: a few dozen grid-wide rounds, each relaxing a few hundred rows in
// parallel (bandwidth-shaped work), per search.  oracle/sap_model.c is the CPU model of exactly this
// algorithm (the experiment that sized it, and the reference the tests compare assignments with).
//
// Incomplete phases.  The next phase start drops every pair that violates the tighter eps and rebuilds the free list
// anyway, so placing the hardest persons of a phase with eps > 1 is wasted work: such a phase stops when <= `partial`
// persons are free (eps-CS holds for every assigned pair at all times; only the last phase must place everybody).
// 46 -> 8 searches and 1 498 -> 721 search rounds at 10k x 10k; with it a gentler schedule (theta = 8) pays.
//
// SAP finish (runs when <= sap_t persons are free; with the default partial >= sap_t only in the eps = 1 phase),
// residual graph at the current eps:
//   assigned person i (object o_i) -> object k != o_i : len = C[i,k]+lambda[k]+eps - (C[i,o_i]+lambda[o_i]) >= 0
//   free person i                  -> object k        : len = C[i,k]+lambda[k] - min_k'(C[i,k']+lambda[k'])   >= 0
//   object o -> every person it holds                 : len = 0
// One search relaxes labels d[o] from ALL free persons at once (label-correcting, any order is
// valid): round 0 = the free persons' rows; every further round takes the dirty objects (label
// lowered since their holders last relaxed) with the smallest labels -- about K rows, chosen by a
// 256-bin histogram threshold -- and relaxes their holders' rows with 64-bit atomicMin on
// (label << 18 | entering slot).  D = the want-th smallest label of an object with a free slot;
// the search ends when no dirty object has a label < D.  Then lambda[o] += D - d[o] for d[o] < D
// (prices only rise, every relaxed arc keeps len >= 0, tree arcs become tight) and the assignment
// is flipped along up to `want` vertex-disjoint tree paths (one warp traces each candidate, the
// paths claim their objects with atomicMin(rank); a path is applied iff it owns all its claims --
// the closest candidate always does).  A new pair (i, o) has C[i,o]+lambda[o]+eps = (old value of
// i): eps-CS is kept exactly, so auction rounds and searches mix freely inside a phase.
//
// Determinism.  A relaxation only counts if it is STRICTLY below the label the object had when the
// round started (snapshot g[k] = lambda[k] - d[k]: in shared memory, or in L2 kept by the slice
// owners when the prices do not fit), so the labels
// and the (label, slot) minima do not depend on the interleaving; the lists built with atomics are
// sets.  Same assignment for any grid size (tested against the CPU model).
//
// Shared memory per CTA (when it fits): 8 B per object -- the price lambda[o] during auction rounds,
// g[o] during a search -- plus 4 B per slot (holder replica) and 4 B per object (tree predecessor).
// HBM traffic: one row (O*4 bytes) per bid and per relaxed row; everything else is L2 / shared.
//
// Tuning knobs (environment, read at launch): CYB_LAP_SAP_T (free persons at which the search
// takes over), CYB_LAP_SAP_K (rows per search round), CYB_LAP_SAP_MULTI (paths per search, <= 32),
// CYB_LAP_WARM (0: every search starts from scratch), CYB_LAP_PARTIAL (0: every phase is completed), CYB_LAP_THETA,
// CYB_LAP_EPS0, CYB_LAP_TEAMS (rows a CTA scans side by side, <= 8), CYB_LAP_SMEM_PRICES=0 / CYB_LAP_SMEM_OWNER=0 (force
// the L2 paths).

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>

#include "common.h"

namespace cyb {
size_t lap_check_workspace_bytes(int64_t n_persons, int64_t n_objects);      // lap_check.cu
}  // namespace cyb

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kPB = 18;                                        // P < 2^18 - kSapMax
constexpr unsigned long long kPM = (1ull << kPB) - 1;
constexpr long long kInf = 1ll << 60;                          // price of an object without capacity / label "unreached"
constexpr long long kGInf = 1ll << 61;                         // g of a priced-out object: never relaxed
constexpr long long kBidLimit = 1ll << 45;                     // 46-bit bid / label field
constexpr int kSapMax = 256;                                   // free persons a search can start from
constexpr int kMultiMax = 32;                                  // augmenting paths per search (one warp each)
constexpr int kRowsMax = 32;                                   // rows a CTA relaxes per chunk
constexpr int kMaxSearch = 1 << 24;
constexpr int kMaxGrid = 160;                                  // CTAs (one per SM; B200: 148)

struct SapParams {
    const int32_t *cost;
    long long ld;
    int P, O;
    const int32_t *soff;     // slot offsets [O+1] (nullptr: every capacity is 1)
    int32_t *person_obj, *slot_owner;
    long long *lambda, *total, *stats;
    long long *slot_price;
    int32_t *person_slot, *minslot, *slot_obj;
    int32_t *list[3];
    int4 *rec[3];
    unsigned long long *bidw[3];
    int32_t *flag;
    unsigned int *bar;
    int *gmm;                // [0] cmin, [1] cmax, [2] status
    // search state
    unsigned long long *dkey;        // [O] (label << 18 | entering slot; sources are P + k), ~0 = unreached
    long long *gsnap;                // [O] lambda - (label at the start of the round), for the variants without shared-memory prices
    unsigned *chgbits[3];            // [(O+31)/32] objects whose label a round lowered, rotating by round
    int32_t *claim;                  // [O + kSapMax] path claims (decreasing base per search)
    int4 *moves;                     // [P] (person, object, slot, -) of the accepted paths
    int32_t *anc;                    // [G][O] per-CTA scratch of the kept / dropped pointer jumping (warm searches)
    int *nmoves;                     // [1]
    int32_t *srcdone;                // [kSapMax]
    unsigned long long *cand[3];     // [O] candidate lists (label << 18 | object), rotating by round
    int *rstat;                      // [3][8] per round {candidates, their slots, eligible slots, -, min eligible label (64 bit)}
    int wpc;                         // label words (32 objects) a CTA classifies per round
    int qcap;
    long long max_rounds;
    int sap_t, sap_k, multi, partial, max_teams;
    int warm;                        // 1: a search that follows another one in the phase starts from the surviving forest
    int theta, eps0_div;
    int packed_reduce, prefetch;
};

struct Best {
    long long b1, b2;
    int j1;
};

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
__device__ __forceinline__ void prefetch_row_l2(const int32_t *r, int n, int tid = threadIdx.x) {
    const unsigned bytes = ((unsigned)n * 4u) & ~15u;
    const unsigned chunk = 8192u;
    const unsigned off = (unsigned)tid * chunk;
    if (off < bytes && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
        const unsigned len = min(chunk, bytes - off);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char *>(r) + off), "r"(len) : "memory");
    }
}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void upd(Best &s, long long h, int j) {
    if (h < s.b2) {
        if (h < s.b1) { s.b2 = s.b1; s.b1 = h; s.j1 = j; }
        else s.b2 = h;
    }
}
__device__ __forceinline__ Best combine(const Best &a, const Best &b) {
    const bool bwins = (b.b1 < a.b1) || (b.b1 == a.b1 && (unsigned)b.j1 < (unsigned)a.j1);
    Best r;
    if (bwins) { r.b1 = b.b1; r.j1 = b.j1; r.b2 = a.b1 < b.b2 ? a.b1 : b.b2; }
    else       { r.b1 = a.b1; r.j1 = a.j1; r.b2 = b.b1 < a.b2 ? b.b1 : a.b2; }
    return r;
}
__device__ __noinline__ void report_timeout(int line, unsigned v, unsigned target) {
    printf("[cyb lap] CTA %d: wait at line %d timed out (counter %u, target %u)\n", blockIdx.x, line, v, target);
}
// Grid barrier with a watchdog.  `abortf` is a global flag: a CTA that waits longer than 5 s (only a
// protocol bug can do that) raises it, every other CTA leaves its wait when it sees it, and the kernel
// returns CYB_ERR_NOT_CONVERGED instead of hanging the GPU.  Returns true when the launch is aborted.
__device__ __forceinline__ bool spin_until(unsigned int *bar, unsigned int target, int *abortf, int line) {
    unsigned int v, spins = 0;
    long long t0 = 0;
    bool dead = false;
    for (;;) {
        // (relaxed polls + one acquire fence at the end measured slower: 50.6 -> 54 ms at cfg2 together with a 6-deep
        // first wave in the relax; both reverted)
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if ((int)(v - target) >= 0) break;
        if ((++spins & 0x3FFu) == 0) {
            int a;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(a) : "l"(abortf) : "memory");
            const long long now = global_ns();
            if (!t0) t0 = now;
            if (a != 0 || now - t0 > 5000000000ll) {
                if (a == 0) report_timeout(line, v, target);
                atomicExch(abortf, 1);
                dead = true;
                break;
            }
        }
    }
    return dead;
}
__device__ __forceinline__ bool grid_barrier(unsigned int *bar, unsigned int &target, unsigned int G, int *abortf, int line) {
    __syncthreads();
    bool dead = false;
    if (threadIdx.x == 0) {
        // release / acquire on the counter itself (no separate fences): the CTA's writes, ordered before this
        // thread by the bar.sync above, are published by the release; the acquire load in spin_until pairs with it
        target += G;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        dead = spin_until(bar, target, abortf, line);
    }
    return __syncthreads_or(dead);
}
// Block-wide exclusive prefix count of `valid`; `total` = number of valid threads.
__device__ __forceinline__ int block_excl_count(bool valid, int *wcnt, int &total) {
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int within = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[w] = __popc(m);
    __syncthreads();
    const int c = lane < kWarps ? wcnt[lane] : 0;
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
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    return ((unsigned long long)mhi << 32) | mlo;
}
// A row is scanned by a TEAM of ts threads (a power of two, 32 <= ts <= kThreads; consecutive warps): the whole CTA when it
// has one row to scan, down to 128 threads when it has eight or more -- each row scan is a chain of load latency, reduction
// and a leader epilogue, and only rows in flight side by side overlap those.  Teams synchronise on their own named barrier.
struct Team {
    int ts, tt, id;          // team size, thread index in the team, team index
};
__device__ __forceinline__ Team make_team(int n_teams) {
    Team tm;
    tm.ts = kThreads / n_teams;
    tm.tt = (int)threadIdx.x & (tm.ts - 1);
    tm.id = (int)threadIdx.x / tm.ts;
    return tm;
}
// (Without shared-memory prices -- 50k columns -- a scan is bound by the price reads through L2, side-by-side rows only
// thrash: measured 284 -> 298 ms; there a CTA keeps scanning one row at a time.)
__device__ __forceinline__ int teams_for(int rows, int max_teams) {
    const int n = rows >= 8 ? 8 : rows >= 4 ? 4 : rows >= 2 ? 2 : 1;
    return min(n, max_teams);
}
__device__ __forceinline__ void team_sync(const Team &tm) {
    if (tm.ts == kThreads) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + tm.id), "r"(tm.ts) : "memory");
}
constexpr long long kPackMax = (1ll << 46) - 1;
__device__ __forceinline__ unsigned long long pack_key(long long v, unsigned j) {
    return ((unsigned long long)min(v, kPackMax) << kPB) | j;
}
__device__ __forceinline__ Best packed_reduce(unsigned long long k1, unsigned long long k2, long long *red_b1,
                                              long long *red_b2, const Team &tm) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int w0 = (tm.id * tm.ts) >> 5, nw = tm.ts >> 5;        // the team's warps
    const unsigned long long w1 = warp_min64(k1);
    const unsigned long long w2 = warp_min64(k1 == w1 ? k2 : k1);
    unsigned long long *rk1 = reinterpret_cast<unsigned long long *>(red_b1);
    unsigned long long *rk2 = reinterpret_cast<unsigned long long *>(red_b2);
    if (lane == 0) { rk1[w] = w1; rk2[w] = w2; }
    team_sync(tm);
    Best s{LLONG_MAX, LLONG_MAX, -1};
    if (w == w0) {
        const unsigned long long q1 = lane < nw ? rk1[w0 + lane] : ~0ull, q2 = lane < nw ? rk2[w0 + lane] : ~0ull;
        const unsigned long long W1 = warp_min64(q1);
        const unsigned long long W2 = warp_min64(q1 == W1 ? q2 : q1);
        const long long v1 = (long long)(W1 >> kPB), v2 = (long long)(W2 >> kPB);
        s.j1 = W1 == ~0ull ? -1 : (int)(W1 & kPM);
        s.b1 = (W1 == ~0ull) ? LLONG_MAX : (v1 == kPackMax ? kInf : v1);
        s.b2 = (W2 == ~0ull) ? LLONG_MAX : (v2 == kPackMax ? kInf : v2);
    }
    team_sync(tm);
    return s;
}

// Team-wide scan of one person's row: min / second-min / argmin of (c-cmin)*S + price[.]; valid in the team's thread 0.
template <bool SMEMP>
__device__ __forceinline__ Best scan_row(const int32_t *__restrict__ r, int n, int cmin, int S,
                                         const long long *__restrict__ price, bool vec_ok,
                                         long long *red_b1, long long *red_b2, int *red_j, bool packed, const Team &tm) {
    Best s{LLONG_MAX, LLONG_MAX, -1};
    const int t = tm.tt, kStep = tm.ts;
    const unsigned long long pol = l2_policy_evict_first();
    int jtail = 0;
    if (vec_ok) {
        const int4 *r4 = reinterpret_cast<const int4 *>(r);
        const int n4 = n >> 2;
#pragma unroll 8
        for (int q = t; q < n4; q += kStep) {
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
    for (int j = jtail + t; j < n; j += kStep) {
        const long long p = SMEMP ? price[j] : __ldcg(price + j);
        upd(s, (long long)(__ldg(r + j) - cmin) * S + p, j);
    }
    if (packed) {
        const unsigned long long k1 = s.j1 >= 0 ? pack_key(s.b1, (unsigned)s.j1) : ~0ull;
        const unsigned long long k2 = s.b2 != LLONG_MAX ? pack_key(s.b2, (unsigned)kPM) : ~0ull;
        return packed_reduce(k1, k2, red_b1, red_b2, tm);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Best o;
        o.b1 = __shfl_xor_sync(0xffffffffu, s.b1, d);
        o.b2 = __shfl_xor_sync(0xffffffffu, s.b2, d);
        o.j1 = __shfl_xor_sync(0xffffffffu, s.j1, d);
        s = combine(s, o);
    }
    const int lane = (int)threadIdx.x & 31, w = (int)threadIdx.x >> 5;
    const int w0 = (tm.id * tm.ts) >> 5, nw = tm.ts >> 5;
    if (lane == 0) { red_b1[w] = s.b1; red_b2[w] = s.b2; red_j[w] = s.j1; }
    team_sync(tm);
    if (w == w0) {
        if (lane < nw) { s.b1 = red_b1[w0 + lane]; s.b2 = red_b2[w0 + lane]; s.j1 = red_j[w0 + lane]; }
        else { s.b1 = LLONG_MAX; s.b2 = LLONG_MAX; s.j1 = -1; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            Best o;
            o.b1 = __shfl_xor_sync(0xffffffffu, s.b1, d);
            o.b2 = __shfl_xor_sync(0xffffffffu, s.b2, d);
            o.j1 = __shfl_xor_sync(0xffffffffu, s.j1, d);
            s = combine(s, o);
        }
    }
    team_sync(tm);
    return s;
}

// Cheapest slot of object o (lowest slot index on ties) when slot `t_new` holds `p_new` and every
// other slot its stored price.
__device__ __forceinline__ void cheapest_slot(const SapParams &P, int o, int t_new, long long p_new, int &ms, long long &mp) {
    const int s0 = P.soff ? __ldg(P.soff + o) : o;
    const int s1 = P.soff ? __ldg(P.soff + o + 1) : o + 1;
    ms = s0; mp = (s0 == t_new) ? p_new : __ldcg(P.slot_price + s0);
#pragma unroll 4
    for (int t = s0 + 1; t < s1; ++t) {
        const long long p = (t == t_new) ? p_new : __ldcg(P.slot_price + t);
        if (p < mp) { mp = p; ms = t; }
    }
}

// SMEMP: prices (auction) / g (search) in shared memory.  SMEMO: slot-owner, tree-predecessor and
// (capacitated) cheapest-slot replicas in shared memory.
template <bool SMEMP, bool SMEMO>
__global__ void __launch_bounds__(kThreads, 1) lap_sap_kernel(const SapParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int np = P.P, no = P.O;
    size_t off = 0;
    long long *sarr = reinterpret_cast<long long *>(smem_raw);
    if (SMEMP) off += ((size_t)no * 8 + 15) / 16 * 16;
    long long *myqd = reinterpret_cast<long long *>(smem_raw + off);   // [qcap] labels of this CTA's frontier objects
    off += ((size_t)P.qcap * 8 + 15) / 16 * 16;
    int *myq = reinterpret_cast<int *>(smem_raw + off);                // [qcap] work queue (persons / frontier objects)
    off += ((size_t)P.qcap * 4 + 15) / 16 * 16;
    unsigned *sfront = reinterpret_cast<unsigned *>(smem_raw + off);   // [(O+31)/32] frontier of the round
    off += ((size_t)((no + 31) / 32) * 4 + 15) / 16 * 16;
    long long *sld = reinterpret_cast<long long *>(smem_raw + off);     // [wpc*32] labels of this CTA's slice (as classified)
    off += (size_t)P.wpc * 32 * 8;
    unsigned *sel_loc = reinterpret_cast<unsigned *>(smem_raw + off);  // [wpc] eligible objects of the slice
    off += ((size_t)P.wpc * 4 + 15) / 16 * 16;
    unsigned *sdirty_loc = reinterpret_cast<unsigned *>(smem_raw + off);   // [wpc] slice objects lowered since their holders last relaxed
    off += ((size_t)P.wpc * 4 + 15) / 16 * 16;
    int *swbase = reinterpret_cast<int *>(smem_raw + off);             // [(O+31)/32] frontier rank of a word's first bit
    off += ((size_t)((no + 31) / 32) * 4 + 15) / 16 * 16;
    unsigned *sreach = reinterpret_cast<unsigned *>(smem_raw + off);   // [(O+31)/32] SMEMO: objects the finished search reached
    off += ((size_t)((no + 31) / 32) * 4 + 15) / 16 * 16;
    unsigned *skept = reinterpret_cast<unsigned *>(smem_raw + off);    // [(O+31)/32] SMEMO: forest nodes the next search keeps
    off += ((size_t)((no + 31) / 32) * 4 + 15) / 16 * 16;
    unsigned *sdrop = reinterpret_cast<unsigned *>(smem_raw + off);    // [(O+31)/32] SMEMO: forest nodes it forgets
    off += ((size_t)((no + 31) / 32) * 4 + 15) / 16 * 16;
    int *sowner = reinterpret_cast<int *>(smem_raw + off);             // SMEMO: [P]
    int *spred = sowner + ((np + 3) & ~3);                             // SMEMO: [O]
    int *sminslot = spred + ((no + 3) & ~3);                           // SMEMO && soff: [O]

    __shared__ long long red_b1[32], red_b2[32];
    __shared__ int red_j[32], wcnt[32];
    __shared__ long long st_tm[8];
    __shared__ long long st_ph[5];     // CTA 0: kernel start, phase-section start, ns in phase starts / auction rounds / before the first phase
    __shared__ long long st_acc[12];   // bids, max bidders, -, small rounds, ns bid / barrier / replay / sap, relax hits, sap ns select, sap ns trace
    __shared__ int ssrc[kSapMax], sfo[kSapMax], ssmap[kSapMax];
    __shared__ long long sfo_d[kSapMax];
    __shared__ int hist[256];
    __shared__ long long sh_ll[4];      // broadcast slots: [0] D, [1] dmin, [2] dmax, [3] T
    __shared__ int sh_i[8];             // [0] wsum, [1] count, [2] nfront, [3] status, [4] n moves
    __shared__ int rw_person[kRowsMax], rw_slot[kRowsMax], rw_qi[kRowsMax], rw_slot2[kRowsMax], rw_qi2[kRowsMax];
    __shared__ long long rw_thr[kRowsMax];
    __shared__ int pth_ok[kMultiMax], pth_len[kMultiMax], pth_obj[kMultiMax];

    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int S = np + 1;
    const bool vec_ok = ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.cost) & 15) == 0);
    unsigned int bar_target = 0;
    // a wait that timed out: the launch is given up (status only; the outputs are undefined)
#define LAP_ABORT() do { if (threadIdx.x == 0) P.stats[0] = CYB_ERR_NOT_CONVERGED; return; } while (0)
#define GRID_BARRIER() do { if (grid_barrier(P.bar, bar_target, G, P.gmm + 5, __LINE__)) LAP_ABORT(); } while (0)
    const long long *price_rd = SMEMP ? sarr : P.lambda;
    const bool capd = P.soff != nullptr;

    auto rowptr = [&](int i) -> const int32_t * { return P.cost + (long long)i * P.ld; };
    auto capacity = [&](int o) -> int { return capd ? __ldg(P.soff + o + 1) - __ldg(P.soff + o) : 1; };
    auto owner_of = [&](int slot) -> int { return SMEMO ? sowner[slot] : __ldcg(P.slot_owner + slot); };
    auto obj_of_slot = [&](int slot) -> int { return capd ? __ldcg(P.slot_obj + slot) : slot; };

    if (b == 0 && t == 0) { st_ph[0] = global_ns(); st_ph[2] = 0; st_ph[3] = 0; st_ph[4] = 0; }
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
        if (lane == 0 && lmin <= lmax) { atomicMin(P.gmm + 0, lmin); atomicMax(P.gmm + 1, lmax); }
        for (int i = b * kThreads + t; i < np; i += G * kThreads) {
            P.slot_price[i] = 0; P.slot_owner[i] = -1; P.person_obj[i] = -1; P.person_slot[i] = -1;
        }
        for (int o = b * kThreads + t; o < no; o += G * kThreads) {
            P.lambda[o] = capacity(o) > 0 ? 0 : kInf;          // a spot that takes no cell is priced out
            P.minslot[o] = capd ? __ldg(P.soff + o) : o;
            P.bidw[0][o] = 0ull; P.bidw[1][o] = 0ull; P.bidw[2][o] = 0ull;
            if (capd) for (int s = __ldg(P.soff + o); s < __ldg(P.soff + o + 1); ++s) P.slot_obj[s] = o;
        }
        for (int o = b * kThreads + t; o < no + kSapMax; o += G * kThreads) P.claim[o] = INT_MAX;
        if (SMEMP) for (int o = t; o < no; o += kThreads) sarr[o] = capacity(o) > 0 ? 0 : kInf;
        if (SMEMO) {
            for (int k = t; k < np; k += kThreads) sowner[k] = -1;
            if (capd) for (int o = t; o < no; o += kThreads) sminslot[o] = __ldg(P.soff + o);
        }
        for (int w = t; w < (no + 31) / 32; w += kThreads) sfront[w] = 0u;
    }
    GRID_BARRIER();
    const int cmin = __ldcg(P.gmm + 0), cmax = __ldcg(P.gmm + 1);
    long long eps = ((long long)cmax - (long long)cmin) * S / P.eps0_div;
    if (eps < 1) eps = 1;

    int rounds = 0, phases = 0, searches = 0, srounds = 0, paths = 0, rid = 0;
    long long srows = 0;
    if (t < 12) st_acc[t] = 0;
    __syncthreads();
    int status = 0;
    // the input contract is |cost| < 2^30; a wider range would overflow the 32-bit (c - cmin)
    if ((long long)cmax - (long long)cmin >= (1ll << 31) - 1 || cmin <= -(1 << 30) || cmax >= (1 << 30)) status = CYB_ERR_OVERFLOW;
    int cur = 0, prevF = 0;
    const bool packed = ((long long)cmax - (long long)cmin + 1) * S < (1ll << 45) && P.packed_reduce;
    auto scan = [&](const int32_t *r, const Team &tm) -> Best {
        return scan_row<SMEMP>(r, no, cmin, S, price_rd, vec_ok, red_b1, red_b2, red_j, packed, tm);
    };
    const int sap_t = min(P.sap_t, kSapMax);

    while (!status) {
        ++phases;
        if (b == 0 && t == 0) { st_ph[1] = global_ns(); if (phases == 1) st_ph[4] = st_ph[1] - st_ph[0]; }
        // ---- phase start: which pairs survive eps-CS at the new eps? --------
        if (phases > 1) {
            // thread 0's certificate terms of a row (its object, slot, cost entry, slot price) are requested ahead of the
            // scan they belong to: no dependent round trip is left after the scan
            // (rows b, b + G, ... of this CTA, dealt over its scanning teams)
            const Team tm = make_team(teams_for(b < np ? (np - b - 1) / G + 1 : 0, P.max_teams));
            const bool lead = tm.tt == 0;
            const int istep = G * (kThreads / tm.ts), i0 = b + G * tm.id;
            int o_nx = i0 < np ? __ldcg(P.person_obj + i0) : -1;
            int ps_nx = (lead && i0 < np) ? __ldcg(P.person_slot + i0) : -1;
            for (int i = i0; i < np; i += istep) {
                const int o = o_nx, ps = ps_nx;
                if (i + istep < np) { o_nx = __ldcg(P.person_obj + i + istep); if (lead) ps_nx = __ldcg(P.person_slot + i + istep); }
                int f = 1;
                if (i + istep < np && P.prefetch) prefetch_row_l2(rowptr(i + istep), no, tm.tt);
                if (o >= 0) {
                    const int32_t *r = rowptr(i);
                    int c_o = 0; long long sp = 0;
                    if (lead) { c_o = __ldg(r + o); sp = __ldcg(P.slot_price + ps); }
                    const Best s = scan(r, tm);
                    if (lead) {
                        const long long alt = (s.j1 == o) ? s.b2 : s.b1;
                        const long long base = (long long)(c_o - cmin) * S;
                        const long long lam = SMEMP ? sarr[o] : __ldcg(P.lambda + o);
                        f = 0;
                        if (alt < kInf / 2) {
                            if (base + lam > alt + eps) f = ps + 2;                       // drop: vacate slot ps
                            else if (base + sp > alt + eps)
                                P.slot_price[ps] = alt + eps - base;                      // clamp (>= lambda[o])
                        }
                    }
                }
                if (lead) P.flag[i] = f;
            }
            GRID_BARRIER();
        }
        int F = 0;
        for (int i0 = 0; i0 < np; i0 += kThreads) {
            const int i = i0 + t;
            const int f = (i < np) ? (phases == 1 ? 1 : __ldcg(P.flag + i)) : 0;
            if (f >= 2) {
                if (i % G == b || !SMEMO) P.slot_owner[f - 2] = -1;      // the slot keeps its price
                if (SMEMO) sowner[f - 2] = -1;
                if (i % G == b) { P.person_obj[i] = -1; P.person_slot[i] = -1; }
            }
            int tot;
            const int pos = F + block_excl_count(f != 0, wcnt, tot);
            if (f != 0 && pos % G == b) { P.list[cur][pos] = i; myq[pos / G] = i; }
            F += tot;
        }
        if (phases > 1 && capd) {
            // clamps / searches may have left an equally cheap slot with a lower index: refresh the argmin
            for (int o = t; o < no; o += kThreads) {
                if (capacity(o) > 1) {
                    int ms; long long mp;
                    cheapest_slot(P, o, -1, 0, ms, mp);
                    if (o % G == b || !SMEMO) P.minslot[o] = ms;
                    if (SMEMO) sminslot[o] = ms;
                }
            }
        }
        __syncthreads();

        if (b == 0 && t == 0) { const long long n = global_ns(); st_ph[2] += n - st_ph[1]; st_ph[1] = n; }
        // A phase with eps > 1 ends as soon as <= partial persons are free: they stay free into the next phase (its
        // start re-derives the free list anyway), only the last phase has to place everybody.
        const int partial = eps > 1 ? P.partial : 0;
        // ---- bidding rounds (more than sap_t persons free) -------------------
        while (F > sap_t && F > partial) {
            if (++rounds > P.max_rounds) { status = CYB_ERR_NOT_CONVERGED; break; }
            if (t == 0) { st_acc[0] += F; if (F > st_acc[1]) st_acc[1] = F; }
            const bool timed = (b == 0 && t == 0 && F <= G);
            const long long tm0 = timed ? global_ns() : 0;
            const int myn = F > b ? (F - b - 1) / G + 1 : 0;
            // this CTA's bidders, dealt over its scanning teams (one team = the whole CTA when it has a single bidder)
            const Team tm = make_team(teams_for(myn, P.max_teams));
            const int nt = kThreads / tm.ts;
            if (tm.id < myn && P.prefetch) prefetch_row_l2(rowptr(myq[tm.id]), no, tm.tt);
            for (int q = tm.id; q < myn; q += nt) {
                const int i = myq[q];
                if (q + nt < myn && P.prefetch) prefetch_row_l2(rowptr(myq[q + nt]), no, tm.tt);
                const Best s = scan(rowptr(i), tm);
                if (tm.tt == 0) {
                    const int o = s.j1;
                    const long long lam = SMEMP ? sarr[o] : __ldcg(P.lambda + o);
                    const long long bid = lam + (s.b2 < kInf / 2 ? s.b2 - s.b1 : 0) + eps;
                    if (bid >= kBidLimit) atomicExch(P.gmm + 2, CYB_ERR_OVERFLOW);
                    const int slot = capd ? (SMEMO ? sminslot[o] : __ldcg(P.minslot + o)) : o;
                    const int prev = owner_of(slot);
                    P.rec[cur][q * G + b] = make_int4(o, slot, prev, 0);
                    atomicMax(P.bidw[cur] + o, ((unsigned long long)bid << kPB) | (kPM - (unsigned long long)i));
                }
            }
            const long long tm1 = timed ? global_ns() : 0;
            GRID_BARRIER();
            const long long tm2 = timed ? global_ns() : 0;
            // ---- resolve: every CTA replays every record ----------------------
            status = __ldcg(P.gmm + 2);
            const int nxt = cur == 2 ? 0 : cur + 1, prv = cur == 0 ? 2 : cur - 1;
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
                    const int wperson = (int)(kPM - (key & kPM));
                    if (wperson == i) {
                        const long long bid = (long long)(key >> kPB);
                        const bool mine = (k % G == b);
                        int ms; long long mp;
                        // sibling slot prices FIRST (loads before the stores to the same line, 148 CTAs storing to one line and then loading from it serialise at the L2 slice)
                        cheapest_slot(P, rc.x, rc.y, bid, ms, mp);
                        if (mine || !SMEMO) {
                            P.slot_owner[rc.y] = i;
                            if (capd) P.minslot[rc.x] = ms;
                        }
                        if (mine) P.slot_price[rc.y] = bid;
                        if (SMEMO) { sowner[rc.y] = i; if (capd) sminslot[rc.x] = ms; }
                        if (SMEMP) { sarr[rc.x] = mp; if (mine) P.lambda[rc.x] = mp; }
                        else P.lambda[rc.x] = mp;
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
        GRID_BARRIER();        // state of the last resolve becomes visible
        {
            const int prv = cur == 0 ? 2 : cur - 1;
            for (int k = b * kThreads + t; k < prevF; k += G * kThreads)
                P.bidw[prv][__ldcg(&P.rec[prv][k].x)] = 0ull;
            prevF = 0;
        }

        if (b == 0 && t == 0) st_ph[3] += global_ns() - st_ph[1];
        // ---- shortest-augmenting-path finish -------------------------------------
        long long step = eps;                    // frontier window of the searches, adapted round by round
        if (b == 0 && t == 0) st_tm[0] = global_ns();
        if (F > partial) {
            if (t < F) ssrc[t] = __ldcg(P.list[cur] + t);
            __syncthreads();
        }
        bool warm = false;                       // the forest of the previous search of this phase is available
        while (F > partial) {
            ++searches;
            if (searches >= kMaxSearch) { status = CYB_ERR_NOT_CONVERGED; break; }
            const int cbase = (kMaxSearch - searches) * kMultiMax;
            // S0: labels unreached, lists empty; shared memory switches from prices to g = lambda - d
            for (int o = b * kThreads + t; o < no; o += G * kThreads) {
                if (!warm) P.dkey[o] = ~0ull;                         // (a warm search keeps the labels written at the end of the previous one)
                if (!SMEMP) { const long long lam = __ldcg(P.lambda + o); P.gsnap[o] = lam >= kInf / 2 ? kGInf : lam - kInf; }
            }
            for (int w = b * kThreads + t; w < 3 * ((no + 31) / 32); w += G * kThreads) P.chgbits[w / ((no + 31) / 32)][w % ((no + 31) / 32)] = 0u;
            if (b == 0 && t < 24) { if ((t & 7) == 4 || (t & 7) == 5) P.rstat[t] = -1; else P.rstat[t] = 0; }      // dmin = all ones
            if (SMEMP) for (int o = t; o < no; o += kThreads) { const long long l = sarr[o]; sarr[o] = l >= kInf / 2 ? kGInf : l - kInf; }
            GRID_BARRIER();
            // (every CTA has finished applying the previous search's moves: their buffers can be reset)
            if (b == 0 && t == 0) P.nmoves[0] = 0;
            if (b == 0 && t < kSapMax) P.srcdone[t] = 0;
            // objects with a free slot, in increasing order (every CTA builds the same list)
            int nfo = 0;
            for (int o0 = 0; o0 < no; o0 += kThreads) {
                const int o = o0 + t;
                bool fr = false;
                if (o < no) {
                    if (!capd) fr = owner_of(o) < 0;
                    else for (int s = __ldg(P.soff + o); s < __ldg(P.soff + o + 1); ++s) fr = fr || owner_of(s) < 0;
                }
                int tot;
                const int pos = nfo + block_excl_count(fr, wcnt, tot);
                if (fr && pos < kSapMax) sfo[pos] = o;
                nfo += tot;
            }
            if (nfo > kSapMax || nfo < 1) { status = CYB_ERR_NOT_CONVERGED; break; }     // cannot happen: free slots == free persons
            const int want = min(min(F, nfo), P.multi);
            __syncthreads();
            // S1: round 0 -- the free persons relax their rows, values relative to the row minimum
            ++rid; ++srounds;
            for (int k = b; k < F; k += G) {
                const int i = ssrc[k];
                const int32_t *r = rowptr(i);
                // over C + g (g = lambda - 2^60: a common shift; negative, hence the unpacked reduction)
                const Best s = scan_row<SMEMP>(r, no, cmin, S, price_rd, vec_ok, red_b1, red_b2, red_j, false, make_team(1));
                if (t == 0) sh_ll[0] = s.b1;
                __syncthreads();
                const long long b1 = sh_ll[0];
                const unsigned long long src = (unsigned long long)(np + k);
                for (int j = t; j < no; j += kThreads) {
                    const long long p = SMEMP ? sarr[j] : __ldcg(P.lambda + j);
                    if (p >= kInf / 2) continue;                      // priced out (kGInf in g form as well)
                    const long long nd = (long long)(__ldg(r + j) - cmin) * S + p - b1;
                    if (nd >= kBidLimit) { atomicExch(P.gmm + 2, CYB_ERR_OVERFLOW); continue; }
                    const unsigned long long key = ((unsigned long long)nd << kPB) | src;
                    if (key < __ldcg(P.dkey + j)) atomicMin(P.dkey + j, key);
                }
                srows += 2;
                __syncthreads();
            }
            GRID_BARRIER();
            // S2: search rounds.  The selection is DISTRIBUTED: every CTA classifies its own slice of the
            // objects (one warp per 32-object word: label, lowered-bitmap word, holder), appends the candidates
            // (eligible labels <= Tg) to a global list and adds its statistics; after a grid barrier every CTA
            // reads the (short) candidate list, derives the same threshold T and frontier from it and takes the
            // frontier entries of rank = b (mod G).  Two grid barriers per round; per-CTA work is a handful of
            // loads.  A CTA's dirty words (lowered since their holders last relaxed) stay in its shared memory.
            long long D = kInf;
            int rr3 = 0;                         // round index mod 3: bitmap read / candidate list of this round
            bool implicit = true;                // round 1: every reached object is dirty
            const int nwords = (no + 31) / 32;
            const int w0 = b * P.wpc, nmy = max(0, min(nwords, w0 + P.wpc) - w0);
            for (int ww = t; ww < P.wpc; ww += kThreads) sdirty_loc[ww] = 0u;
            long long Tg = step;                 // this round's guess: candidates are the eligible labels <= Tg (round 0 leaves labels 0)
            bool repair = warm;                  // first round of a warm search: every surviving tree node relaxes again
            for (;;) {
                ++rid;
                if (b == 0 && t == 0) st_tm[1] = global_ns();
                const int nb = rr3 == 2 ? 0 : rr3 + 1, zb = rr3 == 0 ? 2 : rr3 - 1;
                const unsigned *CB = P.chgbits[rr3];
                unsigned *NB = P.chgbits[nb];
                unsigned long long *CAND = P.cand[rr3];
                int *RS = P.rstat + 8 * rr3;             // {ncand, wC, wE, -, dmin (64 bit), -, -}
                // buffers of the previous round: read by nobody any more, written again in the next round
                for (int w = b * kThreads + t; w < nwords; w += G * kThreads) P.chgbits[zb][w] = 0u;
                if (b == 0 && t == 0) {
                    int *Z = P.rstat + 8 * nb;
                    Z[0] = 0; Z[1] = 0; Z[2] = 0; *reinterpret_cast<unsigned long long *>(Z + 4) = ~0ull;
                }
                // ---- phase 1: all the loads of the phase are requested first (one round trip) ----
                if (t < nfo) { const unsigned long long key = __ldcg(P.dkey + sfo[t]); sfo_d[t] = key == ~0ull ? kInf : (long long)(key >> kPB); }
                const int o_pre = (w0 + warp) * 32 + lane;
                const bool has_pre = warp < nmy && o_pre < no;
                const unsigned long long key_pre = has_pre ? __ldcg(P.dkey + o_pre) : ~0ull;
                const unsigned cbw_pre = (warp < nmy && !implicit) ? __ldcg(CB + w0 + warp) : ~0u;
                if (t == 0) sh_ll[0] = kInf;
                if (t < 256) hist[t] = 0;
                // the objects the previous round lowered, as a bitmap in shared memory + prefix counts: their
                // g = lambda - d and tree predecessor replicas are refreshed below
                int nchg = 0;
                if (SMEMP || SMEMO) {
                    for (int wb = 0; wb < nwords; wb += kThreads) {
                        const int w = wb + t;
                        unsigned cw = 0;
                        if (w < nwords) {
                            cw = implicit ? ((w == nwords - 1 && (no & 31)) ? ((1u << (no & 31)) - 1u) : ~0u) : __ldcg(CB + w);
                            sfront[w] = cw;
                        }
                        const int c = __popc(cw);
                        int inc = c;
#pragma unroll
                        for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += y; }
                        if (lane == 31) wcnt[warp] = inc;
                        __syncthreads();
                        int wv = lane < kWarps ? wcnt[lane] : 0, winc = wv;
#pragma unroll
                        for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(0xffffffffu, winc, dd); if (lane >= dd) winc += y; }
                        const int woff = __shfl_sync(0xffffffffu, winc - wv, warp);
                        const int tot = __shfl_sync(0xffffffffu, winc, 31);
                        if (w < nwords) swbase[w] = nchg + woff + inc - c;
                        nchg += tot;
                        __syncthreads();
                    }
                } else {
                    __syncthreads();
                }
                // D = the want-th smallest label of an object with a free slot
                if (t < nfo) {
                    const long long dm = sfo_d[t];
                    int rank = 0;
                    for (int u = 0; u < nfo; ++u) { const long long du = sfo_d[u]; rank += (du < dm || (du == dm && u < t)) ? 1 : 0; }
                    if (rank == want - 1) sh_ll[0] = dm;
                }
                __syncthreads();
                D = sh_ll[0];
                if (D >= kInf / 2) { status = CYB_ERR_NOT_CONVERGED; break; }     // cannot happen: round 0 reaches every object
                if (step > D) step = D;
                if (step < 1) step = 1;
                if (Tg > D - 1 || repair) Tg = D - 1;                        // every eligible label is below D
                // replicas of the lowered objects: position i of the bitmap -> word by binary search over the prefix counts.
                // The first object of every thread is requested NOW, so that its round trip overlaps the atomicAdd
                // round trip of the classification below.
                auto lowered_object = [&](int i) -> int {
                    int lo = 0, hi = nwords - 1;
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (swbase[mid] <= i) lo = mid; else hi = mid - 1; }
                    return lo * 32 + (int)__fns(sfront[lo], 0, i - swbase[lo] + 1);
                };
                int ro = -1; unsigned long long rkey = ~0ull; long long rlam = 0;
                if ((SMEMP || SMEMO) && t < nchg) {
                    ro = lowered_object(t);
                    rkey = __ldcg(P.dkey + ro);
                    if (SMEMP) rlam = __ldcg(P.lambda + ro);
                }
                // this CTA's slice: eligible = dirty, held and below D; candidates = eligible with label <= Tg
                for (int ww = warp; ww < nmy; ww += kWarps) {
                    const int w = w0 + ww, o = w * 32 + lane;
                    const unsigned long long key = ww == warp ? key_pre : (o < no ? __ldcg(P.dkey + o) : ~0ull);
                    const unsigned cbw = ww == warp ? cbw_pre : (implicit ? ~0u : __ldcg(CB + w));
                    bool el = false; long long d = kInf;
                    if (o < no && key != ~0ull) {
                        d = (long long)(key >> kPB);
                        if ((((cbw | sdirty_loc[ww]) >> lane) & 1u) && d < D) {
                            if (!capd) el = owner_of(o) >= 0;
                            else for (int sl = __ldg(P.soff + o); sl < __ldg(P.soff + o + 1); ++sl) el = el || owner_of(sl) >= 0;
                        }
                    }
                    sld[ww * 32 + lane] = d;
                    if (!SMEMP && o < no && key != ~0ull && ((cbw >> lane) & 1u)) {
                        // without shared-memory prices the round-start snapshot g = lambda - label lives in L2, kept by the slice owners
                        const long long lam = __ldcg(P.lambda + o);
                        P.gsnap[o] = lam >= kInf / 2 ? kGInf : lam - d;
                    }
                    const bool cnd = el && d <= Tg;
                    const unsigned mel = __ballot_sync(0xffffffffu, el), mc = __ballot_sync(0xffffffffu, cnd);
                    int base = 0;
                    if (lane == 0) { sel_loc[ww] = mel; if (mc) base = atomicAdd(RS + 0, __popc(mc)); }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (cnd) CAND[base + __popc(mc & ((1u << lane) - 1u))] = ((unsigned long long)d << kPB) | (unsigned)o;
                    const int cap = el ? capacity(o) : 0;
                    int lE = cap, lC = cnd ? cap : 0; long long lmin = el ? d : kInf;
#pragma unroll
                    for (int dd = 16; dd > 0; dd >>= 1) {
                        lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, dd));
                        lC += __shfl_xor_sync(0xffffffffu, lC, dd);
                        lE += __shfl_xor_sync(0xffffffffu, lE, dd);
                    }
                    if (lane == 0 && lE > 0) {
                        if (lC) atomicAdd(RS + 1, lC);
                        atomicAdd(RS + 2, lE);
                        atomicMin(reinterpret_cast<unsigned long long *>(RS + 4), (unsigned long long)lmin);
                    }
                }
                // (the replicas of the lowered objects: loads requested above, stored here; the rest of a long list follows)
                if (SMEMP || SMEMO) {
                    if (ro >= 0 && rkey != ~0ull) {
                        if (SMEMP) sarr[ro] = rlam >= kInf / 2 ? kGInf : rlam - (long long)(rkey >> kPB);
                        if (SMEMO) spred[ro] = (int)(rkey & kPM);
                    }
#pragma unroll 2
                    for (int i = t + kThreads; i < nchg; i += kThreads) {
                        const int o = lowered_object(i);
                        const unsigned long long key = __ldcg(P.dkey + o);
                        if (key != ~0ull) {
                            if (SMEMP) { const long long lam = __ldcg(P.lambda + o); sarr[o] = lam >= kInf / 2 ? kGInf : lam - (long long)(key >> kPB); }
                            if (SMEMO) spred[o] = (int)(key & kPM);
                        }
                    }
                }
                if (b == 0 && t == 0) st_acc[2] += global_ns() - st_tm[1];
                GRID_BARRIER();                                          // candidates and statistics are complete; labels may be lowered from here on
                // ---- phase 2: every CTA derives the same threshold and frontier from the candidate list ----
                if (b == 0 && t == 0) st_tm[2] = global_ns();
                unsigned long long e_pre = t < no ? __ldcg(CAND + t) : 0ull;            // (speculative: before the count is known)
                if (t == 0) {
                    sh_i[0] = __ldcg(RS + 0); sh_i[1] = __ldcg(RS + 1); sh_i[2] = __ldcg(RS + 2);
                    sh_ll[1] = (long long)__ldcg(reinterpret_cast<unsigned long long *>(RS + 4));
                }
                __syncthreads();
                const int ncand = sh_i[0], wC = sh_i[1], wE = sh_i[2];
                const long long dmin_el = sh_ll[1];
                if (wE == 0) break;                                     // fixed point below D: the search is over
                if (wC == 0) {
                    // the guess selected nothing (first round of the search, or D moved below it): restart from the
                    // smallest eligible label.  Nothing was relaxed: the next round sees the same labels.
                    Tg = dmin_el + step;
                    for (int ww = t; ww < nmy; ww += kThreads) sdirty_loc[ww] = sel_loc[ww];
                    implicit = false;
                    rr3 = nb;
                    GRID_BARRIER();
                    continue;
                }
                if (++srounds + rounds > P.max_rounds) { status = CYB_ERR_NOT_CONVERGED; break; }
                // T = about K rows' worth of the smallest candidates: 256 power-of-two bins over [dmin_el, Tg], upper
                // edge of the first bin where the cumulative slot count reaches K (Tg when the candidates hold fewer)
                long long T = Tg;
                if (wC > P.sap_k && !repair) {
                    int sh = 0;
                    while (((Tg - dmin_el) >> sh) >= 256) ++sh;
                    for (int e = t; e < ncand; e += kThreads) {
                        const unsigned long long ent = e == t ? e_pre : __ldcg(CAND + e);
                        atomicAdd(&hist[(int)(((long long)(ent >> kPB) - dmin_el) >> sh)], capacity((int)(ent & kPM)));
                    }
                    __syncthreads();
                    if (warp == 0) {
                        int c8[8], sum = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) { c8[q] = hist[lane * 8 + q]; sum += c8[q]; }
                        int inc = sum;
#pragma unroll
                        for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += y; }
                        int run = inc - sum, bsel = INT_MAX;
#pragma unroll
                        for (int q = 0; q < 8; ++q) { run += c8[q]; if (run >= P.sap_k && bsel == INT_MAX) bsel = lane * 8 + q; }
                        bsel = __reduce_min_sync(0xffffffffu, (unsigned)bsel);
                        if (bsel == INT_MAX) bsel = 255;
                        if (lane == 0) sh_ll[3] = min(Tg, dmin_el + (((long long)bsel + 1) << sh) - 1);
                    }
                    __syncthreads();
                    T = sh_ll[3];
                }
                // frontier = candidates with label <= T, dealt to the CTAs by their rank in the list
                int nfront = 0;
                for (int e0 = 0; e0 < ncand; e0 += kThreads) {
                    const int e = e0 + t;
                    const unsigned long long ent = e < ncand ? (e0 == 0 ? e_pre : __ldcg(CAND + e)) : ~0ull;
                    const long long d = (long long)(ent >> kPB);
                    const bool fr = e < ncand && d <= T;
                    int tot;
                    const int pos = nfront + block_excl_count(fr, wcnt, tot);
                    if (fr) {
                        const int o = (int)(ent & kPM);
                        if (pos % G == b) { myq[pos / G] = o; myqd[pos / G] = d; }
                    }
                    nfront += tot;
                }
                // this CTA's dirty words: the eligible objects that were not selected
                for (int ww = warp; ww < nmy; ww += kWarps) {
                    const unsigned mf = __ballot_sync(0xffffffffu, sld[ww * 32 + lane] <= T);
                    if (lane == 0) sdirty_loc[ww] = sel_loc[ww] & ~mf;
                }
                // the window doubles when it held fewer than K although more was eligible, halves above 4K
                const bool flood = repair;
                if (repair) repair = false;
                else if (wC < P.sap_k && wC < wE) step *= 2;
                else if (wC > 4 * P.sap_k && step > 1) step /= 2;
                Tg = T + step;
                const int myn = nfront > b ? (nfront - b - 1) / G + 1 : 0;
                __syncthreads();
                if (b == 0 && t == 0) { st_tm[3] = global_ns(); st_acc[8] += st_tm[3] - st_tm[2]; }
                // relax: the holders of this CTA's frontier objects, kRowsMax rows at a time
                int qi = 0, si = 0;          // next frontier object of this CTA, next slot inside it
                while (qi < myn) {
                    __syncthreads();
                    if (t == 0) {
                        // the next <= kRowsMax slots of the queue
                        int nr = 0;
                        while (qi < myn && nr < kRowsMax) {
                            const int o = myq[qi];
                            const int s0 = capd ? __ldg(P.soff + o) : o, s1 = capd ? __ldg(P.soff + o + 1) : o + 1;
                            int s = s0 + si;
                            for (; s < s1 && nr < kRowsMax; ++s) { rw_slot[nr] = s; rw_qi[nr] = qi; ++nr; }
                            if (s >= s1) { ++qi; si = 0; } else si = s - s0;
                        }
                        sh_i[5] = nr; sh_i[6] = qi; sh_i[7] = si;
                    }
                    __syncthreads();
                    const int nr = sh_i[5];
                    qi = sh_i[6]; si = sh_i[7];
                    // holders (empty slots are squeezed out), in slot order
                    {
                        const int i = t < nr ? owner_of(rw_slot[t]) : -1;
                        const unsigned m = __ballot_sync(0xffffffffu, i >= 0);
                        if (t < 32) {
                            if (i >= 0) { const int p = __popc(m & ((1u << lane) - 1u)); rw_person[p] = i; rw_slot2[p] = rw_slot[t]; rw_qi2[p] = rw_qi[t]; }
                            if (t == 0) sh_i[3] = __popc(m);
                        }
                    }
                    __syncthreads();
                    const int nrv = sh_i[3];
                    srows += nrv;
                    // the entry C[i, o] of every row (for its threshold) is requested together with the first
                    // wave of row data: one round trip for both
                    int cval = 0;
                    if (t < nrv) cval = __ldg(rowptr(rw_person[t]) + myq[rw_qi2[t]]);
                    const unsigned long long pol = l2_policy_evict_first();
                    const int n4 = vec_ok ? (no >> 2) : 0;
                    const unsigned total4 = (unsigned)nrv * (unsigned)n4;          // <= 32 rows x 65k int4: 32-bit index arithmetic
                    const unsigned un4 = (unsigned)max(n4, 1);
                    int4 wv0[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned idx = (unsigned)t + (unsigned)u * kThreads;
                        wv0[u] = make_int4(0, 0, 0, 0);
                        if (idx < total4) {
                            const int row = (int)(idx / un4), q = (int)(idx - (unsigned)row * un4);
                            wv0[u] = ld_stream(reinterpret_cast<const int4 *>(rowptr(rw_person[row])) + q, pol);
                        }
                    }
                    if (t < nrv) {
                        // thr = C[i,o] + lambda[o] - d[o] - eps;  a relaxation of k gives  nd = C[i,k] + lambda[k] - thr
                        const int q = rw_qi2[t], o = myq[q];
                        const long long lam_minus_d = SMEMP ? sarr[o] : __ldcg(P.lambda + o) - myqd[q];
                        rw_thr[t] = (long long)(cval - cmin) * S + lam_minus_d - eps;
                    }
                    __syncthreads();
                    // one element: g = lambda[k] - (label of k when the round started) from shared memory (SMEMP) or from the
                    // snapshot array the slice owners keep in L2; the relaxation counts iff it is STRICTLY below that label
                    auto relax = [&](int k, int c, long long thr, unsigned long long slot, long long g) {
                        const long long v = (long long)(c - cmin) * S;
                        if (v + g < thr) {
                            const long long nd = v + __ldcg(P.lambda + k) - thr;
                            if (nd >= kBidLimit) { atomicExch(P.gmm + 2, CYB_ERR_OVERFLOW); return; }
                            // strictly below the round-start label: whoever wins the minimum, the label IS lowered in this
                            // round, so the bit can be set without waiting for the atomic's return value (both are
                            // fire-and-forget reductions: no round trip on the relax path)
                            // (in the repair round of a warm search thousands of rows lower the same objects: a racy look at the
                            // current label first -- most of them lose to the running minimum, and a read is far cheaper than an
                            // atomic serialised at one L2 slice; same minimum.  Elsewhere hits are rare and the extra read only stalls.)
                            const unsigned long long key = ((unsigned long long)nd << kPB) | slot;
                            const unsigned bit = 1u << (k & 31);
                            if (flood) {
                                if (key < __ldcg(P.dkey + k)) atomicMin(P.dkey + k, key);
                                if (!(__ldcg(NB + (k >> 5)) & bit)) atomicOr(NB + (k >> 5), bit);
                            } else {
                                atomicMin(P.dkey + k, key);
                                atomicOr(NB + (k >> 5), bit);
                            }
                        }
                    };
                    auto relax4 = [&](int row, int q, const int4 &c) {
                        const long long thr = rw_thr[row];
                        const unsigned long long slot = (unsigned long long)rw_slot2[row];
                        const int j = q << 2;
                        if (SMEMP) {
                            const longlong2 a = *reinterpret_cast<const longlong2 *>(sarr + j);
                            const longlong2 bb = *reinterpret_cast<const longlong2 *>(sarr + j + 2);
                            relax(j, c.x, thr, slot, a.x); relax(j + 1, c.y, thr, slot, a.y);
                            relax(j + 2, c.z, thr, slot, bb.x); relax(j + 3, c.w, thr, slot, bb.y);
                        } else {
                            const longlong2 a = __ldcg(reinterpret_cast<const longlong2 *>(P.gsnap + j));
                            const longlong2 bb = __ldcg(reinterpret_cast<const longlong2 *>(P.gsnap + j + 2));
                            relax(j, c.x, thr, slot, a.x); relax(j + 1, c.y, thr, slot, a.y);
                            relax(j + 2, c.z, thr, slot, bb.x); relax(j + 3, c.w, thr, slot, bb.y);
                        }
                    };
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned idx = (unsigned)t + (unsigned)u * kThreads;
                        if (idx < total4) { const int row = (int)(idx / un4); relax4(row, (int)(idx - (unsigned)row * un4), wv0[u]); }
                    }
#pragma unroll 4
                    for (unsigned idx = (unsigned)t + 4u * kThreads; idx < total4; idx += kThreads) {
                        const int row = (int)(idx / un4), q = (int)(idx - (unsigned)row * un4);
                        relax4(row, q, ld_stream(reinterpret_cast<const int4 *>(rowptr(rw_person[row])) + q, pol));
                    }
                    // columns beyond the vector part
                    for (int row = 0; row < nrv; ++row) {
                        const int32_t *r = rowptr(rw_person[row]);
                        for (int j = (n4 << 2) + t; j < no; j += kThreads) {
                            const int cj = __ldg(r + j);
                            if (SMEMP) relax(j, cj, rw_thr[row], (unsigned long long)rw_slot2[row], sarr[j]);
                            else relax(j, cj, rw_thr[row], (unsigned long long)rw_slot2[row], __ldcg(P.gsnap + j));
                        }
                    }
                }
                if (b == 0 && t == 0) { st_acc[9] += st_tm[3] - st_tm[1]; st_acc[10] += global_ns() - st_tm[3]; }
                implicit = false;
                rr3 = nb;
                GRID_BARRIER();
                status = __ldcg(P.gmm + 2);
                if (status) break;
            }
            if (status) break;
            if (D >= kInf / 2) { status = CYB_ERR_NOT_CONVERGED; break; }
            if (b == 0 && t == 0) st_tm[4] = global_ns();
            // S3a: price update -- lambda[o] += D - d[o] below D; shared memory goes back to prices
            for (int o0 = 0; o0 < no; o0 += kThreads) {
                const int o = o0 + t;
                const bool mine = (o % G == b);
                unsigned long long key = ~0ull;
                if (o < no && (SMEMP || mine)) {
                    key = __ldcg(P.dkey + o);
                    const long long d = key == ~0ull ? kInf : (long long)(key >> kPB);
                    long long lam;
                    if (SMEMP) { const long long g = sarr[o]; lam = g >= kGInf / 2 ? kInf : g + d; }
                    else lam = __ldcg(P.lambda + o);
                    if (d < D) {
                        lam += D - d;
                        if (mine) {
                            if (lam >= kBidLimit) atomicExch(P.gmm + 2, CYB_ERR_OVERFLOW);
                            P.lambda[o] = lam;
                            const int s0 = capd ? __ldg(P.soff + o) : o, s1 = capd ? __ldg(P.soff + o + 1) : o + 1;
                            for (int s = s0; s < s1; ++s) if (__ldcg(P.slot_price + s) < lam) P.slot_price[s] = lam;
                        }
                    }
                    if (SMEMP) sarr[o] = lam;
                }
                if (SMEMO) {                                           // (SMEMO implies SMEMP: every thread read its key)
                    const unsigned mr = __ballot_sync(0xffffffffu, key != ~0ull);
                    if (lane == 0 && (o >> 5) < (no + 31) / 32) sreach[o >> 5] = mr;
                }
            }
            // S3b: CTA 0 traces the candidate paths (one lane of warp 0 each) and publishes the accepted moves
            if (b == 0) {
                // candidate c = the free object of rank c by (label, object)
                if (t < nfo) {
                    const long long dm = sfo_d[t];
                    int rank = 0;
                    for (int u = 0; u < nfo; ++u) { const long long du = sfo_d[u]; rank += (du < dm || (du == dm && u < t)) ? 1 : 0; }
                    if (rank < want) pth_obj[rank] = sfo[t];
                }
                __syncthreads();
                auto pred_of = [&](int o) -> int { return SMEMO ? spred[o] : (int)(__ldcg(P.dkey + o) & kPM); };
                if (t < want) {
                    int o = pth_obj[t], len = 0;
                    for (;;) {
                        atomicMin(P.claim + o, cbase + t);
                        const int s = pred_of(o);
                        ++len;
                        if (s >= np) { atomicMin(P.claim + no + (s - np), cbase + t); break; }
                        o = obj_of_slot(s);
                        if (len > np) break;                         // cannot happen (the tree has no cycles)
                    }
                    pth_len[t] = len;
                }
                __syncthreads();
                if (t < want) {
                    int o = pth_obj[t], ok = 1, len = 0;
                    for (;;) {
                        ok &= (__ldcg(P.claim + o) == cbase + t);
                        const int s = pred_of(o);
                        if (s >= np) { ok &= (__ldcg(P.claim + no + (s - np)) == cbase + t); break; }
                        o = obj_of_slot(s);
                        if (++len > np) { ok = 0; break; }
                    }
                    pth_ok[t] = ok;
                }
                __syncthreads();
                if (t < want && pth_ok[t]) {
                    int base = 0;
                    for (int u = 0; u < t; ++u) if (pth_ok[u]) base += pth_len[u];
                    int o = pth_obj[t];
                    int slot = -1;
                    {
                        const int s0 = capd ? __ldg(P.soff + o) : o, s1 = capd ? __ldg(P.soff + o + 1) : o + 1;
                        for (int s = s0; s < s1; ++s) if (owner_of(s) < 0) { slot = s; break; }
                    }
                    for (;;) {
                        const int s = pred_of(o);
                        const int p = s >= np ? ssrc[s - np] : owner_of(s);
                        P.moves[base++] = make_int4(p, o, slot, 0);
                        if (s >= np) { P.srcdone[s - np] = 1; break; }
                        o = obj_of_slot(s); slot = s;
                    }
                }
                __syncthreads();
                if (t == 0) {
                    int tot = 0, np_ = 0;
                    for (int u = 0; u < want; ++u) if (pth_ok[u]) { tot += pth_len[u]; ++np_; }
                    P.nmoves[0] = tot;
                    sh_i[4] = np_;
                }
                __syncthreads();
            }
            GRID_BARRIER();
            status = __ldcg(P.gmm + 2);
            if (status) break;
            // S3c: every CTA applies the moves to its replicas; global state has one writer per move
            {
                const int nm = __ldcg(P.nmoves);
                for (int k = t; k < nm; k += kThreads) {
                    const int4 mv = __ldcg(P.moves + k);
                    if (SMEMO) sowner[mv.z] = mv.x;
                    if (k % G == b) {
                        P.slot_owner[mv.z] = mv.x; P.person_obj[mv.x] = mv.y; P.person_slot[mv.x] = mv.z;
                        P.slot_price[mv.z] = SMEMP ? sarr[mv.y] : __ldcg(P.lambda + mv.y);
                    }
                }
                // the free persons that are still free, in order
                const bool still = t < F && __ldcg(P.srcdone + t) == 0;
                const int me = t < F ? ssrc[t] : -1;
                int tot;
                const int pos = block_excl_count(still, wcnt, tot);
                if (still) ssrc[pos] = me;
                if (t < F) ssmap[t] = still ? pos : -1;
                paths += F - tot;
                if (tot == F) { status = CYB_ERR_NOT_CONVERGED; }        // no path applied: cannot happen
                const int Fold = F;
                F = tot;
                __syncthreads();
                warm = false;
                if (SMEMO && P.warm && F > partial && !status) {
                    // ---- the next search starts from the surviving part of this one's shortest-path forest: a node is
                    // kept iff its root is still free and its chain avoids every object of an applied path.  Its label
                    // shifts by D (tree arcs stay consistent under the price update above); the rest is forgotten.
                    // Every CTA derives the same bitmasks from its replicas (predecessors, reached bits, moves).
                    const int nwd = (no + 31) / 32;
                    int *anc = P.anc + (size_t)b * no;
                    for (int w = t; w < nwd; w += kThreads) { skept[w] = 0u; sdrop[w] = 0u; }
                    __syncthreads();
                    for (int k = t; k < nm; k += kThreads) { const int o = __ldcg(P.moves + k).y; atomicOr(&sdrop[o >> 5], 1u << (o & 31)); }
                    __syncthreads();
                    for (int o0 = 0; o0 < no; o0 += kThreads) {
                        const int o = o0 + t;
                        bool kp = false, dp = false;
                        if (o < no) {
                            const bool reached = (sreach[o >> 5] >> (o & 31)) & 1u, onp = (sdrop[o >> 5] >> (o & 31)) & 1u;
                            if (!reached || onp) dp = true;
                            else { const int sl = spred[o]; if (sl >= np) { if (sl - np < Fold && ssmap[sl - np] >= 0) kp = true; else dp = true; } }
                        }
                        // an undecided node starts with its tree parent as ancestor (this CTA's scratch row in L2)
                        if (o < no && !kp && !dp) __stcg(anc + o, obj_of_slot(spred[o]));
                        const unsigned mk = __ballot_sync(0xffffffffu, kp), md = __ballot_sync(0xffffffffu, dp);
                        __syncwarp();
                        if (lane == 0 && (o >> 5) < nwd) { skept[o >> 5] = mk; sdrop[o >> 5] |= md; }
                    }
                    // pointer jumping: an undecided node takes the state of its ancestor, or jumps to the ancestor's ancestor
                    // (chains are tens of nodes deep: a handful of iterations instead of one per level)
                    for (int it = 0; it <= 40; ++it) {
                        __syncthreads();
                        bool undecided = false;
                        for (int o0 = 0; o0 < no; o0 += kThreads) {
                            const int o = o0 + t;
                            bool kp = false, dp = false;
                            if (o < no && !(((skept[o >> 5] | sdrop[o >> 5]) >> (o & 31)) & 1u)) {
                                const int a = __ldcg(anc + o);
                                kp = (skept[a >> 5] >> (a & 31)) & 1u;
                                dp = (sdrop[a >> 5] >> (a & 31)) & 1u;
                                if (!(kp || dp)) { __stcg(anc + o, __ldcg(anc + a)); undecided = true; }
                            }
                            const unsigned mk = __ballot_sync(0xffffffffu, kp), md = __ballot_sync(0xffffffffu, dp);
                            if (lane == 0 && (mk | md)) { atomicOr(&skept[o >> 5], mk); atomicOr(&sdrop[o >> 5], md); }
                        }
                        if (!__syncthreads_or(undecided)) break;
                    }
                    // the owners rewrite the labels (nobody reads them before the barrier of the next search)
                    for (int o = b * kThreads + t; o < no; o += G * kThreads) {
                        unsigned long long key = ~0ull;
                        if ((skept[o >> 5] >> (o & 31)) & 1u) {
                            const unsigned long long old = __ldcg(P.dkey + o);
                            const long long d = (long long)(old >> kPB);
                            int sl = (int)(old & kPM);
                            if (sl >= np) sl = np + ssmap[sl - np];
                            key = ((unsigned long long)(d > D ? d - D : 0) << kPB) | (unsigned)sl;
                        }
                        P.dkey[o] = key;
                    }
                    warm = true;
                }
            }
            if (b == 0 && t == 0) st_acc[11] += global_ns() - st_tm[4];
            if (status) break;
            if (F > partial && !SMEMO) GRID_BARRIER();     // the next search reads slot owners from global memory
        }
        if (b == 0 && t == 0) st_acc[7] += global_ns() - st_tm[0];
        if (status) break;
        GRID_BARRIER();        // moves / prices of the last search become visible
        status = __ldcg(P.gmm + 2);
        if (status) break;
        if (eps == 1) break;
        eps /= P.theta;
        if (eps < 1) eps = 1;
    }
    if (status) atomicExch(P.gmm + 2, status);

    // ---- total cost of the assignment -------------------------------------------
    if (!status) {
        long long sum = 0;
        for (int i = b * kThreads + t; i < np; i += G * kThreads) {
            const int o = __ldcg(P.person_obj + i);
            if (o >= 0) sum += (long long)__ldg(rowptr(i) + o);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (lane == 0 && sum != 0) atomicAdd(reinterpret_cast<unsigned long long *>(P.total), (unsigned long long)sum);
    }
    if (t == 0 && srows) atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 13), (unsigned long long)srows);
    if (b == 0 && t == 0) {
        P.stats[0] = status; P.stats[1] = phases; P.stats[2] = rounds; P.stats[3] = st_acc[0];
        P.stats[4] = phases - 1; P.stats[5] = cmin; P.stats[6] = cmax; P.stats[7] = S;
        P.stats[8] = G; P.stats[9] = SMEMP ? 1 : 0; P.stats[10] = 2 + (SMEMO ? 1 : 0); P.stats[11] = st_acc[1];
        P.stats[12] = (phases - 1) * (long long)np; P.stats[14] = searches; P.stats[15] = srounds;
        P.stats[16] = st_acc[3]; P.stats[17] = st_acc[4]; P.stats[18] = st_acc[5]; P.stats[19] = st_acc[6]; P.stats[20] = st_acc[7];
        P.stats[21] = paths; P.stats[27] = P.warm; P.stats[22] = st_acc[9]; P.stats[23] = st_acc[10]; P.stats[24] = st_acc[11];
        P.stats[25] = st_acc[2]; P.stats[26] = st_acc[8];
        P.stats[28] = st_ph[2]; P.stats[29] = st_ph[3]; P.stats[30] = st_ph[4]; P.stats[31] = global_ns() - st_ph[0];
    }
}

struct SapLayout {
    size_t list[3], rec[3], bidw[3], flag, slot_price, person_slot, minslot, slot_obj, dkey, gsnap, chgbits[3], cand[3], claim, moves, anc, small, total;
};

SapLayout sap_layout(int64_t np, int64_t no) {
    SapLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = cyb::align_up(o + bytes, 256); return r; };
    for (int k = 0; k < 3; ++k) L.list[k] = take((size_t)np * 4);
    for (int k = 0; k < 3; ++k) L.rec[k] = take((size_t)np * 16);
    for (int k = 0; k < 3; ++k) L.bidw[k] = take((size_t)no * 8);
    L.flag = take((size_t)np * 4);
    L.slot_price = take((size_t)np * 8);
    L.person_slot = take((size_t)np * 4);
    L.minslot = take((size_t)no * 4);
    L.slot_obj = take((size_t)np * 4);
    L.dkey = take((size_t)no * 8);
    L.gsnap = take((size_t)no * 8);
    for (int k = 0; k < 3; ++k) L.chgbits[k] = take((size_t)((no + 31) / 32) * 4);
    for (int k = 0; k < 3; ++k) L.cand[k] = take((size_t)no * 8);
    L.claim = take((size_t)(no + kSapMax) * 4);
    L.moves = take((size_t)np * 16);
    L.anc = take((size_t)kMaxGrid * no * 4);
    L.small = take(2048);
    L.total = o;
    return L;
}

}  // namespace

extern "C" size_t cyb_lap_workspace_bytes(int64_t n_persons, int64_t n_objects) {
    if (n_persons <= 0 || n_objects <= 0) return 256;
    return std::max(sap_layout(n_persons, n_objects).total, cyb::lap_check_workspace_bytes(n_persons, n_objects));
}

extern "C" int cyb_lap_solve_i32(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                                 const int32_t *slot_offset_dev, int32_t *person_obj_dev,
                                 int32_t *slot_owner_dev, int64_t *price_dev, int64_t *total_dev,
                                 int64_t *stats_dev, void *workspace_dev, size_t workspace_bytes,
                                 int grid_hint, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int64_t np = n_persons, no = n_objects;
    // (more objects than persons is legal with capacities: spots that take no cell are priced out)
    if (np <= 0 || np >= (1ll << kPB) - kSapMax || no <= 0 || no >= (1ll << kPB) - kSapMax)
        return cyb::set_error(CYB_ERR_INVALID,
                              "cyb_lap_solve_i32: persons=%lld objects=%lld outside 1 <= persons, objects < 2^18 - 256",
                              (long long)np, (long long)no);
    if (!slot_offset_dev && no != np)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: without capacities the problem must be square");
    if (!cost_dev || !person_obj_dev || !slot_owner_dev || !price_dev || !total_dev || !stats_dev || !workspace_dev)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: null pointer argument");
    if (ld < no)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_lap_solve_i32: ld=%lld < objects=%lld", (long long)ld, (long long)no);
    const SapLayout L = sap_layout(np, no);
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
    if (G > kMaxGrid) G = kMaxGrid;
    if (G > np) G = (int)np;
    if (G < 1) G = 1;

    char *ws = static_cast<char *>(workspace_dev);
    SapParams P;
    memset(&P, 0, sizeof(P));
    P.cost = cost_dev; P.ld = ld; P.P = (int)np; P.O = (int)no; P.soff = slot_offset_dev;
    P.person_obj = person_obj_dev; P.slot_owner = slot_owner_dev;
    P.lambda = reinterpret_cast<long long *>(price_dev);
    P.total = reinterpret_cast<long long *>(total_dev); P.stats = reinterpret_cast<long long *>(stats_dev);
    for (int k = 0; k < 3; ++k) {
        P.list[k] = reinterpret_cast<int32_t *>(ws + L.list[k]);
        P.rec[k] = reinterpret_cast<int4 *>(ws + L.rec[k]);
        P.bidw[k] = reinterpret_cast<unsigned long long *>(ws + L.bidw[k]);
        P.chgbits[k] = reinterpret_cast<unsigned *>(ws + L.chgbits[k]);
        P.cand[k] = reinterpret_cast<unsigned long long *>(ws + L.cand[k]);
    }
    P.flag = reinterpret_cast<int32_t *>(ws + L.flag);
    P.slot_price = reinterpret_cast<long long *>(ws + L.slot_price);
    P.person_slot = reinterpret_cast<int32_t *>(ws + L.person_slot);
    P.minslot = reinterpret_cast<int32_t *>(ws + L.minslot);
    P.slot_obj = reinterpret_cast<int32_t *>(ws + L.slot_obj);
    P.dkey = reinterpret_cast<unsigned long long *>(ws + L.dkey);
    P.gsnap = reinterpret_cast<long long *>(ws + L.gsnap);
    P.claim = reinterpret_cast<int32_t *>(ws + L.claim);
    P.moves = reinterpret_cast<int4 *>(ws + L.moves);
    P.anc = reinterpret_cast<int32_t *>(ws + L.anc);
    P.bar = reinterpret_cast<unsigned int *>(ws + L.small);
    P.gmm = reinterpret_cast<int *>(ws + L.small + 16);
    P.nmoves = reinterpret_cast<int *>(ws + L.small + 96);
    P.rstat = reinterpret_cast<int *>(ws + L.small + 1280);              // 3 x 32 bytes
    P.srcdone = reinterpret_cast<int32_t *>(ws + L.small + 128);          // kSapMax ints = 1 KB
    P.qcap = (int)((std::max(np, no) + G - 1) / G);
    P.max_rounds = 2000ll * np + 100000;
    P.prefetch = no * 4 > 131072 ? 1 : 0;
    if (const char *e = getenv("CYB_LAP_PREFETCH")) P.prefetch = atoi(e) ? 1 : 0;
    P.packed_reduce = 1;
    if (const char *e = getenv("CYB_LAP_PACKED")) P.packed_reduce = atoi(e) ? 1 : 0;
    P.theta = 8; P.eps0_div = 4;
    if (const char *e = getenv("CYB_LAP_THETA")) P.theta = std::max(2, atoi(e));
    if (const char *e = getenv("CYB_LAP_EPS0")) P.eps0_div = std::max(1, atoi(e));
    // (constants, not functions of the grid: the assignment must not depend on the grid size)
    // measured on B200 (profiles/r02_lap_knob_sweep.txt): 64 / 296 / 16 is the best single setting over 10k x 10k,
    // 30k x 5k and 25k x 25k (auction rounds cost ~7 us, search rounds ~19 us: the search should start late)
    P.sap_t = 64; P.sap_k = 296; P.multi = 16;
    if (const char *e = getenv("CYB_LAP_SAP_T")) P.sap_t = std::max(1, std::min(kSapMax, atoi(e)));
    if (const char *e = getenv("CYB_LAP_SAP_K")) P.sap_k = std::max(1, atoi(e));
    if (const char *e = getenv("CYB_LAP_SAP_MULTI")) P.multi = std::max(1, std::min(kMultiMax, atoi(e)));
    // Incomplete phases (oracle/sap_model.c `sap_partial`): a phase with eps > 1 stops when <= 64 persons are
    // free instead of placing them with searches whose result the next phase start throws away again; with that, a gentler
    // eps schedule (theta 8 instead of 64) pays.  Measured: cfg2 45.9 -> 23.5 ms, 30k x 5k 100 -> 54, 25k 172 -> 102,
    // 50k 498 -> 283 (search rounds at cfg2: 1 498 -> 721, searches 46 -> 8).
    P.partial = 64;
    if (const char *e = getenv("CYB_LAP_PARTIAL")) P.partial = std::max(0, atoi(e));
    // Warm-started searches (kept forest + repair round, only where the predecessors live in shared memory) cut the search
    // rounds by a quarter (cfg2: 1 984 -> 1 498).  Measured with the pointer-jumping kept / dropped pass: 4k x 4k
    // 19.3 -> 16.8 ms, cfg2 48.4 -> 45.9 ms, 30k x 5k 103 -> 102 ms.
    P.warm = 1;
    if (const char *e = getenv("CYB_LAP_WARM")) P.warm = atoi(e) ? 1 : 0;

    // barrier counters [0..1], gmm = {cmin, cmax, status, -, -, abort flag}
    const int init[16] = {0, 0, 0, 0, INT_MAX, INT_MIN, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    CYB_CUDA_CHECK(cudaMemcpyAsync(ws + L.small, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    CYB_CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), stream));
    CYB_CUDA_CHECK(cudaMemsetAsync(stats_dev, 0, sizeof(int64_t) * CYB_LAP_NSTATS, stream));

    P.wpc = (int)(((no + 31) / 32 + G - 1) / G);
    const size_t base_bytes = cyb::align_up((size_t)P.qcap * 8, 16) + cyb::align_up((size_t)P.qcap * 4, 16) +
                              5 * cyb::align_up((size_t)((no + 31) / 32) * 4, 16) + (size_t)P.wpc * 32 * 8 +
                              2 * cyb::align_up((size_t)P.wpc * 4, 16);
    const size_t price_bytes = cyb::align_up((size_t)no * 8, 16);
    const size_t owner_bytes = (size_t)((np + 3) & ~3) * 4 + (size_t)((no + 3) & ~3) * 4 +
                               (slot_offset_dev ? (size_t)((no + 3) & ~3) * 4 : 0) + 16;
    const size_t static_smem = 8192;              // bound on the kernel's static shared memory
    bool smemp = base_bytes + price_bytes + static_smem <= (size_t)max_smem;
    if (const char *e = getenv("CYB_LAP_SMEM_PRICES")) smemp = smemp && atoi(e);
    size_t dyn = base_bytes + (smemp ? price_bytes : 0);
    bool smemo = smemp && dyn + owner_bytes + static_smem <= (size_t)max_smem;
    if (const char *e = getenv("CYB_LAP_SMEM_OWNER")) smemo = smemo && atoi(e);
    if (smemo) dyn += owner_bytes;
    if (!smemo) P.warm = 0;
    P.max_teams = smemp ? 8 : 1;
    if (const char *e = getenv("CYB_LAP_TEAMS")) P.max_teams = std::max(1, std::min(8, atoi(e)));

    const void *fn = smemp ? (smemo ? (const void *)lap_sap_kernel<true, true> : (const void *)lap_sap_kernel<true, false>)
                           : (const void *)lap_sap_kernel<false, false>;
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
