/* TEST INFRASTRUCTURE ONLY -- sequential CPU model of the synchronous (Jacobi)
 * eps-scaling auction with CAPACITATED objects that the device LAP solver
 * (cytospace_b200/csrc/lap_auction.cu) runs.  It exists so the device algorithm's
 * exactness argument, tie-breaks and round/bid counts can be checked without a
 * GPU; it is never linked into the product library.
 *
 * Problem (transportation form of CytoSPACE's expanded LAP,
 * linear_assignment_solvers.py:63-66): persons i = 0..P-1 (cells), objects
 * o = 0..O-1 (spots) with capacity cap[o], sum cap = P;
 *     min sum_i m[i, obj(i)]   s.t. object o holds exactly cap[o] persons.
 * m is persons x objects, row-major (the transpose of the reference's cost).
 *
 * Every object owns cap[o] SLOTS, each with its own price (the last accepted bid)
 * and holder; the object's price lambda[o] is the minimum slot price.  A free
 * person scans its row: v1 = min_o (M[i,o] + lambda[o]) at o* (lowest o on ties),
 * w = min over o != o*; it bids  b = lambda[o*] + (w - v1) + eps  for the cheapest
 * slot of o* (lowest slot index on ties).  Per object the highest bid of a round
 * wins (lowest person on equal bids), the slot's previous holder becomes free.
 * "Similar objects" never fight each other: the second-best value excludes the
 * sibling slots of o*, so duplicated spot rows cause no price war.
 *
 * Exactness: costs are scaled by (P+1), prices are int64, the last phase runs
 * with eps = 1.  eps-CS w.r.t. lambda then bounds the scaled total within P*eps
 * < P+1 of optimal, i.e. the unscaled total is optimal.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AUCTION_INF ((int64_t)1 << 60)
/* candidate-list experiment (tail only): groups of columns as the device scan assigns them to threads */
int64_t auction_list_hits = 0, auction_list_miss = 0;
int auction_list_groups = 0;       /* 0 = lists off; else number of groups (32 = warps, 1024 = threads) */
int auction_list_cap = 64;
int auction_early_stop = 0;     /* >0: a phase with eps > 1 ends as soon as <= this many persons are free; they carry over */
int64_t auction_phase_log[64][4];   /* per phase: free at start, rounds, bids, tail bids */
int auction_list_refresh = 0;   /* >0: every R tail bids all persons' lists are rebuilt at current prices (idle-CTA sweep model) */
#define LIST_MAX 1024
typedef struct { int n; int64_t bound; int32_t obj[LIST_MAX]; } cand_t;
#define GAP(b1, b2) (((b2) >= AUCTION_INF / 2) ? 0 : (b2) - (b1))   /* no alternative object: bid eps */

/* stats[0]=phases, [1]=rounds, [2]=bids (row scans), [3]=full-matrix passes,
 * [4]=rounds with <=148 bidders, [5]=bids made in the Gauss-Seidel tail.
 * round_log (may be NULL, capacity round_cap): bidders per round. */
int auction_model_i32(int P, int O, const int32_t *m, int64_t ld, const int32_t *cap,
                      int32_t *person_obj, int32_t *slot_owner, int64_t *lambda,
                      int64_t *total, int64_t theta, int64_t eps0_div, int64_t tail_t,
                      int64_t *stats, int32_t *round_log, int64_t round_cap, int variant)
{
    if (P <= 0 || O <= 0) { if (total) *total = 0; return (P == 0) ? 0 : -1; }
    const int64_t S = (int64_t)P + 1;
    int32_t *soff = (int32_t *)malloc(sizeof(int32_t) * ((size_t)O + 1));
    soff[0] = 0;
    for (int o = 0; o < O; ++o) soff[o + 1] = soff[o] + (cap ? cap[o] : 1);
    if (soff[O] != P) { free(soff); return -3; }
    int32_t cmin = INT32_MAX, cmax = INT32_MIN;
    for (int i = 0; i < P; ++i) {
        const int32_t *r = m + (size_t)i * ld;
        for (int o = 0; o < O; ++o) { if (r[o] < cmin) cmin = r[o]; if (r[o] > cmax) cmax = r[o]; }
    }
    int64_t *slot_price = (int64_t *)calloc((size_t)P, sizeof(int64_t));
    int32_t *minslot = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *person_slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int32_t *freel = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int32_t *nextl = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int64_t *bidp = (int64_t *)malloc(sizeof(int64_t) * (size_t)O);
    int32_t *bidr = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *kobj = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    cand_t *lists = auction_list_groups ? (cand_t *)calloc((size_t)P, sizeof(cand_t)) : NULL;
    int64_t gb1[1024], gb2[1024]; int gj[1024];
    /* an object without capacity (a spot that takes no cell, cytospace.py:686-694) is priced out */
    for (int o = 0; o < O; ++o) { lambda[o] = (soff[o + 1] > soff[o]) ? 0 : AUCTION_INF; minslot[o] = soff[o]; bidr[o] = -1; }
    for (int t = 0; t < P; ++t) slot_owner[t] = -1;
    for (int i = 0; i < P; ++i) { person_obj[i] = -1; person_slot[i] = -1; }
    int64_t st[6] = {0, 0, 0, 1, 0, 0};

#define SCAN(i, b1, b2, o1)                                                        \
    do {                                                                           \
        const int32_t *r_ = m + (size_t)(i) * ld;                                  \
        b1 = INT64_MAX; b2 = INT64_MAX; o1 = -1;                                   \
        for (int o_ = 0; o_ < O; ++o_) {                                           \
            int64_t h_ = (int64_t)(r_[o_] - cmin) * S + lambda[o_];                \
            if (h_ < b1) { b2 = b1; b1 = h_; o1 = o_; }                            \
            else if (h_ < b2) b2 = h_;                                             \
        }                                                                          \
    } while (0)
#define REFRESH(o)                                                                 \
    do {                                                                           \
        int ms_ = soff[o]; int64_t mp_ = slot_price[ms_];                          \
        for (int t_ = soff[o] + 1; t_ < soff[(o) + 1]; ++t_)                       \
            if (slot_price[t_] < mp_) { mp_ = slot_price[t_]; ms_ = t_; }          \
        minslot[o] = ms_; lambda[o] = mp_;                                         \
    } while (0)

    int64_t eps = ((int64_t)cmax - (int64_t)cmin) * S / (eps0_div > 0 ? eps0_div : 4);
    if (eps < 1) eps = 1;
    for (;;) {
        ++st[0];
        int nfree = 0;
        if (st[0] == 1) {
            for (int i = 0; i < P; ++i) freel[nfree++] = i;
        } else {
            /* keep the pairs that already satisfy eps-CS at the new eps.  The test uses the person's
             * OWN slot price (>= lambda of its object, which may still rise up to it) against the best
             * alternative object, so the pair stays eps-CS for the rest of the phase.  A vacated slot
             * keeps its price. */
            ++st[3];
            for (int i = 0; i < P; ++i) {
                int64_t b1, b2; int o1;
                SCAN(i, b1, b2, o1);
                const int o = person_obj[i];
                int drop = 1;
                if (o >= 0) {
                    const int64_t alt = (o1 == o) ? b2 : b1;
                    const int64_t base = (int64_t)(m[(size_t)i * ld + o] - cmin) * S;
                    if (alt >= AUCTION_INF / 2) drop = 0;
                    else if (variant == 0) drop = base + slot_price[person_slot[i]] > alt + eps;
                    else {
                        /* test against the OBJECT price, then clamp the own slot price down to the
                         * highest level that is still eps-CS (>= lambda[o], so lambda never drops) */
                        drop = base + lambda[o] > alt + eps;
                        if (!drop && base + slot_price[person_slot[i]] > alt + eps) slot_price[person_slot[i]] = alt + eps - base;
                    }
                }
                if (drop) {
                    if (o >= 0) slot_owner[person_slot[i]] = -1;
                    person_obj[i] = -1; person_slot[i] = -1; freel[nfree++] = i;
                }
            }
        }
        if (st[0] > 1 && variant) for (int o = 0; o < O; ++o) if (soff[o + 1] > soff[o]) REFRESH(o);
        const int64_t ph_r0 = st[1], ph_b0 = st[2], ph_t0 = st[5];
        if (st[0] <= 64) auction_phase_log[st[0] - 1][0] = nfree;
        while (nfree > 0) {
            if (eps > 1 && nfree <= auction_early_stop) break;    /* the free persons bid again in the next phase */
            if (nfree <= tail_t) {
                /* Gauss-Seidel tail (what CTA 0 runs alone on the device): FIFO of free persons */
                int head = 0, tailp = nfree % P, cnt = nfree;
                while (cnt > 0) {
                    const int i = freel[head]; head = (head + 1) % P; --cnt;
                    int64_t b1, b2; int o1;
                    int done = 0;
                    if (auction_list_groups > 0 && lists && auction_list_refresh > 0 && (st[5] % auction_list_refresh) == 0) {
                        const int Gn = auction_list_groups;
                        for (int p_ = 0; p_ < P; ++p_) {
                            const int32_t *r_ = m + (size_t)p_ * ld;
                            for (int g = 0; g < Gn; ++g) { gb1[g] = INT64_MAX; gb2[g] = INT64_MAX; gj[g] = -1; }
                            for (int o_ = 0; o_ < O; ++o_) {
                                const int g = (Gn == 32) ? (((o_ >> 2) % 1024) >> 5) : ((o_ >> 2) % Gn);
                                const int64_t h_ = (int64_t)(r_[o_] - cmin) * S + lambda[o_];
                                if (h_ < gb1[g]) { gb2[g] = gb1[g]; gb1[g] = h_; gj[g] = o_; }
                                else if (h_ < gb2[g]) gb2[g] = h_;
                            }
                            int64_t bound = INT64_MAX;
                            for (int g = 0; g < Gn; ++g) if (gb2[g] < bound) bound = gb2[g];
                            int cntq = 0;
                            for (int g = 0; g < Gn; ++g) if (gj[g] >= 0 && gb1[g] < bound) ++cntq;
                            lists[p_].n = 0; lists[p_].bound = bound;
                            if (cntq <= auction_list_cap && cntq <= LIST_MAX)
                                for (int g = 0; g < Gn; ++g) if (gj[g] >= 0 && gb1[g] < bound) lists[p_].obj[lists[p_].n++] = gj[g];
                        }
                    }
                    if (auction_list_groups && lists && lists[i].n > 0) {
                        const int32_t *r_ = m + (size_t)i * ld;
                        b1 = INT64_MAX; b2 = INT64_MAX; o1 = -1;
                        for (int q = 0; q < lists[i].n; ++q) {
                            const int o_ = lists[i].obj[q];
                            const int64_t h_ = (int64_t)(r_[o_] - cmin) * S + lambda[o_];
                            if (h_ < b1 || (h_ == b1 && o_ < o1)) { b2 = b1; b1 = h_; o1 = o_; }
                            else if (h_ < b2) b2 = h_;
                        }
                        if (b2 < lists[i].bound) { done = 1; ++auction_list_hits; }
                    }
                    if (!done) {
                        SCAN(i, b1, b2, o1);
                        if (auction_list_groups && lists) {
                            ++auction_list_miss;
                            const int32_t *r_ = m + (size_t)i * ld;
                            if (auction_list_groups < 0) {
                                /* exact top-K (K = -groups <= LIST_MAX): bound = (K+1)-th smallest value */
                                const int K = -auction_list_groups;
                                int64_t hv[LIST_MAX + 1]; int ho[LIST_MAX + 1]; int nn = 0;
                                for (int o_ = 0; o_ < O; ++o_) {
                                    const int64_t h_ = (int64_t)(r_[o_] - cmin) * S + lambda[o_];
                                    if (nn <= K || h_ < hv[nn - 1]) {
                                        int p_ = nn < K + 1 ? nn++ : nn - 1;
                                        while (p_ > 0 && hv[p_ - 1] > h_) { hv[p_] = hv[p_ - 1]; ho[p_] = ho[p_ - 1]; --p_; }
                                        hv[p_] = h_; ho[p_] = o_;
                                    }
                                }
                                lists[i].n = 0; lists[i].bound = nn > K ? hv[K] : INT64_MAX;
                                for (int q = 0; q < (nn > K ? K : nn); ++q) lists[i].obj[lists[i].n++] = ho[q];
                            } else {
                            /* rebuild: per-group best / second best */
                            const int Gn = auction_list_groups;
                            for (int g = 0; g < Gn; ++g) { gb1[g] = INT64_MAX; gb2[g] = INT64_MAX; gj[g] = -1; }
                            for (int o_ = 0; o_ < O; ++o_) {
                                const int g = (Gn == 32) ? (((o_ >> 2) % 1024) >> 5) : ((o_ >> 2) % Gn);
                                const int64_t h_ = (int64_t)(r_[o_] - cmin) * S + lambda[o_];
                                if (h_ < gb1[g]) { gb2[g] = gb1[g]; gb1[g] = h_; gj[g] = o_; }
                                else if (h_ < gb2[g]) gb2[g] = h_;
                            }
                            int64_t bound = INT64_MAX;
                            for (int g = 0; g < Gn; ++g) if (gb2[g] < bound) bound = gb2[g];
                            int cntq = 0;
                            for (int g = 0; g < Gn; ++g) if (gj[g] >= 0 && gb1[g] < bound) ++cntq;
                            lists[i].n = 0; lists[i].bound = bound;
                            if (cntq <= auction_list_cap && cntq <= LIST_MAX)
                                for (int g = 0; g < Gn; ++g) if (gj[g] >= 0 && gb1[g] < bound) lists[i].obj[lists[i].n++] = gj[g];
                            }
                        }
                    }
                    const int64_t bid = lambda[o1] + GAP(b1, b2) + eps;
                    const int t = minslot[o1];
                    const int old = slot_owner[t];
                    slot_owner[t] = i; slot_price[t] = bid; person_obj[i] = o1; person_slot[i] = t;
                    if (old >= 0) {
                        person_obj[old] = -1; person_slot[old] = -1;
                        freel[tailp] = old; tailp = (tailp + 1) % P; ++cnt;
                    }
                    REFRESH(o1);
                    ++st[2]; ++st[5];
                }
                nfree = 0;
                break;
            }
            if (round_log && st[1] < round_cap) round_log[st[1]] = nfree;
            ++st[1]; st[2] += nfree;
            if (nfree <= 148) ++st[4];
            /* bids: all computed against the prices of the round start (Jacobi) */
            for (int k = 0; k < nfree; ++k) {
                const int i = freel[k];
                int64_t b1, b2; int o1;
                SCAN(i, b1, b2, o1);
                const int64_t bp = lambda[o1] + GAP(b1, b2) + eps;
                kobj[k] = o1;
                if (bidr[o1] < 0 || bp > bidp[o1] || (bp == bidp[o1] && i < bidr[o1])) { bidp[o1] = bp; bidr[o1] = i; }
            }
            /* resolve in record order (the device replays the records in this order): the winner of
             * an object takes its cheapest slot at its bid, the slot's previous holder becomes free;
             * a loser stays free.  The next free list keeps the record order. */
            int nnext = 0;
            for (int k = 0; k < nfree; ++k) {
                const int i = freel[k], o = kobj[k];
                if (bidr[o] == i) {
                    const int t = minslot[o], old = slot_owner[t];
                    slot_owner[t] = i; slot_price[t] = bidp[o]; person_obj[i] = o; person_slot[i] = t;
                    if (old >= 0) { person_obj[old] = -1; person_slot[old] = -1; nextl[nnext++] = old; }
                } else {
                    nextl[nnext++] = i;
                }
            }
            for (int k = 0; k < nfree; ++k) {
                const int o = kobj[k];
                if (bidr[o] >= 0) { REFRESH(o); bidr[o] = -1; }
            }
            int32_t *tmp = freel; freel = nextl; nextl = tmp; nfree = nnext;
        }
        if (st[0] <= 64) {
            auction_phase_log[st[0] - 1][1] = st[1] - ph_r0; auction_phase_log[st[0] - 1][2] = st[2] - ph_b0;
            auction_phase_log[st[0] - 1][3] = st[5] - ph_t0;
        }
        if (eps == 1) break;
        eps = eps / theta; if (eps < 1) eps = 1;
    }
    int64_t tot = 0;
    for (int i = 0; i < P; ++i) tot += m[(size_t)i * ld + person_obj[i]];
    if (total) *total = tot;
    if (stats) memcpy(stats, st, sizeof(st));
    free(soff); free(slot_price); free(minslot); free(person_slot); free(freel); free(nextl);
    free(bidp); free(bidr); free(kobj); free(lists);
    return 0;
}
