/* TEST INFRASTRUCTURE ONLY -- sequential CPU model of the HYBRID device LAP solver:
 * synchronous (Jacobi) eps-scaling auction rounds for the bulk of every phase (as in
 * auction_model.c) and, once few persons are free, a multi-source SHORTEST AUGMENTING PATH
 * finish whose search is a label-correcting (delta-stepping) relaxation that the device runs
 * level-synchronously over the whole grid.  Used (i) as the CPU experiment that sized the device
 * design -- dependent rounds and row scans per tail against the Gauss-Seidel tail's dependent
 * bids -- and (ii) as the reference the device kernel's assignment is compared with.
 *
 * Same problem and state as auction_model.c (transportation form of CytoSPACE's expanded LAP,
 * linear_assignment_solvers.py:63-66; the solve the reference reaches at
 * linear_assignment_solvers.py:38).  Cs = (m - cmin) * (P+1).
 *
 * The SAP finish keeps the auction's invariant exactly.  Residual graph at the current eps:
 *   person i (assigned to o_i) -> object k != o_i :  len = Cs[i,k] + lambda[k] + eps - (Cs[i,o_i] + lambda[o_i])  >= 0  (eps-CS)
 *   free person i              -> object k        :  len = Cs[i,k] + lambda[k] - min_k'(Cs[i,k'] + lambda[k'])    >= 0
 *   object o -> every person it holds             :  len = 0
 * Labels d[o] are relaxed from all free persons at once; the search ends when no object with
 * d[o] < D still has to be (re)scanned, D = the smallest label of an object with a free slot
 * (multi-path mode: the D at which `want` source trees have reached distinct free objects).
 * Then lambda[o] += D - d[o] for d[o] < D (prices only rise; every relaxed arc keeps len >= 0 and
 * the tree arcs become tight), and the assignment is flipped along the tree path(s): the new pair
 * (i, o) satisfies Cs[i,o] + lambda[o] + eps = (old value of i) i.e. eps-CS with slack eps.  Slot
 * prices of touched objects become max(old, lambda[o]) and lambda[o] for the slots taken.
 * Because a fixed point of the relaxation below D is all the proof needs, labels may be relaxed
 * in any order: each round takes the `K` smallest labelled dirty objects' holders (rows) at once.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SAP_INF ((int64_t)1 << 60)
#define GAP(b1, b2) (((b2) >= SAP_INF / 2) ? 0 : (b2) - (b1))

/* per phase: [0] free at phase start, [1] auction rounds, [2] auction bids, [3] searches,
 * [4] search rounds (dependent grid rounds), [5] rows scanned in searches, [6] paths applied,
 * [7] free persons when the SAP finish took over */
int64_t sap_phase_log[64][8];
/* 1: a search that follows another one in the same phase starts from the surviving part of its shortest-path forest
 * (labels shifted by D, trees of the persons that got assigned dropped) instead of from scratch */
int sap_warm = 0;
/* > 0: CHAINED relaxations.  When the rows of a frontier object lower the label of an object k to nd <= T (inside this
 * round's window, below D) the best such discovery (smallest (nd, k); k must have a holder) is relaxed in the SAME round
 * with the label nd, up to sap_chain hops deep -- a round then advances the search by up to 1 + sap_chain arcs instead of
 * one (the searches are depth-bound: ~90 rounds each at 10k).  Everything is judged against the round-start snapshot, so
 * the result does not depend on the order of the relaxations; an object whose chained label is its final label of the
 * round is not dirty afterwards.  Objects with more than SAP_CHAIN_CAP slots do not start chains (the device relaxes at
 * most that many rows per pass); the repair round of a warm search does not chain. */
int sap_chain = 0;
#define SAP_CHAIN_CAP 32
int sap_partial = 0;      /* > 0: a phase with eps > 1 ends as soon as <= sap_partial persons are free (they stay free into the next phase) */

typedef struct { int64_t d; int32_t o; } lab_t;
static int lab_cmp(const void *a, const void *b) {
    const lab_t *x = (const lab_t *)a, *y = (const lab_t *)b;
    if (x->d != y->d) return x->d < y->d ? -1 : 1;
    return x->o < y->o ? -1 : (x->o > y->o);
}

/* stats: [0] phases, [1] auction rounds, [2] auction bids, [3] searches, [4] search rounds,
 * [5] search row scans, [6] paths, [7] phase-start passes.
 * sap_t: the SAP finish takes over when <= sap_t persons are free (0: never -> pure auction).
 * K: rows relaxed per search round (target; whole objects are taken).
 * multi: 0 = one augmenting path per search, m > 0 = a search continues until min(free, m)
 * source trees have reached distinct free objects (all of them applied after one price update). */
int sap_model_i32(int P, int O, const int32_t *m, int64_t ld, const int32_t *cap,
                  int32_t *person_obj, int32_t *slot_owner, int64_t *lambda, int64_t *total,
                  int64_t theta, int64_t eps0_div, int64_t sap_t, int64_t K, int64_t multi, int64_t *stats)
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
    /* search state */
    int64_t *d = (int64_t *)malloc(sizeof(int64_t) * (size_t)O);
    int64_t *snap = (int64_t *)malloc(sizeof(int64_t) * (size_t)O);
    int64_t *chainlab = (int64_t *)malloc(sizeof(int64_t) * (size_t)O);     /* smallest label an object was chain-relaxed with this round */
    int32_t *cq_o = (int32_t *)malloc(sizeof(int32_t) * (size_t)O * 4 + 64); int64_t *cq_d = (int64_t *)malloc(sizeof(int64_t) * (size_t)O * 4 + 64);
    int32_t *cq_h = (int32_t *)malloc(sizeof(int32_t) * (size_t)O * 4 + 64);
    int32_t *pred = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *nfreeslot = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    lab_t *cand = (lab_t *)malloc(sizeof(lab_t) * (size_t)O);
    int32_t *fo = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *dirty = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *front = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *claim = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *claim_src = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int32_t *slot_obj = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    char *indirty = (char *)malloc((size_t)O);
    char *touched = (char *)malloc((size_t)O);
    char *onpath = (char *)malloc((size_t)O);
    char *vstate = (char *)malloc((size_t)O);
    int32_t *chain = (int32_t *)malloc(sizeof(int32_t) * (size_t)O);
    int32_t *srcmap = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int have_forest = 0;
    for (int o = 0; o < O; ++o) for (int t = soff[o]; t < soff[o + 1]; ++t) slot_obj[t] = o;
    for (int o = 0; o < O; ++o) { lambda[o] = (soff[o + 1] > soff[o]) ? 0 : SAP_INF; minslot[o] = soff[o]; bidr[o] = -1; }
    for (int t = 0; t < P; ++t) slot_owner[t] = -1;
    for (int i = 0; i < P; ++i) { person_obj[i] = -1; person_slot[i] = -1; }
    int64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 1};
    int rc = 0;

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
            ++st[7];
            for (int i = 0; i < P; ++i) {
                int64_t b1, b2; int o1;
                SCAN(i, b1, b2, o1);
                const int o = person_obj[i];
                int drop = 1;
                if (o >= 0) {
                    const int64_t alt = (o1 == o) ? b2 : b1;
                    const int64_t base = (int64_t)(m[(size_t)i * ld + o] - cmin) * S;
                    if (alt >= SAP_INF / 2) drop = 0;
                    else {
                        drop = base + lambda[o] > alt + eps;
                        if (!drop && base + slot_price[person_slot[i]] > alt + eps) slot_price[person_slot[i]] = alt + eps - base;
                    }
                }
                if (drop) {
                    if (o >= 0) slot_owner[person_slot[i]] = -1;
                    person_obj[i] = -1; person_slot[i] = -1; freel[nfree++] = i;
                }
            }
            for (int o = 0; o < O; ++o) if (soff[o + 1] > soff[o]) REFRESH(o);
        }
        const int ph = (int)st[0] - 1;
        have_forest = 0;
        int64_t step = eps;                     /* frontier window of the searches, adapted round by round */
        int64_t ph0[8]; memcpy(ph0, st, sizeof(st));
        if (ph < 64) { memset(sap_phase_log[ph], 0, sizeof(sap_phase_log[ph])); sap_phase_log[ph][0] = nfree; }
        while (nfree > 0) {
            if (eps > 1 && nfree <= sap_partial) break;
            if (nfree <= sap_t) {
                if (ph < 64 && sap_phase_log[ph][7] == 0) sap_phase_log[ph][7] = nfree;
                /* ================= SAP finish: one search, one price update, >= 1 augmentation ============ */
                ++st[3];
                int nfo = 0;
                const int warm_now = sap_warm && have_forest;
                for (int o = 0; o < O; ++o) {
                    if (!warm_now) { d[o] = SAP_INF; pred[o] = -1; }
                    nfreeslot[o] = 0; indirty[o] = 0;
                    for (int t = soff[o]; t < soff[o + 1]; ++t) if (slot_owner[t] < 0) ++nfreeslot[o];
                    if (nfreeslot[o] > 0) fo[nfo++] = o;
                }
                /* round 0: the free persons relax their rows (values relative to the row minimum); source k is
                 * recorded as pred = P + k */
                ++st[4];
                for (int k = 0; k < nfree; ++k) {
                    const int i = freel[k];
                    int64_t b1, b2; int o1;
                    SCAN(i, b1, b2, o1);
                    (void)b2; (void)o1;
                    const int32_t *r = m + (size_t)i * ld;
                    for (int o = 0; o < O; ++o) {
                        if (lambda[o] >= SAP_INF / 2) continue;
                        const int64_t nd = (int64_t)(r[o] - cmin) * S + lambda[o] - b1;
                        if (nd < d[o] || (nd == d[o] && P + k < pred[o])) { d[o] = nd; pred[o] = P + k; }
                    }
                    st[5] += 2;
                }
                int ndirty = 0;
                for (int o = 0; o < O; ++o) if (d[o] < SAP_INF / 2) { dirty[ndirty++] = o; indirty[o] = 1; }
                int want = nfree < nfo ? nfree : nfo;
                if (multi <= 0) want = 1; else if (want > multi) want = (int)multi;
                int64_t D = SAP_INF, T = -1, Tg = step, Tnext = -1;         /* round 0 leaves labels 0: the first guess is one window */
                int repair = warm_now;                                      /* first round of a warm search: every surviving tree node relaxes again */
                for (;;) {
                    /* D = the want-th smallest label of an object with a free slot */
                    int nc = 0;
                    for (int q = 0; q < nfo; ++q) if (d[fo[q]] < SAP_INF / 2) { cand[nc].d = d[fo[q]]; cand[nc].o = fo[q]; ++nc; }
                    qsort(cand, (size_t)nc, sizeof(lab_t), lab_cmp);
                    D = nc >= want ? cand[want - 1].d : SAP_INF;
                    if (D >= SAP_INF / 2) { rc = -6; goto done; }           /* every object with capacity is reached in round 0 */
                    if (step > D) step = D;
                    if (step < 1) step = 1;
                    /* dirty objects that still matter (eligible): label < D and somebody to scan */
                    int nd2 = 0;
                    for (int q = 0; q < ndirty; ++q) {
                        const int o = dirty[q];
                        if (d[o] < D && (soff[o + 1] - soff[o]) - nfreeslot[o] > 0) dirty[nd2++] = o;
                        else indirty[o] = 0;
                    }
                    ndirty = nd2;
                    if (ndirty == 0) break;
                    /* Frontier = about K rows' worth of the smallest eligible labels, found in ONE pass over the
                     * labels: candidates are the eligible objects with label <= Tg (a guess fixed at the end of the
                     * previous round: its threshold plus `step`), histogrammed in 256 power-of-two bins over
                     * [smallest eligible label, Tg]; T = upper edge of the first bin where the cumulative slot count reaches K (Tg when
                     * the candidates hold fewer).  `step` doubles when the window held fewer than K although more was eligible, halves
                     * above 4K (state kept across the searches of a phase).  When the guess selects nothing (first
                     * round of a search, D moved below it) it restarts from the smallest eligible label. */
                    int64_t wC = 0, wE = 0, dmin_el = SAP_INF;
                    if (repair) Tg = D - 1;
                    for (int attempt = 0; attempt < 2; ++attempt) {
                        dmin_el = SAP_INF; wC = 0; wE = 0;
                        if (Tg > D - 1) Tg = D - 1;                        /* every eligible label is below D */
                        for (int q = 0; q < ndirty; ++q) {
                            const int o = dirty[q];
                            if (d[o] <= Tg) wC += soff[o + 1] - soff[o];
                            wE += soff[o + 1] - soff[o];
                            if (d[o] < dmin_el) dmin_el = d[o];
                        }
                        if (wC > 0) break;
                        Tg = dmin_el + step;
                    }
                    {
                        const int64_t base = dmin_el;
                        int sh = 0;
                        while (((Tg - base) >> sh) >= 256) ++sh;
                        int64_t hist[256]; memset(hist, 0, sizeof(hist));
                        for (int q = 0; q < ndirty; ++q) {
                            const int o = dirty[q];
                            if (d[o] <= Tg) hist[(d[o] - base) >> sh] += soff[o + 1] - soff[o];
                        }
                        int64_t Tq = Tg;
                        if (wC > K && !repair) {
                            int64_t cum = 0; int bsel = 255;
                            for (int bb = 0; bb < 256; ++bb) { cum += hist[bb]; if (cum >= K) { bsel = bb; break; } }
                            Tq = base + (((int64_t)bsel + 1) << sh) - 1;
                            if (Tq > Tg) Tq = Tg;
                        }
                        T = Tq;
                    }
                    const int chain_now = repair ? 0 : (sap_chain > 3 ? 3 : sap_chain);
                    if (repair) repair = 0;
                    else if (wC < K && wC < wE) step *= 2;                 /* the window was too small (not: too little left) */
                    else if (wC > 4 * K && step > 1) step /= 2;
                    Tnext = T + step;
                    ++st[4];
                    memcpy(snap, d, sizeof(int64_t) * (size_t)O);
                    int nkeep = 0, nfront = 0;
                    for (int q = 0; q < ndirty; ++q) { const int o = dirty[q]; if (snap[o] <= T) front[nfront++] = o; else dirty[nkeep++] = o; }
                    ndirty = nkeep;
                    for (int q = 0; q < nfront; ++q) indirty[front[q]] = 0;
                    Tg = Tnext;
                    int64_t rows = 0;
                    /* work queue: the frontier (hop 0), then the chained discoveries; the outcome is order-independent */
                    int qn = 0;
                    for (int q = 0; q < nfront; ++q) { cq_o[qn] = front[q]; cq_d[qn] = snap[front[q]]; cq_h[qn] = 0; ++qn; }
                    if (chain_now) for (int o = 0; o < O; ++o) chainlab[o] = SAP_INF;
                    for (int q = 0; q < qn; ++q) {
                        const int o = cq_o[q];
                        const int64_t dl = cq_d[q];
                        const int may_chain = cq_h[q] < chain_now && soff[o + 1] - soff[o] <= SAP_CHAIN_CAP;
                        int kbest = -1; int64_t ndbest = SAP_INF;
                        for (int t = soff[o]; t < soff[o + 1]; ++t) {
                            const int i = slot_owner[t];
                            if (i < 0) continue;
                            ++rows;
                            const int32_t *r = m + (size_t)i * ld;
                            const int64_t base = (int64_t)(r[o] - cmin) * S + lambda[o];
                            for (int k = 0; k < O; ++k) {
                                if (k == o || lambda[k] >= SAP_INF / 2) continue;
                                const int64_t nd = dl + (int64_t)(r[k] - cmin) * S + lambda[k] + eps - base;
                                if (nd < dl) { rc = -5; goto done; }          /* eps-CS violated: model bug */
                                if (nd >= snap[k]) continue;                    /* counts iff strictly below the round-start label */
                                if (nd < d[k] || (nd == d[k] && t < pred[k])) {
                                    d[k] = nd; pred[k] = t;
                                    if (!indirty[k]) { indirty[k] = 1; dirty[ndirty++] = k; }
                                }
                                if (may_chain && nd <= T && nd < D && (soff[k + 1] - soff[k]) - nfreeslot[k] > 0 &&
                                    (nd < ndbest || (nd == ndbest && k < kbest))) { ndbest = nd; kbest = k; }
                            }
                        }
                        if (kbest >= 0) {
                            cq_o[qn] = kbest; cq_d[qn] = ndbest; cq_h[qn] = cq_h[q] + 1; ++qn;
                            if (ndbest < chainlab[kbest]) chainlab[kbest] = ndbest;
                        }
                    }
                    if (chain_now) {
                        /* chain-relaxed with what turned out to be its label at the end of the round: nothing left to do for it */
                        int nk = 0;
                        for (int q2 = 0; q2 < ndirty; ++q2) {
                            const int o2 = dirty[q2];
                            if (chainlab[o2] == d[o2]) indirty[o2] = 0; else dirty[nk++] = o2;
                        }
                        ndirty = nk;
                    }
                    st[5] += rows;
                }
                /* candidates: free objects with label <= D by (label, object); rank = position */
                int nc = 0;
                for (int q = 0; q < nfo; ++q) if (d[fo[q]] <= D) { cand[nc].d = d[fo[q]]; cand[nc].o = fo[q]; ++nc; }
                qsort(cand, (size_t)nc, sizeof(lab_t), lab_cmp);
                if (nc > want) nc = want;
                for (int o = 0; o < O; ++o) claim[o] = INT32_MAX;
                for (int k = 0; k < nfree; ++k) claim_src[k] = INT32_MAX;
                for (int c = 0; c < nc; ++c) {
                    int o = cand[c].o, guard = 0;
                    for (;;) {
                        if (claim[o] > c) claim[o] = c;
                        const int s = pred[o];
                        if (s >= P) { if (claim_src[s - P] > c) claim_src[s - P] = c; break; }
                        o = slot_obj[s];
                        if (++guard > P) { rc = -7; goto done; }
                    }
                }
                /* price update (before the flips: slot prices of the slots taken are the NEW object prices) */
                for (int o = 0; o < O; ++o) {
                    if (d[o] < D) {
                        lambda[o] += D - d[o];
                        for (int t = soff[o]; t < soff[o + 1]; ++t) if (slot_price[t] < lambda[o]) slot_price[t] = lambda[o];
                        touched[o] = 1;
                    } else touched[o] = 0;
                }
                int applied = 0;
                memset(onpath, 0, (size_t)O);
                for (int c = 0; c < nc; ++c) {
                    int o = cand[c].o, ok = 1;
                    for (;;) {
                        if (claim[o] != c) { ok = 0; break; }
                        const int s = pred[o];
                        if (s >= P) { ok = claim_src[s - P] == c; break; }
                        o = slot_obj[s];
                    }
                    if (!ok) continue;
                    o = cand[c].o;
                    int slot = -1;
                    for (int t = soff[o]; t < soff[o + 1]; ++t) if (slot_owner[t] < 0) { slot = t; break; }
                    for (;;) {
                        const int s = pred[o];
                        const int p = s >= P ? freel[s - P] : slot_owner[s];
                        slot_owner[slot] = p; slot_price[slot] = lambda[o]; person_obj[p] = o; person_slot[p] = slot;
                        touched[o] = 1; onpath[o] = 1;
                        if (s >= P) break;
                        o = slot_obj[s]; slot = s;
                    }
                    ++applied;
                }
                if (applied == 0) { rc = -8; goto done; }
                st[6] += applied;
                for (int o = 0; o < O; ++o) if (touched[o]) REFRESH(o);
                {
                    int nn = 0;
                    for (int k = 0; k < nfree; ++k) { const int i = freel[k]; srcmap[k] = -1; if (person_obj[i] < 0) { srcmap[k] = nn; nextl[nn++] = i; } }
                    int32_t *tmp = freel; freel = nextl; nextl = tmp; nfree = nn;
                }
                if (sap_warm && nfree > 0) {
                    /* keep the part of the forest whose root is still free and whose chain avoids every object of an
                     * applied path: labels shift by D (tree arcs stay consistent under the price update), the rest is
                     * forgotten.  vstate: 0 unknown, 1 kept, 2 dropped. */
                    for (int o = 0; o < O; ++o) vstate[o] = (d[o] >= SAP_INF / 2 || onpath[o]) ? 2 : 0;
                    for (int o = 0; o < O; ++o) {
                        if (vstate[o]) continue;
                        int x = o, depth = 0, res;
                        for (;;) {                                            /* walk up to the first decided node */
                            chain[depth++] = x;
                            const int sl = pred[x];
                            if (sl >= P) { res = srcmap[sl - P] >= 0 ? 1 : 2; break; }
                            x = slot_obj[sl];
                            if (vstate[x]) { res = vstate[x]; break; }
                        }
                        for (int q = 0; q < depth; ++q) vstate[chain[q]] = (char)res;
                    }
                    for (int o = 0; o < O; ++o) {
                        if (vstate[o] == 1) {
                            d[o] = d[o] > D ? d[o] - D : 0;
                            if (pred[o] >= P) pred[o] = P + srcmap[pred[o] - P];
                        } else { d[o] = SAP_INF; pred[o] = -1; }
                    }
                    have_forest = 1;
                } else have_forest = 0;
                continue;
            }
            have_forest = 0;
            ++st[1]; st[2] += nfree;
            for (int k = 0; k < nfree; ++k) {
                const int i = freel[k];
                int64_t b1, b2; int o1;
                SCAN(i, b1, b2, o1);
                const int64_t bp = lambda[o1] + GAP(b1, b2) + eps;
                kobj[k] = o1;
                if (bidr[o1] < 0 || bp > bidp[o1] || (bp == bidp[o1] && i < bidr[o1])) { bidp[o1] = bp; bidr[o1] = i; }
            }
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
        if (ph < 64) for (int q = 1; q <= 6; ++q) sap_phase_log[ph][q] = st[q] - ph0[q];
        if (eps == 1) break;
        eps = eps / theta; if (eps < 1) eps = 1;
    }
done:;
    int64_t tot = 0;
    for (int i = 0; i < P; ++i) if (person_obj[i] >= 0) tot += m[(size_t)i * ld + person_obj[i]];
    if (total) *total = tot;
    if (stats) memcpy(stats, st, sizeof(st));
    free(soff); free(slot_price); free(minslot); free(person_slot); free(freel); free(nextl);
    free(bidp); free(bidr); free(kobj); free(d); free(snap); free(pred); free(nfreeslot);
    free(cand); free(fo); free(dirty); free(front); free(claim); free(claim_src); free(slot_obj); free(indirty); free(touched); free(onpath); free(vstate); free(chain); free(srcmap);
    return rc;
}
