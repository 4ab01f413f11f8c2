/* TEST INFRASTRUCTURE ONLY -- sequential CPU model of the synchronous
 * (Jacobi) epsilon-scaling auction that the device LAP solver
 * (cytospace_b200/csrc/lap_auction.cu) runs.  It exists so the device
 * algorithm's exactness argument and its round/bid counts can be checked
 * without a GPU; it is never linked into the product library.
 *
 * Problem: min sum_i c[i, x(i)] over permutations x; rows = spot slots,
 * columns = cells (same orientation as lapjv_oracle.c).
 *
 * Exactness: costs are scaled by (n+1), prices are int64, the last phase runs
 * with eps = 1.  eps-complementary-slackness then bounds the scaled total
 * within n*eps = n < n+1 of optimal, i.e. the unscaled total is optimal
 * (Bertsekas 1988, integer-data corollary).
 *
 * Tie-breaks (shared with the device kernels): a row's best column is the
 * lowest j attaining min_j (C[i,j] + p[j]); a column's winning bid is the
 * highest bid price, lowest row index on equal prices.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ROWP(i) (cost + (size_t)(row_map ? row_map[(i)] : (i)) * (size_t)ld)

/* stats[0]=phases, [1]=rounds, [2]=bids (row scans), [3]=full-matrix passes,
 * [4]=rounds with <=148 bidders, [5]=bids made in the Gauss-Seidel tail.
 * round_log (may be NULL, capacity round_cap): bidders per round. */
int auction_model_i32(int n, const int32_t *cost, int64_t ld, const int32_t *row_map,
                      int32_t *rowsol, int32_t *colsol, int64_t *price,
                      int64_t *total, int64_t theta, int64_t eps0_div, int keep_cs, int64_t stop_free,
                      int64_t *stats, int32_t *round_log, int64_t round_cap, int64_t tail_t)
{
    if (n <= 0) { if (total) *total = 0; return n == 0 ? 0 : -1; }
    const int64_t S = (int64_t)n + 1;
    int32_t cmin = INT32_MAX, cmax = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        const int32_t *r = ROWP(i);
        for (int j = 0; j < n; ++j) { if (r[j] < cmin) cmin = r[j]; if (r[j] > cmax) cmax = r[j]; }
    }
    int32_t *freel = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *nextl = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int64_t *bidp  = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);   /* best bid price per column */
    int32_t *bidr  = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);   /* bidding row per column */
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (!freel || !nextl || !bidp || !bidr || !touched) return -2;
    for (int j = 0; j < n; ++j) { price[j] = 0; colsol[j] = -1; bidr[j] = -1; }
    for (int i = 0; i < n; ++i) rowsol[i] = -1;
    int64_t st[6] = {0, 0, 0, 1, 0, 0};

    int64_t eps = ((int64_t)cmax - (int64_t)cmin) * S / (eps0_div > 0 ? eps0_div : 4);
    if (eps < 1) eps = 1;
    for (;;) {
        ++st[0];
        /* phase start: decide who is free */
        int nfree = 0;
        if (st[0] == 1 || !keep_cs) {
            for (int i = 0; i < n; ++i) { rowsol[i] = -1; freel[nfree++] = i; }
            for (int j = 0; j < n; ++j) colsol[j] = -1;
        } else {
            /* keep pairs that already satisfy eps-CS at the new eps */
            ++st[3];
            for (int i = 0; i < n; ++i) {
                const int32_t *r = ROWP(i);
                int64_t m = INT64_MAX;
                for (int j = 0; j < n; ++j) { int64_t h = (int64_t)r[j] * S + price[j]; if (h < m) m = h; }
                int j0 = rowsol[i];
                if (j0 < 0 || (int64_t)r[j0] * S + price[j0] > m + eps) {
                    if (j0 >= 0) colsol[j0] = -1;
                    rowsol[i] = -1; freel[nfree++] = i;
                }
            }
        }
        while (nfree > (eps == 1 ? 0 : stop_free)) {
            if (nfree <= tail_t) {
                /* Gauss-Seidel tail (what CTA 0 runs alone on the device): FIFO of free rows, every
                 * bid sees the prices left by the previous one; the lone bidder always wins. */
                int head = 0, tailp = nfree;                 /* circular queue in freel[0..n) */
                int cnt = nfree;
                while (cnt > 0) {
                    int i = freel[head]; head = (head + 1) % n; --cnt;
                    const int32_t *r = ROWP(i);
                    int64_t b1 = INT64_MAX, b2 = INT64_MAX; int j1 = -1;
                    for (int j = 0; j < n; ++j) {
                        int64_t h = (int64_t)r[j] * S + price[j];
                        if (h < b1) { b2 = b1; b1 = h; j1 = j; }
                        else if (h < b2) b2 = h;
                    }
                    price[j1] += (n > 1 ? b2 - b1 : 0) + eps;
                    int old = colsol[j1];
                    colsol[j1] = i; rowsol[i] = j1;
                    if (old >= 0) { rowsol[old] = -1; freel[tailp % n] = old; tailp = (tailp + 1) % n; ++cnt; }
                    ++st[2]; ++st[5];
                }
                nfree = 0;
                break;
            }
            if (round_log && st[1] < round_cap) round_log[st[1]] = nfree;
            ++st[1]; st[2] += nfree;
            if (nfree <= 148) ++st[4];
            int ntouched = 0;
            for (int k = 0; k < nfree; ++k) {
                int i = freel[k];
                const int32_t *r = ROWP(i);
                int64_t b1 = INT64_MAX, b2 = INT64_MAX; int j1 = -1;
                for (int j = 0; j < n; ++j) {
                    int64_t h = (int64_t)r[j] * S + price[j];
                    if (h < b1) { b2 = b1; b1 = h; j1 = j; }
                    else if (h < b2) b2 = h;
                }
                int64_t gamma = (n > 1 ? b2 - b1 : 0) + eps;
                int64_t bp = price[j1] + gamma;
                if (bidr[j1] < 0) { touched[ntouched++] = j1; bidp[j1] = bp; bidr[j1] = i; }
                else if (bp > bidp[j1] || (bp == bidp[j1] && i < bidr[j1])) { bidp[j1] = bp; bidr[j1] = i; }
            }
            /* resolve: the winner of every touched column takes it at its bid
             * price; the previous owner (if any) becomes free.  The next free
             * list is losers + displaced owners; its order is irrelevant to a
             * Jacobi round (all bids of a round see the same prices). */
            int nnext = 0;
            for (int t = 0; t < ntouched; ++t) {
                int j = touched[t];
                int w = bidr[j];
                int old = colsol[j];
                if (old >= 0) { rowsol[old] = -1; nextl[nnext++] = old; }
                colsol[j] = w; rowsol[w] = j; price[j] = bidp[j];
                bidr[j] = -1;
            }
            for (int k = 0; k < nfree; ++k) { int i = freel[k]; if (rowsol[i] < 0) nextl[nnext++] = i; }
            int32_t *tmp = freel; freel = nextl; nextl = tmp; nfree = nnext;
        }
        if (eps == 1) break;
        eps = eps / theta; if (eps < 1) eps = 1;
    }
    int64_t tot = 0;
    for (int i = 0; i < n; ++i) tot += ROWP(i)[rowsol[i]];
    if (total) *total = tot;
    if (stats) memcpy(stats, st, sizeof(st));
    free(freel); free(nextl); free(bidp); free(bidr); free(touched);
    return 0;
}
