/* TEST INFRASTRUCTURE ONLY -- never linked, imported or called by the product
 * path (cytospace_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker /
 * reported CPU baseline.
 *
 * CPU restatement of the dense Jonker-Volgenant shortest-augmenting-path LAP
 * that CytoSPACE reaches through the third-party wheel `lapjv==1.3.14`
 * (src-d/lapjv; pinned in /root/reference/README.md:64-66; imported at
 * cytospace/linear_assignment_solvers/linear_assignment_solvers.py:16-18 and
 * called at :38; the alternative `lap==0.4.0` at :13-15/:36, pinned in
 * /root/reference/environment.yml:9).
 *
 * PARITY UNPINNED against the wheel itself: neither wheel nor its source is in
 * /root/reference or installed in the build container, and the reference ships
 * no tests or golden vectors.  The algorithm is therefore restated from the
 * published method (Jonker & Volgenant, Computing 38, 1987: column reduction,
 * reduction transfer, two augmenting-row-reduction passes, Dijkstra
 * augmentation).  What IS pinned (tests/test_oracle.py):
 *   - total cost == scipy.optimize.linear_sum_assignment (independent C++
 *     implementation) on every golden matrix and on random property cases;
 *   - dual feasibility c[i,j]-u[i]-v[j] >= 0, tight on assigned pairs;
 *   - committed golden permutations (tests/golden/) so the restatement itself
 *     cannot drift.
 * Tie-break contract of this restatement: every comparison is strict `<`
 * scanning indices in increasing order, i.e. the lowest index wins on equal
 * values.
 *
 * Two instantiations:
 *   lapjv_i32_*  int32 costs, int64 duals/total (exact; the matrix the GPU solves)
 *   lapjv_f64_*  float64 costs (the reference's own dtype, cytospace.py:326-329)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define FN(name) lapjv_i32_##name
#define COST_T int32_t
#define DUAL_T int64_t
#define DUAL_MAX INT64_MAX
#include "lapjv_body.inc"
#undef FN
#undef COST_T
#undef DUAL_T
#undef DUAL_MAX

#define FN(name) lapjv_f64_##name
#define COST_T double
#define DUAL_T double
#define DUAL_MAX DBL_MAX
#include "lapjv_body.inc"
#undef FN
#undef COST_T
#undef DUAL_T
#undef DUAL_MAX

/* Total cost of a given assignment on the int32 matrix (checker helper). */
int64_t lapjv_i32_assignment_cost(int n, const int32_t *cost, int64_t ld,
                                  const int32_t *row_map, const int32_t *rowsol)
{
    int64_t t = 0;
    for (int i = 0; i < n; ++i) {
        size_t r = (size_t)(row_map ? row_map[i] : i);
        t += cost[r * (size_t)ld + (size_t)rowsol[i]];
    }
    return t;
}

/* Minimum reduced cost c[i,j]-u[i]-v[j] over the whole matrix (>= 0 iff the
 * duals are feasible) -- checker helper for device-produced duals. */
int64_t lapjv_i32_min_reduced_cost(int n, const int32_t *cost, int64_t ld,
                                   const int32_t *row_map, const int64_t *u,
                                   const int64_t *v)
{
    int64_t m = INT64_MAX;
    for (int i = 0; i < n; ++i) {
        const int32_t *r = cost + (size_t)(row_map ? row_map[i] : i) * (size_t)ld;
        for (int j = 0; j < n; ++j) {
            int64_t h = (int64_t)r[j] - u[i] - v[j];
            m = h < m ? h : m;
        }
    }
    return m;
}
