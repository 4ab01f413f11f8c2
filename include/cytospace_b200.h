/* cytospace_b200.h -- C ABI of the B200-native CytoSPACE assignment hot path.
 *
 * One shared library, libcytospace_b200.so, built from cytospace_b200/csrc/ for
 * sm_100a.  Plain pointers and sizes only: every *_dev pointer is a CUDA device
 * pointer owned by the caller (the Python host layer takes them from torch
 * tensors), `stream` is a cudaStream_t passed as void*.  The library never
 * frees caller memory and keeps no state between calls except the per-thread
 * error string.  Every function returns 0 on success, a negative cyb_status
 * otherwise; cyb_last_error() describes the last failure on this thread.
 *
 * Reference interfaces replaced (paths relative to the CytoSPACE repo):
 *   cyb_standardise / cyb_cost_gemm_i32 / cyb_cost_build_pearson
 *       <- matrix_correlation_pearson, cytospace/common/common.py:190-199,
 *          called from calculate_cost,
 *          cytospace/linear_assignment_solvers/linear_assignment_solvers.py:53-55
 *          (the `location_repeat` gather at :63-66 is replaced by `row_map`,
 *          resolved by index inside the LAP row scans; normalize_data,
 *          common.py:142-147, is the optional fused `log_tpm` pre-step)
 *   cyb_quantise_f64
 *       <- the float64 cost matrix handed to call_solver,
 *          linear_assignment_solvers.py:34-40 (entry P2: a host cost matrix
 *          supplied by an unmodified CytoSPACE), integerised with the scale
 *          precedent of cytospace/cytospace.py:337
 *   cyb_lap_solve_i32
 *       <- lapjv.lapjv(cost) as called by call_solver,
 *          linear_assignment_solvers.py:34-40 (third-party lapjv==1.3.14) from
 *          solve_linear_assignment_problem, cytospace/cytospace.py:323-332
 *   cyb_rank_columns / cyb_cost_gemm_euclid_i32 / cyb_cost_build
 *       <- the other two `--distance-metric` values (argument_parser.py:72-74):
 *          matrix_correlation_spearman, cytospace/common/common.py:202-215
 *          (pd.DataFrame(v).rank() then the Pearson formula) and
 *          scipy cdist(.., 'euclidean'), linear_assignment_solvers.py:51,59
 *   cyb_expand_rows_noise_i32
 *       <- the integerised lap_CSPR matrix, cytospace/cytospace.py:334-340:
 *          int(1e6 * cost[location_repeat, :] + 10 * U(0,1) + 1)
 *   cyb_lap_check_i32
 *       <- no reference counterpart: on-device optimality certificate
 *          (eps-complementary slackness of the returned prices)
 */
#ifndef CYTOSPACE_B200_H
#define CYTOSPACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYB_ABI_VERSION 9

/* status codes */
#define CYB_OK                 0
#define CYB_ERR_INVALID       -1   /* bad argument (null pointer, n <= 0, misaligned, unsupported size) */
#define CYB_ERR_CUDA          -2   /* a CUDA runtime / driver call failed */
#define CYB_ERR_WORKSPACE     -3   /* workspace too small */
#define CYB_ERR_OVERFLOW      -4   /* price range exceeded the 46-bit bid field */
#define CYB_ERR_NOT_CONVERGED -5   /* round cap hit (should not happen) */
#define CYB_ERR_UNSUPPORTED   -6   /* device is not sm_100 / cooperative launch unavailable */

/* input element types of the expression matrices */
#define CYB_F64 0
#define CYB_F32 1

/* distance metrics (--distance-metric, argument_parser.py:72-74) */
#define CYB_METRIC_PEARSON   0
#define CYB_METRIC_SPEARMAN  1
#define CYB_METRIC_EUCLIDEAN 2

/* operand precision of the correlation GEMM (fp16 operands, fp32 accumulate) */
#define CYB_PREC_F16    0   /* one fp16 pass                                         */
#define CYB_PREC_F16X3  1   /* hi/lo split, 3 products in one K-concatenated GEMM:
                               operand error ~2^-22, fp32-accumulation bound         */

/* number of int64 entries written to stats_dev by cyb_lap_solve_i32 */
#define CYB_LAP_NSTATS 32
/* stats_dev layout (csrc/lap_sap.cu; times are globaltimer ns as seen by CTA 0):
 *  [0] status (0 ok)        [1] eps-scaling phases      [2] auction rounds
 *  [3] bids (= row scans in auction rounds)  [4] phase starts with a re-check pass (phases - 1)
 *  [5] cost minimum         [6] cost maximum            [7] scale S = persons + 1
 *  [8] grid size used       [9] 1 if prices / search snapshot were shared-memory resident
 *  [10] memory variant (2: holders / predecessors in global memory, 3: in shared memory)
 *  [11] max bidders in a round  [12] rows re-checked at phase starts
 *  [13] rows relaxed in searches  [14] searches  [15] search rounds
 *  [16] auction rounds with at most one bidder per CTA, and for those:
 *  [17] time in the bid scan  [18] time in the grid barrier  [19] time in the record replay
 *  [20] time in searches    [21] augmenting paths applied
 *  [22] search time: selection  [23] relaxation  [24] price update + augmentation + warm-start bookkeeping
 *  [25] selection: classification pass  [26] selection: threshold / deal
 *  [27] 1 if searches were warm-started
 *  [28] time in phase starts (re-check pass + free list)  [29] time in all auction rounds
 *  [30] time before the first phase (state init, cost range)  [31] total kernel time */

int         cyb_abi_version(void);
const char *cyb_last_error(void);

/* Device properties the host layer needs for planning (sm count, cc, bytes). */
int cyb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                    size_t *total_mem_bytes);

/* ---------------------------------------------------------------- cost build */

/* Columns of the K-major operand one standardised column occupies:
 * round_up(G, 64) for CYB_PREC_F16, 3 * round_up(G, 64) for CYB_PREC_F16X3. */
int64_t cyb_operand_k(int64_t n_genes, int precision);

/* Bytes of scratch cyb_standardise needs (per-column partial sums). */
size_t cyb_standardise_workspace_bytes(int64_t n_genes, int64_t n_cols);

/* Per-column standardisation + transpose of one expression matrix.
 *   x_dev      [n_genes x n_cols] row-major (genes x cells, as the reference's
 *              DataFrame.to_numpy()), leading dimension ld_x elements
 *   log_tpm    1: apply normalize_data (TPM, log2(x+1)) first; 0: x is already
 *              normalised (what solve_linear_assignment_problem receives)
 *   operand_b  0: this matrix is the A operand (rows of the cost matrix: cells), 1: the
 *              B operand (columns: spots); only matters for CYB_PREC_F16X3
 *              (A = [hi|hi|lo], B = [hi|lo|hi])
 *   z_dev      fp16 [n_cols x cyb_operand_k()] K-major: z = (x - mean) / sigma,
 *              K padding zero-filled by this call
 *   colstat_dev float64 [2 x n_cols] out: mean, population sigma
 *   zero_var_dev int32[1]: incremented by the number of sigma == 0 columns
 *              (their z is all-zero => r = 0; the reference yields NaN,
 *              common.py:196-197 -- the host layer raises on it)            */
int cyb_standardise(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols,
                    int64_t ld_x, int log_tpm, int precision, int operand_b,
                    void *z_dev, double *colstat_dev, int32_t *zero_var_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream);

/* cost[a, b] = rint(-scale * sum_k za[a,k] * zb[b,k])  as int32,
 * a TMA-fed tcgen05 GEMM with TMEM accumulators and a fused quantise epilogue.
 *   za_dev [n_a x k] fp16, zb_dev [n_b x k] fp16 (K-major, k % 64 == 0,
 *   16-byte aligned); cost_dev [n_a x ld_cost] int32 row-major.
 * With z from cyb_standardise, scale = 1e6 / n_genes gives rint(-1e6 * r).   */
int cyb_cost_gemm_i32(const void *za_dev, const void *zb_dev, int64_t n_a,
                      int64_t n_b, int64_t k, float scale, int32_t *cost_dev,
                      int64_t ld_cost, void *stream);

/* Bytes of device workspace cyb_cost_build_pearson needs. */
size_t cyb_cost_build_workspace_bytes(int64_t n_genes, int64_t n_a, int64_t n_b, int precision);

/* calculate_cost, Pearson branch, in one call: standardise both matrices and run the GEMM.
 *   cost[i, j] = rint(-cost_scale * pearson(a[:, i], b[:, j]))   int32 [n_a x ld_cost]
 * a_dev [n_genes x n_a], b_dev [n_genes x n_b] (genes x columns, row-major).  The caller picks
 * the orientation: a = ST, b = scRNA gives the reference's `cost` (spots x cells,
 * linear_assignment_solvers.py:55); a = scRNA, b = ST its transpose (cells x spots), the
 * layout the capacitated LAP scans.  colstat_*_dev: float64 [2 x n] out (may be NULL).   */
int cyb_cost_build_pearson(const void *a_dev, const void *b_dev, int x_dtype,
                           int64_t n_genes, int64_t n_a, int64_t n_b,
                           int64_t ld_a, int64_t ld_b, int log_tpm, int precision,
                           double cost_scale, int32_t *cost_dev, int64_t ld_cost,
                           double *colstat_a_dev, double *colstat_b_dev,
                           int32_t *zero_var_dev, void *workspace_dev,
                           size_t workspace_bytes, void *stream);

/* ------------------------------------------------- Spearman / Euclidean / lap_CSPR */

/* Bytes of scratch cyb_rank_columns needs. */
size_t cyb_rank_workspace_bytes(int64_t n_genes, int64_t n_cols);

/* Per-column average ranks, `pd.DataFrame(x).rank().values` (common.py:207-208; method
 * "average", ascending, 1-based; ties are exact ties of the float64 values, NaN counts as 0):
 *   rank[g, c] = #{g': x[g',c] < x[g,c]} + (#{g': x[g',c] == x[g,c]} + 1) / 2
 *   x_dev     [n_genes x n_cols] row-major, float64 or float32, leading dimension ld_x
 *   log_tpm   1: rank normalize_data(x) (common.py:142-147) instead of x
 *   rank_dev  float32 [n_genes x ld_rank] out (exact: ranks are multiples of 0.5 below 2^24)
 * One CTA per column: the column is sorted in shared memory (bitonic network on
 * order-preserving 64-bit keys), every element binary-searches its tie group.            */
int cyb_rank_columns(const void *x_dev, int x_dtype, int64_t n_genes, int64_t n_cols,
                     int64_t ld_x, int log_tpm, float *rank_dev, int64_t ld_rank,
                     void *workspace_dev, size_t workspace_bytes, void *stream);

/* Euclidean distance from the standardised operands of cyb_standardise and its colstat output
 * ([mean | sigma] per column):  with r = (1/n_genes) * sum_k za[a,k] * zb[b,k],
 *   |a - b|^2 = n_genes * ((mu_a - mu_b)^2 + (sd_a - sd_b)^2 + 2 sd_a sd_b (1 - r))
 *   cost[a, b] = rint(scale * sqrt(.))
 * -- the same tcgen05 GEMM as cyb_cost_gemm_i32 with a float64 epilogue; no cancellation of the
 * large norms.  bad_dev int32[1] is incremented per entry >= 2^30 (clamped).               */
int cyb_cost_gemm_euclid_i32(const void *za_dev, const void *zb_dev, int64_t n_a, int64_t n_b,
                             int64_t k, int64_t n_genes, float scale, const double *colstat_a_dev,
                             const double *colstat_b_dev, int32_t *cost_dev, int64_t ld_cost,
                             int32_t *bad_dev, void *stream);

/* Bytes of device workspace cyb_cost_build needs for `metric`. */
size_t cyb_cost_build_metric_workspace_bytes(int metric, int64_t n_genes, int64_t n_a, int64_t n_b,
                                             int precision);

/* calculate_cost (linear_assignment_solvers.py:42-59) for any --distance-metric:
 *   CYB_METRIC_PEARSON    cost = rint(-cost_scale * pearson(a_i, b_j))
 *   CYB_METRIC_SPEARMAN   cost = rint(-cost_scale * pearson(rank(a_i), rank(b_j)))
 *   CYB_METRIC_EUCLIDEAN  cost = rint( cost_scale * ||a_i - b_j||_2)
 * Arguments as cyb_cost_build_pearson.  bad_dev int32[1] is incremented per zero-variance
 * column (correlations) or per unrepresentable entry (Euclidean).                         */
int cyb_cost_build(int metric, const void *a_dev, const void *b_dev, int x_dtype,
                   int64_t n_genes, int64_t n_a, int64_t n_b, int64_t ld_a, int64_t ld_b,
                   int log_tpm, int precision, double cost_scale, int32_t *cost_dev,
                   int64_t ld_cost, int32_t *bad_dev, void *workspace_dev,
                   size_t workspace_bytes, void *stream);

/* out[i, j] = cost[row_map[i], j] + noise_lo + floor(noise_span * U(seed, i, j))  (int32):
 * the slot expansion cost[location_repeat, :] (linear_assignment_solvers.py:63-66) with the
 * integer tie noise of the lap_CSPR path (cytospace.py:337-340: + 10 * rand + 1, i.e.
 * noise_lo = 1, noise_span = 10).  U is a counter-based hash (splitmix64 of seed, i, j), not
 * the reference's MT19937 stream.  row_map_dev NULL: identity; noise_span 0: no noise.     */
int cyb_expand_rows_noise_i32(const int32_t *cost_dev, int64_t ld, int64_t n_rows_out,
                              int64_t n_cols, const int32_t *row_map_dev, uint64_t seed,
                              int noise_lo, int noise_span, int32_t *out_dev, int64_t ld_out,
                              void *stream);

/* out[i, j] = rint(scale * in[i, j]) as int32 (entry P2: an n_rows x n_cols float64
 * cost matrix built by the reference itself).  bad_dev int32[1] is incremented by
 * the number of non-finite or out-of-range (|scale*x| >= 2^30) entries.           */
int cyb_quantise_f64(const double *in_dev, int64_t n_rows, int64_t n_cols, int64_t ld_in,
                     double scale, int32_t *out_dev, int64_t ld_out, int32_t *bad_dev,
                     void *stream);

/* ----------------------------------------------------------------------- LAP */

/* Bytes of device workspace cyb_lap_solve_i32 / cyb_lap_check_i32 need. */
size_t cyb_lap_workspace_bytes(int64_t n_persons, int64_t n_objects);

/* Exact dense assignment in transportation form:
 *     min sum_i cost[i, obj(i)]   s.t. object o holds exactly cap[o] persons.
 * For CytoSPACE the persons are the cells, the objects the spots and cap =
 * cell_number_to_node_assignment: the same optimisation problem as lapjv on the
 * expanded matrix cost[location_repeat, :] (linear_assignment_solvers.py:63-66) with the
 * same optimal total, without materialising the expansion.
 *   cost_dev        int32 [n_persons x ld] row-major (cells x spots); |cost| < 2^30
 *   slot_offset_dev int32[n_objects + 1], exclusive prefix sum of the capacities
 *                   (slot_offset[n_objects] == n_persons), or NULL: every capacity is 1
 *                   (then n_objects == n_persons: the plain square LAP)
 *   person_obj_dev  int32[n_persons] out: object (spot) of person (cell) i -- with cap == 1
 *                   this is lapjv's row_ind for the matrix as given
 *   slot_owner_dev  int32[n_persons] out: person held by slot t; the slots of object o are
 *                   slot_offset[o] .. slot_offset[o+1]-1 -- with cap == 1 this is lapjv's
 *                   col_ind, the vector CytoSPACE consumes (linear_assignment_solvers.py:38)
 *   price_dev       int64[n_objects] out: object prices in units of 1/(n_persons+1) cost
 *   total_dev       int64[1] out: sum_i cost[i, person_obj[i]]
 *   stats_dev       int64[CYB_LAP_NSTATS] out (layout above)
 *   grid_hint       0 = auto (one CTA per SM); otherwise the number of CTAs
 * Synchronous eps-scaling auction (Jacobi rounds over the grid; the last few bidders of a
 * phase are finished by one CTA, as a Gauss-Seidel FIFO or as in-CTA Jacobi rounds depending
 * on what is shared-memory resident, stats[10]) in one persistent cooperative kernel; costs
 * scaled by n_persons+1, last phase eps = 1 => the returned assignment is optimal for the
 * integer matrix.  Deterministic for a given device and problem: lowest object index wins a
 * person's tie, highest bid then lowest person index wins an object; independent of the grid
 * size.  Limits: n_persons < 2^18, |cost| < 2^30.                                         */
int cyb_lap_solve_i32(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                      const int32_t *slot_offset_dev, int32_t *person_obj_dev,
                      int32_t *slot_owner_dev, int64_t *price_dev, int64_t *total_dev,
                      int64_t *stats_dev, void *workspace_dev, size_t workspace_bytes,
                      int grid_hint, void *stream);

/* Optimality certificate / row-scan pass: for every person computes
 * m_i = min_o (cost[i,o]*(n_persons+1) + price[o]) and writes
 *   out_dev[0] = max_i ( cost[i,obj(i)]*(n_persons+1) + price[obj(i)] - m_i )
 *                (<= 1 together with out[2] == out[3] == 0 certifies optimality),
 *   out_dev[1] = total cost, out_dev[2] = persons without a valid object,
 *   out_dev[3] = objects whose holder count differs from their capacity.
 * One coalesced pass over the whole matrix (n_persons*n_objects*4 bytes): the HBM roofline
 * probe of the LAP row scan.  Workspace: cyb_lap_workspace_bytes().  out_dev: int64[4]. */
int cyb_lap_check_i32(const int32_t *cost_dev, int64_t ld, int64_t n_persons, int64_t n_objects,
                      const int32_t *slot_offset_dev, const int32_t *person_obj_dev,
                      const int64_t *price_dev, int64_t *out_dev, void *workspace_dev,
                      size_t workspace_bytes, void *stream);

/* ------------------------------------------------------ multi-GPU data plane */

/* The chunked problem (apply_linear_assignment, cytospace/cytospace.py:430-467) on one rank per GPU: the
 * reference pickles per-chunk column blocks `scRNA_norm_np[:, index_sc_list[i]]` / `st_norm_np[:, index_st_list[i]]`
 * (:434-443; all spots for --sampling-sub-spots, :438) to worker processes and collects the index lists as they
 * complete (:453-467).  Here: cyb_gather_columns cuts the blocks on rank 0's GPU, cyb_dist_send / cyb_dist_recv
 * move them over NVLink (cyb_dist_broadcast for the shared ST block), cyb_dist_all_gather returns the indices.
 * NCCL is bound at run time (libnccl.so.2); a communicator is an opaque handle owned by the caller.
 * Sizes are in bytes; calls are asynchronous on `stream` like NCCL's own. */
#define CYB_DIST_ID_BYTES 128

/* rank 0: a fresh communicator id (ncclGetUniqueId) to hand to every rank out of band. */
int cyb_dist_unique_id(void *id_out);
/* every rank, on its current CUDA device: join the communicator (collective; ncclCommInitRank). */
int cyb_dist_init(const void *id_bytes, int n_ranks, int rank, void **comm_out);
int cyb_dist_destroy(void *comm);
int cyb_dist_broadcast(void *comm, void *buf_dev, size_t bytes, int root, void *stream);
int cyb_dist_send(void *comm, const void *buf_dev, size_t bytes, int peer, void *stream);
int cyb_dist_recv(void *comm, void *buf_dev, size_t bytes, int peer, void *stream);
/* sends / receives issued between the two calls progress concurrently (ncclGroupStart / ncclGroupEnd): rank 0
 * feeds several owners at once, so that its NVLink egress -- not one peer-to-peer channel set -- is the limit. */
int cyb_dist_group_start(void);
int cyb_dist_group_end(void);
/* recv_dev holds n_ranks * bytes_per_rank bytes, rank r's part at r * bytes_per_rank. */
int cyb_dist_all_gather(void *comm, const void *send_dev, void *recv_dev, size_t bytes_per_rank, void *stream);

/* out[r, j] = x[r, cols[j]]: the columns of one chunk (cytospace.py:434-443) as a dense block.
 *   x_dev [n_rows x ld_x] float64 / float32 row-major, cols_dev int32[n_cols_out], out_dev [n_rows x ld_out]. */
int cyb_gather_columns(const void *x_dev, int x_dtype, int64_t n_rows, int64_t ld_x,
                       const int32_t *cols_dev, int64_t n_cols_out, void *out_dev, int64_t ld_out,
                       void *stream);

/* out[i] = (float) x[i] for a contiguous float64 buffer of n elements; *inexact_dev (int32, zeroed by the caller) is
 * OR-ed with 1 when some value does not survive the round trip exactly (NaN counts as inexact).  The wire format of the
 * chunk blocks: raw count matrices (cytospace.py:398: the input of normalize_data) are exact in float32, which halves
 * the NVLink bytes of cyb_dist_send without changing a single bit of the result (the kernels widen on load). */
int cyb_narrow_f64_to_f32(const double *x_dev, int64_t n, float *out_dev, int32_t *inexact_dev, void *stream);

/* ------------------------------------------------------------ host staging */

/* dst_dev[0..bytes) = src_host[0..bytes) for PAGEABLE host memory -- the numpy arrays the reference passes to the
 * solver (cytospace/cytospace.py:398-443; `.to_numpy()` of the input DataFrames) -- through a process-wide ring of
 * pinned pieces filled by worker threads with non-temporal stores, one DMA per piece on an internal copy stream.
 * Returns when every piece is enqueued (src_host may be reused); `stream` waits for the last DMA.  Ordered after the
 * work already queued on `stream`.  Environment: CYB_STAGE_THREADS (default min(16, cores)), CYB_STAGE_PIECE_MB (8),
 * CYB_STAGE_PIECES (3 x threads). */
int cyb_stage_upload(const void *src_host, void *dst_dev, size_t bytes, void *stream);

/* The same upload for a float64 host array that lands on the device as float32: the worker threads narrow while they
 * copy (round to nearest even, like numpy's astype), so half the bytes cross PCIe.  *inexact_host (may be NULL) is set to
 * 1 when some value does not survive the round trip exactly (NaN counts as inexact), else 0.  The kernels accept
 * float32 input and widen on load: count matrices (integers below 2^24) are exact; on log2(TPM+1) data the rounding
 * (2^-24 relative per value) moves the correlation by < 1e-8 -- 0.2 % of the integer cost entries change by one unit,
 * next to the 2e-6 tolerance of the fp16 hi/lo GEMM itself (tests/test_gpu_path.py). */
int cyb_stage_upload_f64_as_f32(const double *src_host, float *dst_dev, size_t n, int32_t *inexact_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CYTOSPACE_B200_H */
