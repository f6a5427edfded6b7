/* gsi.h -- C ABI of libgsi.so: the sm_100a implementation of the per-user graph-signal
 * interpolation path of Dhole/collaborative_filtering.
 *
 * The reference exposes no plugin/FFI interface: its boundary is a set of CLI
 * programs whose math is inlined in main()/apply().  The entry points below are what a C++ host
 * with the reference's argv/cwd/file behaviour binds instead of that inlined math; each one
 * cites the reference code it replaces (file:line into the reference tree).  Plain C types,
 * caller-owned buffers, int status returns, no exceptions across the boundary, one opaque
 * handle per GPU.  There is no CPU fallback: every entry point fails with GSI_ERR_CUDA when no
 * device is usable.
 *
 * ID conventions (precompute_local.cpp:19,107; knn.cpp:19,103): user ids downstream are
 * user' = 2147483647 - user; movie ids are used as-is and index the weight table directly
 * (row/col 0 unused, precompute_local.cpp:153).
 */
#ifndef GSI_H_
#define GSI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsi_ctx gsi_ctx;

enum {
    GSI_OK = 0,
    GSI_ERR_INVALID = 1,   /* bad argument (null pointer, unsorted/duplicate ids, ...)        */
    GSI_ERR_CUDA = 2,      /* CUDA runtime error or no device; see gsi_last_error()            */
    GSI_ERR_NOMEM = 3,     /* device or pinned-host allocation failed                          */
    GSI_ERR_STATE = 4,     /* call order (e.g. precompute before weights were set)             */
    GSI_ERR_CAPACITY = 5,  /* caller-provided output buffer too small; totals[] says how much  */
    GSI_ERR_SINK = 6       /* the record sink returned non-zero                                 */
};

/* ---- context ------------------------------------------------------------------------------ */

/* One context per GPU.  `stream` is a cudaStream_t (or NULL: the context creates its own
 * non-blocking stream); all kernels of the context are launched on it. */
int gsi_create(gsi_ctx** out, int device, void* stream);
int gsi_destroy(gsi_ctx* ctx);
/* Last error text of the context (or of the calling thread when ctx is NULL). */
const char* gsi_last_error(const gsi_ctx* ctx);
/* Library / build identification, e.g. "gsi 0.1 sm_100a". */
const char* gsi_version(void);
/* Device workspace budget in bytes for one chunk of users (default 8 GiB). */
int gsi_set_workspace_limit(gsi_ctx* ctx, int64_t bytes);
/* Largest n that takes the CTA-resident (shared-memory Jacobi) kernel; larger users take the Householder /
 * divide-and-conquer path.  Default 44 (measured crossover), GSI_SMALL_MAX=32..160 overrides it. */
int gsi_small_max(const gsi_ctx* ctx);
/* Synchronise the context's stream. */
int gsi_sync(gsi_ctx* ctx);

/* ---- item-similarity table  (replaces precompute_local.cpp:113-158, the dense `weights`
 *      MatrixXd filled from ./out_fin_* ; the 2000-id cap of :116 is lifted) ------------------ */

/* Dense directed table, rows x rows doubles, row-major: w[m1 * rows + m2] = weight(m1 -> m2).
 * Copied to the device (synchronous). */
int gsi_set_weights_host(gsi_ctx* ctx, const double* w, int rows);
/* Same, but `d_w` already lives on this context's device; borrowed until the next
 * gsi_set_weights_* / gsi_destroy. */
int gsi_set_weights_device(gsi_ctx* ctx, const double* d_w, int rows);
/* Edge list form of ./out_fin_* ("m1 m2 w" lines): builds the zero-initialised dense table of
 * (max id + 1)^2 on the device and scatters the edges.  (m1, m2) pairs must be unique (knn2
 * emits each directed edge once); `rows_out` receives max id + 1. */
int gsi_set_weights_edges(gsi_ctx* ctx, const int32_t* m1, const int32_t* m2, const double* w,
                          int64_t n_edges, int* rows_out);
/* Device pointer and row count of the current table (for NCCL broadcast by the host side). */
int gsi_get_weights(gsi_ctx* ctx, double** d_w, int* rows);

/* ---- precompute  (replaces compute_eigens(), precompute_local_threads.cpp:100-213, i.e. the
 *      body of the user loop precompute_local.cpp:165-282: gather W, degree, normalised
 *      Laplacian, full symmetric eigensolve, sig_min, cutoff) ---------------------------------- *
 *
 * Users are given as CSR over rated movie ids: user u owns items[offsets[u] .. offsets[u+1]),
 * strictly ascending (the reference's per-user map dedups, :108; its iteration order is hash
 * order, we define ascending -- SURVEY.md B6).
 *
 * Per user the call produces what one out_eigen_ record holds (README.md:14-19):
 *   sig_min[offsets[u] + i]    (double)  ||row_i(L)||_2 (float accumulate) + 0.01     (:169-182)
 *   k[u]                                 number of kept eigenpairs, >= 2             (:185-194)
 *   lam[lam_off[u] + j], j < k           eigenvalues ascending
 *   vec[vec_off[u] + i*k + j]            eigenvectors, row-major n x k (row i = movie i), unit
 *                                        norm, sign: the largest-|.| component is positive
 * Records are laid out in PROCESSING order (users are bucketed by n); lam_off / vec_off locate
 * each user's record and are not monotonic in u.  For n == 1 the reference reads uninitialised
 * memory (:190-194); we define k = 2, lam = {1, 0}, vec = {1, 0}.
 */

/* Everything on the device.  d_lam / d_vec have capacity lam_cap / vec_cap doubles; upper
 * bounds are sum(max(n,2)) and sum(n*max(n,2)).  totals[0], totals[1] (host) receive the doubles
 * actually used.  Asynchronous on the context's stream except for a few scalar read-backs. */
int gsi_precompute_device(gsi_ctx* ctx, int64_t n_users, const int64_t* h_offsets,
                          const int32_t* d_items, double* d_sig_min, int32_t* d_k,
                          int64_t* d_lam_off, int64_t* d_vec_off, double* d_lam, int64_t lam_cap,
                          double* d_vec, int64_t vec_cap, int64_t* totals);

/* Host buffers in, host buffers out (copies included).  Same layout as above. */
int gsi_precompute_host(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets,
                        const int32_t* items, double* sig_min, int32_t* k, int64_t* lam_off,
                        int64_t* vec_off, double* lam, int64_t lam_cap, double* vec,
                        int64_t vec_cap, int64_t* totals);

/* Streaming form for the CLI hosts (the analogue of save_output(),
 * precompute_local_threads.cpp:89-98: records are handed over as they complete).  The sink is
 * called once per chunk from the calling thread with pointers into pinned host staging memory
 * that stay valid until it returns. */
typedef struct gsi_record_chunk {
    int64_t n_records;
    const int64_t* user_index;  /* [n_records] index u of the record in the caller's CSR       */
    const int32_t* n;           /* [n_records] rated movies                                      */
    const int32_t* k;           /* [n_records] kept eigenpairs                                   */
    const int64_t* lam_off;     /* [n_records] into lam                                          */
    const int64_t* vec_off;     /* [n_records] into vec                                          */
    const double* lam;
    const double* vec;
    const double* sig_min;      /* whole-batch array, index with the caller's offsets[u] + i    */
} gsi_record_chunk;
typedef int (*gsi_record_sink)(void* opaque, const gsi_record_chunk* chunk);
int gsi_precompute_stream(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets,
                          const int32_t* items, gsi_record_sink sink, void* opaque);

/* ---- predict  (replaces neigh_program::gather/apply, local_calc_precomp.cpp:200-380: for every
 *      (movie m, test user u) pair select the known rows K = rated(u) n out-neighbours(m), cut the
 *      spectrum at the per-movie sig_min, drop empty columns, solve the band-limited least squares
 *      problem and return the clamped squared error) ------------------------------------------- *
 *
 * The users' records are given in the layout gsi_precompute_* produces (what load_precomputed_data
 * :406-482 parses from out_eigen_).  The item graph is not passed separately: edge m -> j exists iff
 * (float)weights(m, j) > 0.1 (graph_loader :122-136), and the dense table set with
 * gsi_set_weights_* holds exactly the out_fin_ lines the reference builds its graph from.
 *
 *   w_lim[offsets[u] + i]    cutoff for the pair (u, movie i).  The reference reads
 *                            sigs_min[movie_ind] from a vector that is never cleared between
 *                            records (bug B1, :414,437,440,271); the host decides whether to pass
 *                            that or the record's own sig_min.
 *   ratings[offsets[u] + i]  the user's own rating of movie i (out_test_rat_, float -> double
 *                            :138-160); it is both the ground truth of pair (u, i) and a known
 *                            rating for the user's other pairs.
 *   pair_mask                optional [nnz] bytes; 0 skips the pair (--pct sampling is per movie,
 *                            :221).  NULL = every pair.
 * Outputs, all [nnz], aligned with items: err (float, (real - clamp(pred,1,5))^2 :322-327,357),
 * kk (#K :359), pred (unclamped), status (GSI_PRED_*), cols (columns used).
 * Ill-posed pairs (kk < cols, or a non-positive Cholesky pivot; the reference inverts the singular
 * Gram blindly, B3) get pred = mean of the known ratings; kk == 0 gives NaN as in the reference.
 */
enum { GSI_PRED_OK = 0, GSI_PRED_EMPTY = 1, GSI_PRED_UNDERDETERMINED = 2, GSI_PRED_SINGULAR = 3, GSI_PRED_SKIPPED = 4 };

int gsi_predict_host(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets, const int32_t* items,
                     const double* w_lim, const double* ratings, const int32_t* k,
                     const int64_t* lam_off, const int64_t* vec_off, const double* lam,
                     int64_t lam_len, const double* vec, int64_t vec_len, const uint8_t* pair_mask,
                     float* err, int32_t* kk, double* pred, int32_t* status, int32_t* cols);

/* Same with every array already on the device (records straight from gsi_precompute_device;
 * h_offsets and h_k are host copies the planner needs).  pair_mask is a host array or NULL. */
int gsi_predict_device(gsi_ctx* ctx, int64_t n_users, const int64_t* h_offsets, const int32_t* h_k,
                       const int64_t* d_offsets, const int32_t* d_items, const double* d_w_lim,
                       const double* d_ratings, const int32_t* d_k, const int64_t* d_lam_off,
                       const int64_t* d_vec_off, const double* d_lam, const double* d_vec,
                       const uint8_t* pair_mask, float* d_err, int32_t* d_kk, double* d_pred,
                       int32_t* d_status, int32_t* d_cols, int64_t* n_pairs_done);

/* ---- local_calc  (replaces vertex_program::gather/apply of the per-MOVIE variant, local_calc.cpp:246-526;
 *      SURVEY.md 8f.2).  For every movie m that has test ratings: the local graph {m} + out-neighbours(m)
 *      (edge a -> b iff (float)weights(a, b) > 0.1, carrying that float, graph_loader :102-117), its adjacency
 *      with row and column 0 = w(m -> i) (:324-335), normalised Laplacian L (:347-374) and eigenpairs of the
 *      lower triangle (:378); for every test user u of m: unrated nodes (incl. m itself) -> exact cutoff
 *      w_lim = sqrt(lambda_min(L_h L_h^T)) with L_h = the unrated rows of L (:417-436), lim = max(2, first
 *      lambda > w_lim) (:443-451), least squares on the rated rows of U[:, :lim] (:456-491), clamp, squared
 *      error (:494-499).  No out_eigen_ records are involved; the item graph comes from gsi_set_weights_*.
 *
 * Test ratings are given per user in CSR form exactly as for gsi_predict_host (items need not be sorted,
 * each (user, movie) at most once); rating 0 means "not rated" as in the reference (:406-413).  Outputs are
 * [nnz], aligned with items: pair t = (user u, movie items[t]).  status GSI_PRED_SKIPPED marks pairs the
 * reference emits no line for: masked out (pair_mask, --pct), movie id outside the table, or a local graph
 * of fewer than 3 nodes (:271-272).  kk == 0 gives NaN (GSI_PRED_EMPTY); kk < lim (only possible through
 * the max(.,2) rule) GSI_PRED_UNDERDETERMINED with pred = mean of the known ratings.  cols = lim.
 * w_lim may be NULL.  Self edges of the table are ignored.
 * How the cutoff is computed: lambda_min of P[unrated, unrated], P = L L^T per movie, by Lanczos through the movie's P
 * (no per-pair matrix); a pair whose Ritz value has not settled after 96 steps, and every pair when the environment
 * variable GSI_LC_EXACT=1 is set, is solved exactly instead (gathered matrix -> Householder tridiagonalisation ->
 * Sturm-count bracket).  Local graphs of more than 9,216 nodes take the two-stage tridiagonalisation (46,000 nodes is the limit).  Neighbour order is
 * ascending movie id (the reference: hash order); where the two directions of an edge carry different weights the
 * reference's result depends on that order, with the bit-symmetric weights knn2 emits it does not. */
int gsi_local_calc_host(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets, const int32_t* items,
                        const double* ratings, const uint8_t* pair_mask, float* err, int32_t* kk,
                        double* pred, int32_t* status, int32_t* cols, double* w_lim);

/* ---- knn chain  (replaces the GraphLab vertex programs of knn.cpp:160-298, the edge transform
 *      weights_calc knn2.cpp:127-146 and knn_program / error_vertex_data knn3.cpp:185-256) ------ *
 *
 * Ratings are CSR by user (ids ascending, unique), `rows` = max movie id + 1.
 */

/* knn2: cosine weight of every directed item pair over its common TRAIN raters (float
 * accumulators, > 5 common raters, knn2.cpp:127-146) -> the out_fin_ edge list (w > 0.01,
 * :155-163), kept on the device in ascending (m1, m2) order.  With install_weights != 0 the dense
 * table of the context is (re)built from those edges exactly as precompute_local would read them
 * back from the out_fin_ text (weights rounded to 6 significant digits). */
int gsi_knn_build_host(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets, const int32_t* items,
                       const float* ratings, int rows, int install_weights, int64_t* n_edges);
/* Copies the edge list of the last gsi_knn_build_host to the host (cap >= n_edges). */
int gsi_knn_edges_host(gsi_ctx* ctx, int32_t* m1, int32_t* m2, float* w, int64_t cap);
/* knn step 1, out_edg_*: co[a * rows + b] = 1 iff some user (train OR validate role,
 * knn.cpp:218-227) rated both a and b, a != b. */
int gsi_knn_corated_host(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets, const int32_t* items,
                         int rows, uint8_t* co);
/* knn3 over the context's weight table and the validate users' ratings: per movie the float sum of
 * (r - round(pred))^2 and the number of test ratings (knn3.cpp:234-256), plus has_edge[m] = movie
 * m is an endpoint of a kept edge ((float)w > 0.1) -- together they give num_vertices (:263). */
int gsi_knn3_host(gsi_ctx* ctx, int64_t n_users, const int64_t* offsets, const int32_t* items,
                  const float* ratings, float* movie_err_sum, int32_t* movie_cnt, uint8_t* has_edge);

/* ---- Chebyshev polynomial graph filter (cheby.cpp) ---------------------------------------------- *
 * y = 0.5 c0 T0 + sum_{k >= 1} c_k T_k(x) on the normalised Laplacian of ONE graph, interval [0, 2]
 * (cheby.cpp:17-19): T0 = x, T1 = L x - x, T_{k+1} = 2 (L T_k - T_k) - T_{k-1}, with
 * L = I - D^-1/2 W D^-1/2 and d_i = the sum of the out-edge weights of i (degree_program, :155-183).
 * Replaces the three GraphLab engines of cheby.cpp:312-375 (degree, init values, ncoef - 2 synchronous
 * supersteps).  The graph is a CSR over out-edges with the vertices numbered 0 .. nv-1: every kept line
 * `a b w` of graph_topology (w > 0.1, :96-99) contributes a -> b and b -> a, duplicates are kept.
 * ncoef >= 2; with ncoef == 2 the reference reads coeff[2] out of bounds (:262) -- here the filter stops
 * after c1.  All pointers are host memory; y receives nv values. */
int gsi_cheby_filter_host(gsi_ctx* ctx, int64_t nv, const int64_t* row_off, const int32_t* col, const double* w,
                          const double* x, int ncoef, const double* coef, double* y);

/* ---- device group: every GPU of the box behind ONE call  (replaces the thread pool of
 *      precompute_local_threads.cpp:300-314, which spreads the users over all workers of the box by itself, and the
 *      engine threads that run apply() in local_calc_precomp.cpp:566-572) ------------------------------------------ *
 *
 * A group owns one context per device and one host thread per device while a call runs.  Users are independent units
 * (each compute_eigens() task reads only the shared read-only weights, precompute_local_threads.cpp:25,120-123): they
 * are dealt over the devices by longest-processing-time-first on the cost n^3 + 64 n^2 (deterministic), every device
 * runs gsi_precompute_stream on its share, and there is no data-path collective.  The weight table is copied to the
 * first device once and replicated to the others over NVLink: ncclBroadcast when an NCCL library can be loaded
 * (libnccl.so.2, resolved at run time), otherwise a tree of peer copies (gsi_group_broadcast_path says which).
 * Records of a user are bit-identical to what a single context computes (the route of a user depends on its n only).
 */
typedef struct gsi_group gsi_group;
/* `devices` = n_devices CUDA ordinals; n_devices <= 0 (or devices == NULL) takes every visible device. */
int gsi_group_create(gsi_group** out, int n_devices, const int* devices);
int gsi_group_destroy(gsi_group* g);
int gsi_group_size(const gsi_group* g);
/* Context of member i (for per-device calls: timing, workspace limit, ...); owned by the group. */
gsi_ctx* gsi_group_ctx(gsi_group* g, int i);
const char* gsi_group_last_error(const gsi_group* g);
/* "single", "nccl" or "peer": how the last gsi_group_set_weights_host replicated the table. */
const char* gsi_group_broadcast_path(const gsi_group* g);
int gsi_group_set_workspace_limit(gsi_group* g, int64_t bytes);
/* gsi_set_weights_host on member 0, then replication to every other member (device to device). */
int gsi_group_set_weights_host(gsi_group* g, const double* w, int rows);
/* gsi_precompute_stream over all devices.  The sink is called from the device threads, one call at a time (serialised
 * by the group); user_index refers to the caller's CSR, sig_min to the caller's offsets, exactly as for one context.
 * The order in which chunks arrive depends on timing; the records themselves do not. */
int gsi_group_precompute_stream(gsi_group* g, int64_t n_users, const int64_t* offsets, const int32_t* items,
                                gsi_record_sink sink, void* opaque);
/* gsi_predict_host over all devices: users (with all their pairs, next to their U -- SURVEY.md 8e) are dealt by
 * longest-processing-time-first on the cost pairs * n * k^2; outputs land in the caller's [nnz] arrays. */
int gsi_group_predict_host(gsi_group* g, int64_t n_users, const int64_t* offsets, const int32_t* items,
                           const double* w_lim, const double* ratings, const int32_t* k,
                           const int64_t* lam_off, const int64_t* vec_off, const double* lam,
                           int64_t lam_len, const double* vec, int64_t vec_len, const uint8_t* pair_mask,
                           float* err, int32_t* kk, double* pred, int32_t* status, int32_t* cols);

/* ---- measurement helpers ------------------------------------------------------------------- */

/* Accumulated device time (ms, CUDA events on the context's stream) per kernel class since the
 * last reset, and launch counts.  Classes: see GSI_T_* . */
enum {
    GSI_T_EIG_CTA = 0,     /* fused gather+Laplacian+sig_min+CTA-resident Jacobi (n <= 160)     */
    GSI_T_LAP = 1,         /* gather / degree / Laplacian / sig_min kernels of the large path    */
    GSI_T_BJ_GRAM = 2,     /* block-Jacobi Gram (DMMA)                                           */
    GSI_T_BJ_INNER = 3,    /* block-Jacobi 32x32 rotation solve                                  */
    GSI_T_BJ_UPDATE = 4,   /* block-Jacobi panel update (DMMA)                                   */
    GSI_T_FINALIZE = 5,    /* norms, sort, cutoff, emit                                          */
    GSI_T_COMPACT = 6,     /* offset scan + compaction                                           */
    GSI_T_PREDICT = 7,
    GSI_T_KNN = 8,
    GSI_T_TRD = 9,         /* Householder tridiagonalisation (large path)                        */
    GSI_T_DC = 10,         /* divide & conquer on the tridiagonal: secular/deflation kernels      */
    GSI_T_DC_GEMM = 11,    /* divide & conquer merge GEMMs (DMMA)                                 */
    GSI_T_BT = 12,         /* back-transformation of the kept eigenvectors (DMMA) + emit          */
    GSI_T_CHEBY = 13,      /* Chebyshev graph filter supersteps                                   */
    GSI_T_SBR = 14,        /* two-stage tridiagonalisation: dense -> band (DMMA), bulge chasing   */
    GSI_T_BT2 = 15,        /* back-transformation through the bulge-chasing reflectors (DMMA)     */
    GSI_T_COUNT = 16
};
int gsi_timing_enable(gsi_ctx* ctx, int on);
int gsi_timing_reset(gsi_ctx* ctx);
/* ms[c]: device time of the launches that were bracketed by events; samples[c]: how many launches
 * that time covers (the block-Jacobi rounds are sampled 1 in 8); launches[c]: all launches. */
int gsi_timing_get(gsi_ctx* ctx, double* ms /*[GSI_T_COUNT]*/, int64_t* launches /*[GSI_T_COUNT]*/,
                   int64_t* samples /*[GSI_T_COUNT]*/);
/* Measured FP64 FMA throughput of this device in TFLOP/s (register-resident DFMA chains on every
 * SM, CUDA events) -- the roofline denominator for the Jacobi kernels (MEASURED_PEAKS.json holds
 * only HBM and bf16 tensor numbers).  `use_dmma` != 0 measures mma.sync m8n8k4 f64 instead. */
int gsi_measure_fp64_tflops(gsi_ctx* ctx, int use_dmma, double* tflops);

/* ---- stage-wise test hook of the large-n eigensolver ---------------------------------------- *
 * Runs the Householder / divide-and-conquer / back-transform pipeline that gsi_precompute_* uses for
 * n > GSI large-path threshold on ONE dense symmetric matrix given by the caller (host, n x n,
 * column-major, n >= 33), and returns the intermediate results so that each stage can be compared
 * with the oracle on its own (tests/test_eigh_stages_gpu.py):
 *   d[n], e[n-1], tau[n-1]   tridiagonal T and Householder scalars;  V[n*n] reflectors (column j:
 *                            v_j with v_j[j+1] = 1, zeros above)
 *   lam[n]                   eigenvalues of T (= of A), ascending
 *   U[n*k]                   eigenvectors of A for lam <= thr (at least 2), column-major n x k
 * `team` is the number of CTAs that cooperate on the matrix (0 = the planner's choice).  Any output
 * pointer may be NULL. */
int gsi_debug_eigh(gsi_ctx* ctx, int n, const double* a, float thr, int team, double* d, double* e,
                   double* tau, double* v, double* lam, double* u, int32_t* k);

/* Test hook of the two-stage tridiagonalisation (by default the route of users with n > 9,216; GSI_SBR_MIN overrides): stage 1 (dense -> band of
 * half-width 64) on ONE dense symmetric matrix (host, column-major, n >= 130).  ab[j * 128 + dd] = B(j + dd, j), dd = 0 .. 64
 * (zero beyond); B is orthogonally similar to A.  With d / e non-NULL stage 2 (bulge chasing) runs as well and the tridiagonal
 * comes back. */
int gsi_debug_band(gsi_ctx* ctx, int n, const double* a, double* ab, double* d, double* e);

/* Test / measurement hook of the FP64-equivalent tensor-core GEMM (csrc/tc_gemm.cu: INT8 slices on tcgen05.mma.kind::i8,
 * INT32 accumulators in tensor memory; the engine behind the divide-and-conquer merges that replace the QR-iteration half of
 * precompute_local.cpp:231).  C (m x n) = A (m x k) * B (k x n), host buffers, column-major with the given leading dimensions.
 * slices = 6, 7 or 8 digits of 7 bits per operand entry.  The device part runs `reps` times (>= 1); ms_slice / ms_gemm return
 * the CUDA-event time of ONE repetition of the slicing kernels and of the tcgen05 kernel (either may be NULL). */
int gsi_debug_tc_gemm(gsi_ctx* ctx, int m, int n, int k, const double* a, int64_t lda, const double* b, int64_t ldb,
                      double* c, int64_t ldc, int slices, int reps, double* ms_slice, double* ms_gemm);

#ifdef __cplusplus
}
#endif
#endif /* GSI_H_ */
