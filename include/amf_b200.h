/*
 * amf_b200.h -- C ABI of the B200-native (sm_100a) accelerator for the python-pmf hot path
 * of autonlab/active-matrix-factorization.
 *
 * The reference has no FFI: its seam is the Python class API backed by Cython
 * (python-pmf/pmf_cy.pxd:6-36).  Each entry point below names the reference routine whose
 * inner loop it replaces (paths relative to python-pmf/).  The Python host mirror in
 * active_matrix_factorization_b200/ binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; amf_last_error() gives the
 *     message of the calling thread's last failure.  No function falls back to the CPU.
 *   - pointers named *_d are device pointers on the current CUDA device, *_h host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All work
 *     is enqueued asynchronously unless the function name ends in _host or says it syncs.
 *   - dtype: AMF_F32 (fast mode) or AMF_F64 (parity mode, the reference's precision).
 *   - factor matrices are row-major (rows, ld) with ld >= d, ld*sizeof(T) a multiple of 16
 *     bytes, and columns d..ld-1 equal to zero.
 *   - user / item ids are checked against the matrix shape once, where a structure is built
 *     from them (amf_ratings_create, amf_ratings_append, amf_pool_create: AMF_ERR_INVALID, as the
 *     reference asserts in pmf_cy.pyx:139-140); the per-step scoring calls that take raw
 *     candidate arrays (amf_score_candidates, amf_score_pred_host*, amf_mn_score_candidates,
 *     amf_bayes_sample_stats) trust them.
 *   - the library keeps no global mutable state besides per-thread error text; handles may be
 *     used from different host threads as long as one handle is not used concurrently.
 */
#ifndef AMF_B200_H
#define AMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMF_F32 0
#define AMF_F64 1

#define AMF_OK 0
#define AMF_ERR_INVALID 1
#define AMF_ERR_CUDA 2
#define AMF_ERR_UNSUPPORTED 3

const char* amf_last_error(void);
int amf_version(void);
/* number of visible CUDA devices and the compute capability (major*10+minor) of the current one;
 * fails (AMF_ERR_CUDA) when there is no usable device -- callers must not fall back. */
int amf_device_info(int* n_devices, int* sm_arch, int* n_sms);

/* ------------------------------------------------------------------------------------------
 * Rating list (replaces `ratings`, the (nnz,3) float64 array iterated row by row in
 * pmf_cy.pyx:184-186 and :217-221, and the adjacency dicts of bayes_pmf.py:241-255).
 * Device-resident, stored twice: user-major (CSR) and item-major (CSC), each entry
 * {other-side index:int32, rating:T}; order inside a row/column is the input order (stable).
 * ------------------------------------------------------------------------------------------ */
typedef struct amf_ratings amf_ratings_t;

/* i_d/j_d/r_d: device COO arrays of length nnz (r_d of `dtype`).  Duplicates are kept. */
int amf_ratings_create(amf_ratings_t** out, int32_t n_users, int32_t n_items, int64_t nnz,
                       const int32_t* i_d, const int32_t* j_d, const void* r_d, int dtype,
                       void* stream);
/* same from host arrays (copies them to the device first) */
int amf_ratings_create_host(amf_ratings_t** out, int32_t n_users, int32_t n_items, int64_t nnz,
                            const int32_t* i_h, const int32_t* j_h, const void* r_h, int dtype);
int amf_ratings_destroy(amf_ratings_t* h);
int64_t amf_ratings_nnz(const amf_ratings_t* h);
/* device pointers of the two layouts (for tests / Gibbs):  ptr int64[rows+1], idx int32[nnz],
 * val T[nnz].  side 0 = user-major, 1 = item-major. */
int amf_ratings_layout(const amf_ratings_t* h, int side, const int64_t** ptr_d,
                       const int32_t** idx_d, const void** val_d);
/* sum and count of ratings -> mean_rating (pmf_cy.pyx:63); synchronises. */
/* Appends ratings (device arrays, same value type as the list) without re-sorting: they are kept
 * as an unsorted tail that amf_pmf_loss_grad adds on top of the sorted passes, and are merged
 * into the sorted lists (as if the whole list had been created in order of arrival) when the
 * tail exceeds max(65536, nnz/32) entries, by amf_ratings_compact, or by any other consumer of
 * the sorted lists (amf_ratings_layout, amf_ratings_mean, amf_gibbs_half_sweep).  This is
 * add_rating / add_ratings (pmf_cy.pyx:128-156) for a device-resident list: an active-learning
 * step costs O(new ratings), not a re-upload and two sorts of everything. */
int amf_ratings_append(amf_ratings_t* h, int64_t n_new, const int32_t* i_d, const int32_t* j_d,
                       const void* r_d, void* stream);
int amf_ratings_compact(amf_ratings_t* h, void* stream);

/* Which copy of the list amf_pmf_loss_grad runs on.  AUTO: the tiled copy (item / user tiles of
 * the factor matrices resident in shared memory; the "bundled runs" layout: one lane per (row, tile)
 * run segment, 16-bit row inside the tile + the rating = 6 or 10 bytes per rating and side plus
 * ~20 % padding, built on first use) when nnz >= 2^20 and the padded factor row is 64, 128 or 256
 * bytes, else the row-sorted lists; ROWS / TILED force one (TILED fails with AMF_ERR_UNSUPPORTED if it cannot). */
#define AMF_LAYOUT_AUTO 0
#define AMF_LAYOUT_ROWS 1
#define AMF_LAYOUT_TILED 2
int amf_ratings_set_layout(amf_ratings_t* h, int mode);
int amf_ratings_mean(const amf_ratings_t* h, double* mean_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * PMF objective and gradient (pmf_cy.pyx:170-193 log_likelihood, :204-223 gradient,
 * :225-234 update_sigma).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  double sigma_sq, sigma_u_sq, sigma_v_sq; /* pmf_cy.pyx:40-42 */
  double mean_offset;                      /* mean_rating if subtract_mean else 0 (:165-168) */
} amf_pmf_params_t;

/* Fused loss + gradient at (U, V):
 *   sums_d[0] = sum (r - r_hat)^2, sums_d[1] = |U|^2, sums_d[2] = |V|^2   (double[3], device)
 *   dU = -U/sigma_u_sq + sum_j V_j (r - r_hat)/sigma_sq, dV symmetric     (ascent direction)
 * dU_d/dV_d may both be NULL: loss only.  log-likelihood =
 *   -sums[0]/(2 sigma_sq) - sums[1]/(2 sigma_u_sq) - sums[2]/(2 sigma_v_sq). */
int amf_pmf_loss_grad(const amf_ratings_t* h, int dtype, int d, int ld, const void* U_d,
                      const void* V_d, const amf_pmf_params_t* p, void* dU_d, void* dV_d,
                      double* sums_d, void* stream);

/* amf_pmf_loss_grad in two calls, for a caller that puts a collective between them (multi-GPU:
 * all-reduce dU while dV is still being computed).  part 0: prior terms of both sides, the pass
 * that completes dU_d and sums_d (and the appended tail, which adds to both gradients); part 1:
 * the pass that completes dV_d.  Both parts must be run, in this order, with the same arguments.
 * max_ctas > 0 caps the grid of the tiled passes (leaves SMs to the concurrent collective). */
int amf_pmf_loss_grad_part(const amf_ratings_t* h, int dtype, int d, int ld, const void* U_d,
                           const void* V_d, const amf_pmf_params_t* p, void* dU_d, void* dV_d,
                           double* sums_d, int part, int max_ctas, void* stream);

/* The whole line-search fit (pmf_cy.pyx:257-305 fit_lls consumed by fit) in ONE launch, for small
 * problems: one cooperative launch (CTAs meet at a grid barrier three times per trial) runs
 * trial point, fused objective + gradient, accept / reject, step-size update (x1.25 / x0.5 in
 * fp64) and the convergence tests (gain < stop_thresh, lr < min_lr) with the reference's control
 * flow; U_d / V_d (rows x ld, padding zero) are updated in place.
 * trace_d[0 .. min(steps, trace_cap)) receives the objective of every accepted step.  max_steps
 * <= 0: no cap.  Uses scalar gathers and atomics: meant for lists up to a few 1e5 ratings, where a
 * trial is launch-bound; larger lists should drive amf_pmf_loss_grad + amf_axpy from the host.
 * workspace_d: 256-byte aligned.  Merges an appended tail first. */
typedef struct { double lr; double ll; int32_t steps; int32_t trials; int32_t converged; int32_t pad_; } amf_fit_result_t;
int64_t amf_pmf_fit_workspace_bytes(const amf_ratings_t* h, int dtype, int ld);
int amf_pmf_fit_lls(const amf_ratings_t* h, int dtype, int d, int ld, void* U_d, void* V_d,
                    const amf_pmf_params_t* p, double lr, double min_lr, double stop_thresh,
                    int max_steps, double* trace_d, int trace_cap, amf_fit_result_t* result_d,
                    void* workspace_d, int64_t workspace_bytes, void* stream);

/* Line-search trial point of fit_lls (pmf_cy.pyx:271-272): X_new = X + lr * G, elementwise
 * over `count` elements (the padded (rows, ld) block). */
int amf_axpy(int dtype, int64_t count, const void* X_d, const void* G_d, double lr,
             void* Xnew_d, void* stream);

/* Gradient of an explicit COO mini-batch with atomics (gradient(ratings=batch),
 * pmf_cy.pyx:205,211-212 as used by fit_minibatches :335-336).  dU/dV must already hold the
 * prior term or zeros; this only adds the data term.  sums_d[0] gets the squared error. */
int amf_pmf_grad_coo(int dtype, int64_t nnz, const int32_t* i_d, const int32_t* j_d,
                     const void* r_d, int d, int ld, const void* U_d, const void* V_d,
                     const amf_pmf_params_t* p, void* dU_d, void* dV_d, double* sums_d,
                     void* stream);
/* dX = -X / sigma_x_sq and |X|^2 -> *norm2_d (added) : the prior half of the gradient */
int amf_pmf_prior(int dtype, int64_t count, const void* X_d, double sigma_x_sq, void* dX_d,
                  double* norm2_d, void* stream);

/* Momentum SGD update of fit_minibatches (pmf_cy.pyx:338-344), elementwise over `count`:
 *   inc = momentum * inc + scale * G ;  X += inc */
int amf_momentum_step(int dtype, int64_t count, void* inc_d, const void* G_d, double momentum,
                      double scale, void* X_d, void* stream);

/* End-to-end variant with HOST buffers (what ProbabilisticMatrixFactorization.gradient() /
 * .log_likelihood(users, items) call): copies U, V (n*d and m*d, tightly packed row-major, of
 * `dtype`) to the device, runs amf_pmf_loss_grad, copies dU, dV (may be NULL) and the three sums
 * back, and synchronises. */
int amf_pmf_loss_grad_host(const amf_ratings_t* h, int dtype, int d, const void* U_h,
                           const void* V_h, const amf_pmf_params_t* p, void* dU_h, void* dV_h,
                           double* sums_h);

/* ------------------------------------------------------------------------------------------
 * Candidate scoring with fused arg-best (active_pmf.py:739-770 _get_key_vals + :737 chooser;
 * criteria :416-421 pred, :432-439 _prob_ge_cutoff, :392-400 approx_pred_mean_var,
 * :502-524 pred_variance with normal_exps_cy.pyx:111-135 exp_dotprod_sq).
 * ------------------------------------------------------------------------------------------ */
#define AMF_CRIT_PRED 0          /* U_i . V_j (MAP)                                         */
#define AMF_CRIT_APPROX_MEAN 1   /* E[U_i . V_j] under the normal approximation            */
#define AMF_CRIT_PRED_VARIANCE 2 /* Var[U_i . V_j] under the normal approximation          */
#define AMF_CRIT_PROB_GE 3       /* norm.sf(cutoff, loc=E, scale=Var)  (reference quirk)    */

/* Strided view of the Gaussian approximation N(mean, cov) over all factors.  Element (k,l) of
 * the d x d block A_i = Cov(U_i, U_i) is cov_uu[i*uu_stride + k*uu_ld + l]; B_j likewise;
 * C_ij[k,l] = Cov(U_ki, V_lj) = cov_uv[i*uv_stride_i + j*uv_stride_j + k*uv_ld + l] or
 * cov_uv == NULL for a block-diagonal posterior.  With the reference's full k x k matrix
 * (k=(n+m)d, active_pmf.py:136-142): mean_u=mean, mean_v=mean+n*d, cov_uu=cov,
 * uu_stride=d*k+d, uu_ld=k, cov_vv=cov+n*d*k+n*d, cov_uv=cov+n*d, uv_stride_i=d*k,
 * uv_stride_j=d, uv_ld=k.  All of `dtype`. */
typedef struct {
  const void* mean_u; int64_t mean_u_stride;
  const void* mean_v; int64_t mean_v_stride;
  const void* cov_uu; int64_t uu_stride, uu_ld;
  const void* cov_vv; int64_t vv_stride, vv_ld;
  const void* cov_uv; int64_t uv_stride_i, uv_stride_j, uv_ld;
} amf_normal_view_t;

/* Scores ncand candidates (ci_d[c], cj_d[c]).  scores_d (T[ncand]) may be NULL when only the
 * winner is wanted.  best_d: device record {double value; int64 index} of the max
 * (maximize != 0) or min, lowest candidate index on ties, index -1 if ncand == 0 or all NaN.
 * For AMF_CRIT_PRED U_d/V_d/ld are used; for the others `nv`.  index_base is added to the
 * reported index (shard offset on multi-GPU runs). */
typedef struct { double value; int64_t index; } amf_best_t;
int amf_score_candidates(int criterion, int dtype, int64_t ncand, const int32_t* ci_d,
                         const int32_t* cj_d, int d, int ld, const void* U_d, const void* V_d,
                         const amf_normal_view_t* nv, double cutoff, void* scores_d,
                         int maximize, int64_t index_base, amf_best_t* best_d, void* stream);

/* Reduces n winner records (e.g. the all-gathered per-GPU winners of a sharded pool) to one with
 * the rule of amf_score_candidates: best value, lowest index on ties, records with index < 0 or a
 * NaN value never win; out = {0, -1} if none does.  recs_d and out_d may not alias. */
int amf_best_reduce(const amf_best_t* recs_d, int n, int maximize, amf_best_t* out_d, void* stream);

/* approx_pred_covs (active_pmf.py:324-390) for `count` Gaussian approximations at once: mean_d
 * (count, k), cov_d (count, k, k), k = (n + m) d, the reference's layout (active_pmf.py:136-142);
 * out_d (count, n*m, n*m) = covariance between all pairs of predicted cells (Isserlis).  Exact mode
 * sizes only. */
int amf_pred_covs(int32_t n, int32_t m, int d, int count, const double* mean_d,
                  const double* cov_d, double* out_d, void* stream);
/* np.linalg.slogdet of `count` k x k matrices a_d (count, k, k; overwritten by their LU factors):
 * LU with partial pivoting, one CTA per matrix; sign_d in {-1, 0, 1}, logdet_d = log|det| (-inf
 * for a singular matrix).  The log det of _approx_entropy and _pred_entropy_bound
 * (active_pmf.py:526-530, 559-574). */
int amf_slogdet_batched(int k, int count, double* a_d, int* sign_d, double* logdet_d, void* stream);

/* predicted_matrix (pmf_cy.pyx:410-420): out_d (n, m) row-major = U V' + offset, in the compute type. */
int amf_predicted_matrix(int dtype, int32_t n, int32_t m, int d, int ld, const void* U_d,
                         const void* V_d, double offset, void* out_d, void* stream);
/* The sums behind rmse / rmse_on (pmf_cy.pyx:25-29, 422-426) without the N x M matrix:
 * sums_d[0] = sum over the selected cells of (real_d[i*m+j] - U_i.V_j - offset)^2, sums_d[1] = their
 * number; mask_d (n*m bytes, nonzero = selected) may be NULL for all cells; real_d is fp64. */
int amf_sq_error_dense(int dtype, int32_t n, int32_t m, int d, int ld, const void* U_d,
                       const void* V_d, double offset, const double* real_d,
                       const unsigned char* mask_d, double* sums_d, void* stream);

/* Winner exchange of a sharded pool over NVLink peer memory (one process per GPU, one node): the
 * all-gather of the per-GPU winners + amf_best_reduce as ONE kernel per rank -- remote stores of the
 * 16-byte record into every peer's mailbox, a release flag, an acquire wait for all peers' flags,
 * the reduction (`chooser` over the Pool.map chunks, active_pmf.py:765-770, on several GPUs).
 *   amf_peer_create   allocates this rank's mailbox and returns its 64-byte CUDA IPC handle;
 *   amf_peer_connect  takes the handles of all ranks (world x 64 bytes, rank order) and maps them
 *                     (AMF_ERR_UNSUPPORTED if a peer cannot be mapped: keep the NCCL path);
 *   amf_peer_best_reduce  mine_d -> out_d (may alias); a collective: every rank calls it the same
 *                     number of times in the same order; a peer that never arrives traps after ~1 s. */
typedef struct amf_peer amf_peer_t;
int amf_peer_create(amf_peer_t** out, int world, int rank, unsigned char* ipc_handle_out);
int amf_peer_connect(amf_peer_t* p, const unsigned char* all_handles);
int amf_peer_best_reduce(amf_peer_t* p, const amf_best_t* mine_d, int maximize, amf_best_t* out_d,
                         void* stream);
int amf_peer_destroy(amf_peer_t* p);

/* Candidate pool handle: the pool bucketed once by item tile (tile_rows items) in the "bundled
 * runs" layout: the candidates one user has inside one tile are a run, one lane of the scoring
 * kernel owns a run segment (<= 64 candidates) with the whole user row in registers, and 32
 * equally long segments form the bundle a warp works on.  A candidate is stored as its 16-bit row
 * inside the tile (tile_rows <= 65535).  The kernel keeps the tile of V resident in shared memory
 * (TMA bulk copies) and reads one item row per candidate from there -- one shared-memory wavefront
 * per candidate at d = 32 fp32; built from the caller's (i, j) arrays, remembers the caller's
 * order.  Replaces the `pool` list / `unrated` set iterated in active_pmf.py:725-770 when the same
 * pool is scored repeatedly (every active-learning step).
 * amf_pool_max_tile_rows: the tallest tile the scoring kernel can hold for padded factor rows of
 * row_bytes (16, 32, 64, 128 or 256; 0 = width not supported by the pool kernel). */
int amf_pool_max_tile_rows(int row_bytes);
typedef struct amf_pool amf_pool_t;
int amf_pool_create(amf_pool_t** out, int64_t ncand, const int32_t* ci_d, const int32_t* cj_d,
                    int32_t n_users, int32_t n_items, int tile_rows, void* stream);
int amf_pool_destroy(amf_pool_t* h);
int64_t amf_pool_size(const amf_pool_t* h);
/* Removes candidates (given by their positions in the caller's order) from the pool in O(1)
 * each: they no longer compete for the winner and their score slots are left untouched.  This
 * is `unrated.difference_update(new_items)` (pmf_cy.pyx:152) for a device-resident pool: the
 * active loop queries one candidate per step and keeps scoring the rest. */
int amf_pool_remove(amf_pool_t* h, int64_t n, const int64_t* idx_d, void* stream);
/* AMF_CRIT_PRED over the pool: scores_d (T[ncand], caller's order) may be NULL; best_d as in
 * amf_score_candidates (index = position in the caller's order + index_base, lowest wins ties).
 * The pool holds the work counters and the score scratch of the launch: calls on one pool must be
 * stream-ordered (one at a time); distinct pools may be scored concurrently. */
int amf_pool_score_pred(const amf_pool_t* h, int dtype, int d, int ld, const void* U_d,
                        const void* V_d, void* scores_d, int maximize, int64_t index_base,
                        amf_best_t* best_d, void* stream);

/* The same with the cross-GPU winner exchange fused into the scoring kernel (sharded pools, one
 * process per GPU): the last CTA to finish reduces this GPU's partial winners and exchanges the
 * record with the other GPUs over NVLink peer memory (amf_peer_*), so that best_d holds the winner
 * over ALL shards when the kernel ends -- scoring, arg-best and the "all-gather + chooser" of
 * active_pmf.py:765-770 in ONE launch.  A collective: every rank calls it in the same order. */
int amf_pool_score_pred_peer(const amf_pool_t* h, int dtype, int d, int ld, const void* U_d,
                             const void* V_d, void* scores_d, int maximize, int64_t index_base,
                             amf_peer_t* peer, amf_best_t* best_d, void* stream);

/* End-to-end host variant of AMF_CRIT_PRED: host candidate arrays and factors in, scores
 * (may be NULL) and winner out; synchronises. */
int amf_score_pred_host(int dtype, int64_t ncand, const int32_t* ci_h, const int32_t* cj_h,
                        int32_t n, int32_t m, int d, const void* U_h, const void* V_h,
                        void* scores_h, int maximize, amf_best_t* best_h);

/* Same, for a pool sorted by user and given as row offsets: the candidates of user i are
 * cj_h[cand_ptr_h[i] .. cand_ptr_h[i+1]) (n+1 offsets, cand_ptr_h[0] = 0, cand_ptr_h[n] = ncand),
 * candidate index = position in cj_h.  Half the host->device bytes of the (ci, cj) form. */
int amf_score_pred_host_csr(int dtype, const int64_t* cand_ptr_h, const int32_t* cj_h, int32_t n,
                            int32_t m, int d, const void* U_h, const void* V_h, void* scores_h,
                            int maximize, amf_best_t* best_h);

/* Same with 16-bit item ids, for pools over at most 65536 items (m <= 65536, else an error):
 * 2 bytes per candidate cross PCIe, widened on the device piece by piece behind the copy. */
int amf_score_pred_host_csr16(int dtype, const int64_t* cand_ptr_h, const uint16_t* cj16_h,
                              int32_t n, int32_t m, int d, const void* U_h, const void* V_h,
                              void* scores_h, int maximize, amf_best_t* best_h);

/* ------------------------------------------------------------------------------------------
 * Bayesian PMF (bayes_pmf.py:189-216 sample_feature inside the sweeps of :283-300;
 * :433-455 predict / pred_variance / :528-538 prob_ge_cutoff over a list of samples).
 * ------------------------------------------------------------------------------------------ */
/* One half-sweep: for every row n of the side being sampled
 *   Lambda = alpha + beta F'F,  cov = Lambda^-1,  mean = cov (beta F'(r - mean_offset) + alpha mu),
 *   out[n] = chol(cov) z[n] + mean            (lower Cholesky factor, like np.linalg.cholesky)
 * where F = other[idx of row n].  side 0 samples users (other = items), 1 samples items.
 * z_d: (rows, d) standard normals in the reference's draw order; alpha_d (d,d), mu_d (d).
 * other_d/out_d are tightly packed (rows, d).  Always computed in fp64 when dtype==AMF_F64. */
int amf_gibbs_half_sweep(const amf_ratings_t* h, int side, int dtype, int d, const void* other_d,
                         const void* alpha_d, const void* mu_d, double beta, double mean_offset,
                         const void* z_d, void* out_d, void* stream);

/* Same for rows [row_begin, row_end) only (row_end < 0: to the last row): the multi-GPU split of
 * a half-sweep; z_d and out_d are still indexed by absolute row. */
int amf_gibbs_half_sweep_rows(const amf_ratings_t* h, int side, int dtype, int d,
                              const void* other_d, const void* alpha_d, const void* mu_d,
                              double beta, double mean_offset, const void* z_d, void* out_d,
                              int32_t row_begin, int32_t row_end, void* stream);

/* Fast mode of the half-sweep: the standard normals are generated inside the kernel (Philox4x32-10,
 * counter = (row, component, stream_id), key = seed: stateless, any row order, any number of
 * GPUs) and the row is sampled from ONE Cholesky factor of the precision,
 *   Lambda = R R',  out[n] = R^-T (R^-1 rhs + z)  ~  N(Lambda^-1 rhs, Lambda^-1),
 * instead of chol(inv(Lambda)) (bayes_pmf.py:208-216): same conditional distribution, different
 * map from z to the sample, so chains are equal in law to the reference's, not draw for draw.
 * Use a fresh stream_id for every half-sweep of a chain. */
int amf_gibbs_half_sweep_device_rng(const amf_ratings_t* h, int side, int dtype, int d,
                                    const void* other_d, const void* alpha_d, const void* mu_d,
                                    double beta, double mean_offset, uint64_t seed,
                                    uint64_t stream_id, void* out_d, int32_t row_begin,
                                    int32_t row_end, void* stream);
/* Fast mode of sample_hyperparam (bayes_pmf.py:158-186 with sample_wishart :41-59) on the device:
 * mean and covariance of the factor rows feats_d (rows, d; tightly packed), the Normal-Wishart
 * posterior and one draw of it -- mu_out_d (d) and alpha_out_d (d, d) in the compute type, ready
 * for amf_gibbs_half_sweep_device_rng -- without a host round trip.  prior_d: device doubles
 * [inv(W0) (d*d, row-major) | mu0 (d) | beta0 | dof0].  Same distributions as the reference's host
 * code (including the scalar np.dot of bayes_pmf.py:176), Philox counters keyed by (seed,
 * stream_id): use a stream_id no half-sweep of the chain uses.  A scale matrix that is not positive
 * definite raises the handle's sticky failure flag (amf_gibbs_status).  The moment sums live in
 * the handle: draws on one handle must be stream-ordered (one chain per handle).  d <= 32, rows >= 2. */
int amf_gibbs_hyper_device(const amf_ratings_t* h, int dtype, int d, int64_t rows,
                           const void* feats_d, const double* prior_d, uint64_t seed,
                           uint64_t stream_id, void* mu_out_d, void* alpha_out_d, void* stream);
/* n_samples consecutive samples of the fast-mode chain (BayesianPMF.samples, bayes_pmf.py:227-302,
 * in law) enqueued by ONE call: per sample the Normal-Wishart draws of both sides
 * (amf_gibbs_hyper_device) and num_gibbs rounds of the two half-sweeps
 * (amf_gibbs_half_sweep_device_rng); nothing is read back, the host does not take part.  The chain
 * starts from users_d (n, d) / items_d (m, d); sample s is written to out_users_d[s] (n, d) and
 * out_items_d[s] (m, d) (the last one is the state to continue from).  Counters: hyper draws use
 * (1 << 40) + id, half-sweeps id, with id running from stream_id0 by 2 + 2 num_gibbs per sample --
 * the sequence a host loop over the two entry points above makes, so both give the same chain. */
int amf_gibbs_chain_device(const amf_ratings_t* h, int dtype, int d, int n_samples, int num_gibbs,
                           const void* users_d, const void* items_d, const double* prior_u_d,
                           const double* prior_v_d, double beta, double mean_offset, uint64_t seed,
                           uint64_t stream_id0, void* out_users_d, void* out_items_d, void* stream);
/* `count` chains over the SAME rating list in one launch, chain p with one extra rating
 * (ex_row_d[p] = its row on the side being sampled, ex_col_d[p] = the row of `other` it pairs
 * with, ex_val_d[p] = its value; all NULL: none) -- the per-candidate, per-value models of the
 * Bayesian lookahead (bayes_pmf.py:560-598: deepcopy + add_rating + a fresh chain each) without
 * copies of the list.  other_d (count, other_rows, d), alpha_d (count, d, d), mu_d (count, d),
 * out_d (count, rows, d); offsets_d (count) per-chain mean offsets or NULL for `mean_offset`.
 * Device random numbers as in amf_gibbs_half_sweep_device_rng, chain p on its own counters;
 * d <= 32. */
int amf_gibbs_half_sweep_batched(const amf_ratings_t* h, int side, int dtype, int d, int count,
                                 const void* other_d, const void* alpha_d, const void* mu_d,
                                 double beta, double mean_offset, const double* offsets_d,
                                 const int32_t* ex_row_d, const int32_t* ex_col_d,
                                 const double* ex_val_d, uint64_t seed, uint64_t stream_id,
                                 void* out_d, void* stream);
/* The same normals as an array: out_d[(row, k)] = z of (seed, stream_id, row, k). */
int amf_philox_normal(int dtype, uint64_t seed, uint64_t stream_id, int64_t rows, int d,
                      void* out_d, void* stream);

/* 1 in *failed if any row of ANY half-sweep on this handle since the previous call of this
 * function met a non-positive-definite precision/covariance (np.linalg.cholesky would have
 * raised LinAlgError).  The flag is sticky: half-sweeps only set it, this call reads and
 * clears it, so one status call per chain step covers all its half-sweeps; synchronises. */
int amf_gibbs_status(const amf_ratings_t* h, int* failed, void* stream);

/* Sample statistics of S posterior samples at ncand cells: Us_d (S, n, d), Vs_d (S, m, d)
 * tightly packed.  Any of mean_d / var_d / prob_d may be NULL.  var is the population variance
 * (np.var, bayes_pmf.py:448); prob = fraction of samples with prediction >= cutoff.
 * best_d (nullable) selects on `select` (0 mean, 1 var, 2 prob).
 * Dense form: ci_d = cj_d = NULL and ncand = n*m scores every cell (c = i*m + j, outputs of n*m
 * entries) as a blocked product over shared-memory tiles -- the form to use when `which` is the
 * whole matrix or most of it (then gather the wanted cells from the dense outputs). */
int amf_bayes_sample_stats(int dtype, int64_t ncand, const int32_t* ci_d, const int32_t* cj_d,
                           int S, int32_t n, int32_t m, int d, const void* Us_d, const void* Vs_d,
                           double mean_offset, double cutoff, void* mean_d, void* var_d,
                           void* prob_d, int select, int maximize, int64_t index_base,
                           amf_best_t* best_d, void* stream);

/* The dense form of amf_bayes_sample_stats (mean / variance of every cell over S samples, fp32)
 * on the tensor cores: per sample one 128 x 128 tile product U_s V_s^T per CTA issued with
 * tcgen05.mma (TF32 inputs split hi + lo, three products: fp32-class accuracy; fp32 accumulators
 * in tensor memory), operands by 2-D TMA, the variance accumulated about the first sample's
 * prediction, which is subtracted on the tensor cores (negated-A products into the same
 * accumulator).  amf_bayes_sample_stats routes dense fp32 calls without a prob output here
 * (AMF_B200_DENSE_TC=0 in the environment keeps them on the CUDA-core kernel).  d <= 32, S >= 2,
 * select 0 (mean) or 1 (variance). */
int amf_bayes_sample_stats_dense_tc(int S, int32_t n, int32_t m, int d, const float* Us_d,
                                    const float* Vs_d, double mean_offset, float* mean_d,
                                    float* var_d, int select, int maximize, int64_t index_base,
                                    amf_best_t* best_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Variational full-covariance approximation ("exact mode"), batched, fp64:
 * active_pmf.py:202-240 kl_divergence, normal_exps_cy.pyx:140-303 normal_gradient,
 * active_pmf.py:36-50 project_psd, :251-288 fit_normal_kls, and the lookahead of :635-704
 * (_exp_with_rij: one re-fit per candidate and rating value), whose per-problem criteria
 * :526-530 _approx_entropy and :605-606 _total_variance are computed in the same launch.
 * B independent problems share the COO rating list (ri,rj,rr) and each may append one extra
 * rating (extra_i[b] < 0: none).  One CTA per problem; k = (n+m)*d.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t n, m, d;
  double sigma_sq, sigma_u_sq, sigma_v_sq;
  double learning_rate; /* normal_learning_rate, active_pmf.py:145 (1e-4) */
  double min_eig;       /* active_pmf.py:146 (1e-5) */
  double kl_stop;       /* convergence threshold on the KL decrease, active_pmf.py:277 (.005) */
  double min_lr;        /* active_pmf.py:286 (1e-10) */
  int32_t max_steps;    /* <= 0: run to convergence */
} amf_normal_fit_params_t;

#define AMF_NORMAL_FIT 0      /* fit_normal_kls on every problem; mean/cov updated in place  */
#define AMF_NORMAL_KL 1       /* kl_out[b] = KL(mean_b, cov_b)                               */
#define AMF_NORMAL_GRADIENT 2 /* work_b[0:k] = dKL/dmean, work_b[2k : 2k+k*k] = dKL/dcov     */
#define AMF_NORMAL_PROJECT 3  /* cov_b <- project_psd(cov_b, min_eig)                        */

/* doubles of scratch needed PER PROBLEM in work_d */
int64_t amf_normal_workspace_doubles(int32_t n, int32_t m, int d);

/* mean_d (B,k), cov_d (B,k,k) in/out; work_d (B, workspace) scratch/out; kl_out_d (B);
 * steps_out_d (B) accepted steps; kl_trace_d (B, trace_len) KL after each accepted step or NULL;
 * entropy_out_d / totvar_out_d (B) or NULL: log det cov / sum_ij Var[Ui.Vj] after the fit. */
int amf_normal_batched(int mode, int B, int64_t nnz, const int32_t* ri_d, const int32_t* rj_d,
                       const double* rr_d, const int32_t* extra_i_d, const int32_t* extra_j_d,
                       const double* extra_r_d, const amf_normal_fit_params_t* p, double* mean_d,
                       double* cov_d, double* work_d, double* kl_out_d, int32_t* steps_out_d,
                       double* kl_trace_d, int trace_len, double* entropy_out_d,
                       double* totvar_out_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Scalable-mode variational posterior ("blocks"): the KL objective of active_pmf.py:202-240
 * restricted to block-diagonal covariances -- one d x d block per user row and per item column
 * instead of the k x k matrix of active_pmf.py:136,190-200 (k = (N+M) d: 2595 at the drugbank
 * configuration, 26,250 at movielens-100k) -- and the lookahead of active_pmf.py:635-704
 * (_exp_with_rij with _approx_entropy :526-530 / _total_variance :605-606) by local re-fits.
 * All tables fp64, tightly packed: mean (rows, d), cov / prec (rows, d, d), h = prec @ mean
 * (rows, d), logdet = log det cov (rows).  d <= 32.
 * ------------------------------------------------------------------------------------------ */
/* One coordinate half-sweep of the block-restricted objective over every row of `side`
 * (0 users, 1 items): Lambda_i = I/prior_var + sum_{j rated by i} (n_j n_j^T + B_j)/sigma_sq,
 * h_i = sum_j (r_ij - mean_offset) n_j / sigma_sq, A_i = Lambda_i^-1, m_i = A_i h_i, where
 * (n_j, B_j) = (other_mean_d, other_cov_d) is the other side's posterior.  other_cov_d == NULL
 * drops the B_j term (curvature of the MAP objective); mean_d == NULL leaves the means alone.
 * *fail_d (nullable, device) is set to 1 if a precision is not positive definite. */
int amf_blocks_half_sweep(const amf_ratings_t* h, int side, int d, const double* other_mean_d,
                          const double* other_cov_d, double prior_var, double sigma_sq,
                          double mean_offset, double* prec_d, double* h_d, double* cov_d,
                          double* mean_d, double* logdet_d, int* fail_d, void* stream);

/* fit_normal on the block family (active_pmf.py:251-288 restricted to blocks): full sweeps (all
 * user rows given the items' posterior, then all item columns) until no mean moves by more than
 * `tol` or max_sweeps are done -- the loop of amf_blocks_half_sweep calls, with the largest move of
 * a mean and the failure flag read once per sweep, run by the library.  The means must hold the
 * starting point; use_cov_term = 0 drops the B_j / A_i terms.  *sweeps_done = sweeps run; *fail_d
 * (device) is left set if a precision was not positive definite. */
int amf_blocks_fit(const amf_ratings_t* h, int d, double sigma_u_sq, double sigma_v_sq,
                   double sigma_sq, double mean_offset, int use_cov_term, int max_sweeps, double tol,
                   double* mean_u_d, double* cov_u_d, double* prec_u_d, double* h_u_d,
                   double* logdet_u_d, double* mean_v_d, double* cov_v_d, double* prec_v_d,
                   double* h_v_d, double* logdet_v_d, int* fail_d, int* sweeps_done, void* stream);

/* out_d[0 : d*d] = sum_i cov_i, out_d[d*d : 2 d*d] = sum_i mean_i mean_i^T over one side: the
 * d x d sums that _total_variance (active_pmf.py:605-606) is a bilinear form of. */
int amf_blocks_sums(int64_t rows, int d, const double* mean_d, const double* cov_d, double* out_d,
                    void* stream);

typedef struct {
  int32_t n, m, d;
  const double *mean_u, *cov_u, *prec_u, *h_u, *logdet_u;
  const double *mean_v, *cov_v, *prec_v, *h_v, *logdet_v;
  const double* sums;  /* 4 d*d: sum A_i, sum m_i m_i^T, sum B_j, sum n_j n_j^T (total variance) */
  double sigma_sq;
  double entropy0;     /* log det of the whole covariance = sum of all logdet entries */
} amf_blocks_view_t;

#define AMF_LOOK_ENTROPY 0         /* _approx_entropy of the re-fitted model          */
#define AMF_LOOK_TOTAL_VARIANCE 1  /* _total_variance of the re-fitted model          */
#define AMF_WEIGHTS_NONE 0         /* raw evaluations only (evals_d)                   */
#define AMF_WEIGHTS_DISCRETE 1     /* Delta-cdf of N(rij_mean, rij_sd^2) at nv+1 bounds (:687-689) */
#define AMF_WEIGHTS_NODES 2        /* v_q = rij_mean + rij_sd * values[q], given weights (:691-699) */

/* E_v[fn(model + (i, j, v))] for ncand candidates, one lane group per candidate: for each of the
 * nv values `rounds` x (row i given column j, column j given row i) -- a rank-d update of each
 * d x d precision, a Cholesky and (where needed) an inverse per update -- then the criterion
 * and the expectation.  bounds_or_weights_d: nv+1 bounds (first/last ignored = -inf/+inf) in
 * DISCRETE mode, nv weights in NODES mode.  evals_d (ncand, nv) and scores_d (ncand) may be
 * NULL; best_d as in amf_score_candidates (over scores; meaningless for AMF_WEIGHTS_NONE). */
int amf_blocks_lookahead(const amf_blocks_view_t* v, int what, int rounds, int64_t ncand,
                         const int32_t* ci_d, const int32_t* cj_d, int nv, const double* values_d,
                         int weight_mode, const double* bounds_or_weights_d,
                         const double* rij_mean_d, const double* rij_sd_d, double* evals_d,
                         double* scores_d, int maximize, int64_t index_base, amf_best_t* best_d,
                         int* fail_d, void* stream);

/* pred_variance under the block posterior is a dot product (active_pmf.py:502-524 with a zero
 * cross block): Var_ij = <A_i, B_j + n_j n_j^T> + <m_i m_i^T, B_j>.  Packs one side's rows into
 * d(d+1) numbers each (symmetric halves; user side [vech2 A ; vech2 m m^T] with doubled
 * off-diagonals, item side [vech(B + n n^T) ; vech B]), zero-padded to ld_out, of `dtype` --
 * amf_score_candidates / amf_pool_score_pred with AMF_CRIT_PRED on the packed tables
 * (d = d(d+1)) then scores the variance criterion with the same SDDMM kernels as `pred`. */
int amf_blocks_pack(int dtype, int64_t rows, int d, const double* mean_d, const double* cov_d,
                    int side, int ld_out, void* out_d, void* stream);

/* norm.sf(cutoff, loc = mean, scale = var) elementwise (the reference passes the variance as the
 * scale, active_pmf.py:438-439) with the fused arg-best; out_d may be NULL. */
int amf_prob_ge(int dtype, int64_t n, const void* mean_d, const void* var_d, double cutoff,
                void* out_d, int maximize, int64_t index_base, amf_best_t* best_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Matrix-normal approximation MN(mean, Sigma, Omega) (the variant the
 * reference's drugbank / movielens runs use): matrix_normal_exps_cy.pyx:159-216
 * mn_kl_divergence, :219-485 matrixnormal_gradient, mn_active_pmf.py:242-288 fit_normal_kls,
 * :513-521 _approx_entropy, :597-598 _total_variance and the lookahead of :627-697.
 * mean (B, N+M, d), sig (B, N+M, N+M), om (B, d, d), all fp64; same batching and parameter
 * struct as amf_normal_batched.  Modes: AMF_NORMAL_FIT, AMF_NORMAL_KL, AMF_NORMAL_GRADIENT
 * (gradient outputs in work_b: d/dmean at [0, nui*d), d/dSigma at [2*nui*d, +nui*nui),
 * d/dOmega at [2*nui*d + 5*nui*nui, +d*d), nui = N+M).  Modes AMF_MN_KL_SPARSE (4) and
 * AMF_MN_GRADIENT_SPARSE (5) leave out the log-det / inverse terms so that a caller with one
 * large problem can do that dense algebra grid-wide (cuSOLVER) instead of inside one CTA.
 * ------------------------------------------------------------------------------------------ */
#define AMF_MN_KL_SPARSE 4
#define AMF_MN_GRADIENT_SPARSE 5
int64_t amf_mn_workspace_doubles(int32_t n, int32_t m, int d);
int amf_mn_batched(int mode, int B, int64_t nnz, const int32_t* ri_d, const int32_t* rj_d,
                   const double* rr_d, const int32_t* extra_i_d, const int32_t* extra_j_d,
                   const double* extra_r_d, const amf_normal_fit_params_t* p, double* mean_d,
                   double* sig_d, double* om_d, double* work_d, double* kl_out_d,
                   int32_t* steps_out_d, double* kl_trace_d, int trace_len, double* entropy_out_d,
                   double* totvar_out_d, void* stream);
/* AMF_CRIT_APPROX_MEAN / _PRED_VARIANCE / _PROB_GE under the matrix-normal approximation
 * (mn_active_pmf.py:300-315, :431-438, :505-511): per candidate only Sigma[i,i], Sigma[j,j],
 * Sigma[i,j], the two mean rows and Omega are read. */
int amf_mn_score_candidates(int criterion, int dtype, int64_t ncand, const int32_t* ci_d,
                            const int32_t* cj_d, int32_t n, int32_t m, int d, const void* mean_d,
                            const void* sig_d, const void* om_d, double cutoff, void* scores_d,
                            int maximize, int64_t index_base, amf_best_t* best_d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMF_B200_H */
