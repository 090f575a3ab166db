/*
 * safeopt_b200 -- C ABI of the B200-native SafeOpt hot path.
 *
 * The reference (befelix/SafeOpt) has no FFI: its seam is Python duck typing on a GPy
 * model.  Each entry point below names the reference call it
 * stands in for (paths relative to /root/reference).  All functions return an int
 * status (SO_OK == 0, negative = error); no C++ exception crosses this boundary.
 *
 * Ownership: the caller owns every buffer passed in (device buffers are typically
 * torch tensors' data_ptr()); the library owns only the opaque handle and the fit
 * state / workspaces hanging off it.  Hot calls (so_posterior_*, so_sets_*,
 * so_expander_*, so_swarm_*) never allocate and are asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).
 * A handle is bound to one device and is not thread-safe.
 *
 * Pointer naming: `_h` = host memory, `_d` = device memory on the handle's device.
 */
#ifndef SAFEOPT_B200_H
#define SAFEOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_ABI_VERSION 4

/* status codes */
#define SO_OK                  0
#define SO_ERR_BAD_ARG        -1
#define SO_ERR_UNSUPPORTED    -2   /* kernel family / shape the device path does not implement */
#define SO_ERR_NOT_PD         -3   /* Cholesky of K + (noise+1e-8) I failed */
#define SO_ERR_CUDA           -4
#define SO_ERR_NOT_FITTED     -5
#define SO_ERR_CAPACITY       -6
#define SO_ERR_NO_DEVICE      -7
#define SO_ERR_TIMEOUT        -8   /* a peer rank never published its record (cross-rank exchange) */

/* stationary kernel families (GPy class names RBF / Matern32 / Matern52) */
#define SO_KERNEL_RBF          0
#define SO_KERNEL_MATERN32     1
#define SO_KERNEL_MATERN52     2

/* how so_posterior_* combines its safe bit into S (gp_opt.py:481 is an AND over GPs) */
#define SO_SAFE_NONE           0   /* do not touch S */
#define SO_SAFE_WRITE          1   /* S[row]  = (l > fmin) */
#define SO_SAFE_AND            2   /* S[row] &= (l > fmin) */

/* swarm fitness kinds (gp_opt.py:901-1013 `swarm_type`) */
#define SO_SWARM_GREEDY        0
#define SO_SWARM_MAXIMIZERS    1
#define SO_SWARM_EXPANDERS     2
#define SO_SWARM_SAFE_SET      3

typedef struct so_handle so_handle;

/* Record produced by so_sets_reduce_safe: everything gp_opt.py:504-512, :634-636 and
 * :705-712 need from one streaming pass.  Rows are GLOBAL row indices (row0 + local). */
typedef struct so_safe_record {
    int64_t n_safe;        /* count_nonzero(S)                                  */
    double  max_l0;        /* max(Q[S,0])  (-inf if none)                       */
    int64_t argmax_l0;     /* first row attaining it (-1 if none)               */
    double  max_u0;        /* max(Q[S,1])                                       */
    int64_t argmax_u0;     /* first row attaining it (ucb query, gp_opt.py:635) */
    int64_t reserved[3];
} so_safe_record;

/* Record produced by so_sets_maximizers (gp_opt.py:511-513 and the M part of :642-644). */
typedef struct so_max_record {
    int64_t n_max;         /* count_nonzero(M)                                            */
    double  max_width0;    /* max(u0[M]-l0[M])   (unscaled; host divides by scaling[0])   */
    double  best_value;    /* max over M of max_i((u_i-l_i)/scaling_i)                    */
    int64_t best_row;      /* first row attaining best_value (-1 if none)                 */
    int64_t reserved[4];
} so_max_record;

/* ------------------------------------------------------------------ lifetime */
int         so_abi_version(void);
const char* so_status_string(int status);
/* Create a context on CUDA device `device` able to hold `max_gps` fitted GPs. */
int         so_create(int device, int max_gps, so_handle** out);
int         so_destroy(so_handle* h);
/* Human-readable detail of the last failing call on this handle ("" if none). */
const char* so_last_error(const so_handle* h);
int         so_num_sms(const so_handle* h);

/* ------------------------------------------------------------------ K1: fit
 * Stands in for GPy `set_XY` -> ExactGaussianInference.inference, reached from
 * safeopt/gp_opt.py:227, :267, :275 (and the constructor).  Builds
 * Ky = K(X,X) + (noise_var + 1e-8) I, its lower Cholesky L, L^-1 (packed for the
 * tensor-core contraction) and alpha = Ky^-1 Y, all in fp64 on the device.
 *   X_h (N x d row-major), Y_h (N), lengthscale_h (d entries; a non-ARD kernel passes
 *   its single lengthscale replicated d times).
 * Synchronises `stream` before returning so that SO_ERR_NOT_PD can be reported. */
int so_fit(so_handle* h, int gp, const double* X_h, const double* Y_h, int N, int d,
           int kernel_kind, const double* lengthscale_h, double variance, double noise_var,
           void* stream);
/* Asynchronous variant for N <= 512 (the one-launch cluster fit): returns right after the launch, so the host can queue the
 * table build and the posterior behind it while the factorisation runs.  The kernel leaves its status in mapped pinned
 * memory; so_fit_status returns it (SO_OK / SO_ERR_NOT_PD) and must be called after the caller's next synchronisation of
 * `stream` (safeopt_b200 does so when the records of the set pass arrive).  Larger N behaves like so_fit. */
int so_fit_async(so_handle* h, int gp, const double* X_h, const double* Y_h, int N, int d,
                 int kernel_kind, const double* lengthscale_h, double variance, double noise_var,
                 void* stream);
int so_fit_status(so_handle* h, int gp);
/* f4 -- one-point updates of an existing fit (same hyper-parameters), O(N^2) instead of O(N^3):
 * so_fit_append stands in for the `set_XY(vstack(X, x), vstack(Y, y))` of
 * safeopt/gp_opt.py:227 (`_add_data_point`, reached from add_new_data_point :230-255):
 * bordered Cholesky row l = L^-1 k_new, new row of L^-1, alpha and z updated in place, one
 * fragment block row re-packed.  x_new_h: d doubles.  Returns SO_ERR_CAPACITY when the
 * handle's buffers are full and SO_ERR_NOT_PD when the bordered matrix is not positive
 * definite -- in both cases the caller falls back to so_fit on the full data.
 * so_fit_remove_last stands in for `set_XY(X[:-1], Y[:-1])` (gp_opt.py:267, :275).
 * After either call so_grid_prepare must run again (like after so_fit). */
int so_fit_append(so_handle* h, int gp, const double* x_new_h, double y_new, void* stream);
int so_fit_remove_last(so_handle* h, int gp, void* stream);
/* Fit of a GP that shares inputs, kernel and noise with the already fitted `src_gp` (SafeOpt's GPs share their inputs,
 * safeopt/gp_opt.py:121-130): K, L and L^-1 are copied device-to-device, only alpha = Ky^-1 Y is computed.  Equal X and
 * hyper-parameters are the caller's responsibility.  Y_h: the N targets of `gp`. */
int so_fit_like(so_handle* h, int gp, int src_gp, const double* Y_h, void* stream);
/* Test/diagnostic read-back (synchronous): any of the outputs may be NULL.
 *   L_h, Linv_h: N x N row-major; alpha_h: N. */
int so_fit_export(so_handle* h, int gp, double* L_h, double* Linv_h, double* alpha_h);

/* ------------------------------------------------------------------ grid description
 * Stands in for safeopt/utilities.py:21-54 (`linearly_spaced_combinations`) when the
 * optimiser's parameter_set is bit-identical to such a grid: rows are then generated
 * from the row index on the device instead of being read from HBM.
 *   axis_values_h: concatenated per-axis linspace values (sum(n_h) doubles, exactly the
 *   doubles NumPy produced); n_h: points per axis; d axes.
 * Row order is the reference's: axis 1 slowest, then axis 0, then axes 2..d-1 fastest
 * (d == 1: the axis itself).  Must be called again after every so_fit of that gp
 * (it rebuilds the per-axis kernel factor tables); synchronous on `stream` order. */
int so_grid_define(so_handle* h, int d, const int32_t* n_h, const double* axis_values_h, void* stream);
int so_grid_prepare(so_handle* h, int gp, void* stream);
/* Same, building the large per-slow-index operand table only for grid rows [row0, row0+M) -- the
 * row block a rank evaluates with so_posterior_grid (contiguous row blocks). */
int so_grid_prepare_rows(so_handle* h, int gp, int64_t row0, int64_t M, void* stream);

/* ------------------------------------------------------------------ K2: posterior + bounds + safe bit
 * Stands in for `gp.predict_noiseless(self.inputs)` (safeopt/gp_opt.py:469) fused with
 * :471-476 (Q columns 2*gp, 2*gp+1) and this GP's factor of :481.
 * Local rows [0, M) of this rank; `row0` is the global index of local row 0 (grid path
 * only needs it to decode indices; both paths report global rows elsewhere).
 *   mean_d, var_d : (M) fp64, may be NULL
 *   Q_d           : (M x q_stride) fp64 row-major, writes columns q_col, q_col+1; may be NULL
 *   S_d           : (M) uint8, combined per safe_mode with (l > fmin); may be NULL
 * so_posterior_rows reads explicit candidates Xstar_d (M x d row-major fp64).
 * so_posterior_grid generates them from the grid defined by so_grid_define.
 * so_posterior_rows_simple is a one-thread-per-row DFMA cross-check of the same maths
 * (tests only; not a fallback -- nothing in the product calls it). */
int so_posterior_rows(so_handle* h, int gp, const double* Xstar_d, int64_t M, double beta, double fmin,
                      double* mean_d, double* var_d, double* Q_d, int q_stride, int q_col,
                      uint8_t* S_d, int safe_mode, void* stream);
int so_posterior_grid(so_handle* h, int gp, int64_t row0, int64_t M, double beta, double fmin,
                      double* mean_d, double* var_d, double* Q_d, int q_stride, int q_col,
                      uint8_t* S_d, int safe_mode, void* stream);
int so_posterior_rows_simple(so_handle* h, int gp, const double* Xstar_d, int64_t M,
                             double* mean_d, double* var_d, void* stream);
/* GPs that share their inputs X, kernel (family, lengthscales, variance) and noise share K(X,X), its
 * factor and the kernel rows k(x*, X): V = L^-1 k and the variance are common, only the mean
 * V . (L^-1 y_g) differs.  The _multi variants evaluate n <= 4 such GPs with ONE contraction
 * (the reference calls predict_noiseless once per GP, safeopt/gp_opt.py:468-476 and :969-973).
 *   gps_h    : n GP indices (the first one provides the factorisation); the library checks size,
 *              kernel and noise for equality -- equal X is the caller's responsibility
 *   fmin_h, q_col_h : n entries; mean_dh / var_dh : host arrays of n device pointers (entries or
 *              the arrays themselves may be NULL)
 *   S_d      : combined per safe_mode with AND_g (l_g > fmin_g)
 * Returns SO_ERR_CAPACITY when the shared-memory tile has no room for the extra partial sums; the
 * caller then evaluates the GPs one by one. */
int so_posterior_rows_multi(so_handle* h, int n, const int* gps_h, const double* Xstar_d, int64_t M,
                            double beta, const double* fmin_h, double* const* mean_dh,
                            double* const* var_dh, double* Q_d, int q_stride, const int* q_col_h,
                            uint8_t* S_d, int safe_mode, void* stream);
int so_posterior_grid_multi(so_handle* h, int n, const int* gps_h, int64_t row0, int64_t M,
                            double beta, const double* fmin_h, double* const* mean_dh,
                            double* const* var_dh, double* Q_d, int q_stride, const int* q_col_h,
                            uint8_t* S_d, int safe_mode, void* stream);
/* fp32 arithmetic mode of the grid path (BASELINE config 3: "fp32", posterior within 1e-4 relative): the same
 * contraction on the 5th-generation tensor cores (tcgen05.mma kind::tf32, error-compensated 3xTF32 split, fp32
 * accumulators in TMEM); fit, epilogue (bounds, safe bit) and everything downstream stay fp64.  N <= 256, RBF, grid.
 *   so_grid_prepare_f32   : packs the TF32 hi/lo operand planes for grid rows [row0, row0+M); call after so_grid_prepare(_rows)
 *                           of the same GP (again after every fit / one-point update).
 *   so_posterior_grid_f32 : so_posterior_grid_multi in that mode (n <= 4 GPs sharing X, kernel and noise; the first one
 *                           provides the operands). */
int so_grid_prepare_f32(so_handle* h, int gp, int64_t row0, int64_t M, void* stream);
int so_posterior_grid_f32(so_handle* h, int n, const int* gps_h, int64_t row0, int64_t M, double beta,
                          const double* fmin_h, double* const* mean_dh, double* const* var_dh, double* Q_d,
                          int q_stride, const int* q_col_h, uint8_t* S_d, int safe_mode, void* stream);
/* Diagnostic, host only (no device work): the assignment of the NB block rows of L^-1 to the eight warps of
 * the contraction for a fit with NB = ceil(N/8) block rows, as the posterior kernels use it when NB is not a
 * multiple of 32.  table_h: 8 passes x 8 warps x 4 slots (int16, ascending per pass, -1 = unused). */
int so_debug_row_plan(int NB, int16_t* table_h, int* npass_h);
/* The same plan with `slots` (4 or 6) block rows per warp and pass -- 6 is what the grid kernel uses for
 * NB = 36..48 (one pass over a 32-row tile).  table_h: 8 passes x 8 warps x 6 slots (int16, -1 = unused). */
int so_debug_row_plan_slots(int NB, int slots, int16_t* table_h, int* npass_h);
/* Diagnostic, host only: the tile plans the posterior kernels would use for NB block rows on a device with
 * smem_limit bytes of opt-in shared memory and num_sms SMs (n_extra = further GPs sharing the launch).
 *   out_h[0..9]   grid kernel: status, BT, RG, CG, T, npass, ring (0/1), ring stages, shared-memory bytes, warps
 *   out_h[10..16] explicit-rows kernel for M candidates of dimension d: status, BT, RG, CG, T, npass, shared-memory bytes
 *   out_h[17]     grid kernel: block rows per warp and pass (4; 2 for the 16-warp / two-CTA variants; 6 for NB = 36..48) */
int so_debug_tile_plans(int NB, int d, int64_t M, int n_extra, int64_t smem_limit, int num_sms, int64_t* out_h);
/* Materialise grid rows [row0, row0+M) as an (M x d) row-major array (tests, query point). */
int so_grid_rows(so_handle* h, int64_t row0, int64_t M, double* X_d, void* stream);

/* ------------------------------------------------------------------ K3: set logic (streaming passes over Q)
 * so_sets_reduce_safe : any(S), max l0[S], argmax (gp_opt.py:504, :512, :635, :708-712)
 * so_sets_maximizers  : M = S & (u0 >= max_l0); max width over M; best scaled width over M
 *                       (gp_opt.py:511-513, :642-644)
 * so_sets_candidates  : s = S & ~M & (max_i((u_i-l_i)/scaling_i) > max_var)
 *                           & any_i(u_i-l_i > thr_i)            (gp_opt.py:531-536)
 *                       writes the mask (may be NULL) and appends (key,row) of every
 *                       candidate, key = max_i(u_i-l_i) unscaled (gp_opt.py:551), to a
 *                       compact list of capacity `cap`; *n_cand_d counts ALL candidates.
 * scaling_h / thr_h are G host doubles (thr_i = threshold_i * beta).  Records and
 * counters live in device memory and must be zero/initialised by so_sets_* itself. */
int so_sets_reduce_safe(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0,
                        const uint8_t* S_d, so_safe_record* rec_d, void* stream);
int so_sets_maximizers(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0,
                       const uint8_t* S_d, double max_l0, const double* scaling_h,
                       uint8_t* Mmask_d, so_max_record* rec_d, void* stream);
int so_sets_candidates(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0,
                       const uint8_t* S_d, const uint8_t* Mmask_d, double max_var,
                       const double* scaling_h, const double* thr_h,
                       uint8_t* cand_mask_d, double* cand_key_d, int64_t* cand_row_d,
                       int64_t cap, int64_t* n_cand_d, void* stream);

/* Chained variants: the scalar linking two passes is taken from device memory instead of the host --
 * max_l0 = max_r safe_recs_d[r].max_l0 and max_var = max_r max_recs_d[r].max_width0 / scaling_h[0] over the
 * n_recs per-rank records of the previous pass (as all-gathered; n_recs = 1 on a single GPU).  The three
 * passes then run back to back on the stream and the host reads every record with one copy at the end. */
int so_sets_maximizers_chain(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0,
                             const uint8_t* S_d, const so_safe_record* safe_recs_d, int n_recs,
                             const double* scaling_h, uint8_t* Mmask_d, so_max_record* rec_d, void* stream);
int so_sets_candidates_chain(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0,
                             const uint8_t* S_d, const uint8_t* Mmask_d, const so_max_record* max_recs_d,
                             int n_recs, const double* scaling_h, const double* thr_h,
                             uint8_t* cand_mask_d, double* cand_key_d, int64_t* cand_row_d,
                             int64_t cap, int64_t* n_cand_d, void* stream);

/* Fused variant: the three passes above AND the cross-rank exchange of their records in ONE cooperative launch
 * (compute_sets, safeopt/gp_opt.py:483-536; query-point records of :634-644).  Between the passes every rank's
 * record is written straight into every rank's exchange buffer over NVLink (so_xchg_*), so no collective and no host
 * round trip separates them; with one rank the same kernel runs against the local buffer.
 *   with_candidates = 0 skips the candidate pass (full_sets: every safe row is a candidate, gp_opt.py:527-528).
 *   result_h: NULL, or a host buffer of SO_SETS_RESULT_BYTES(world) bytes; if given, the call copies the combined
 *   records there and synchronises `stream` before returning:
 *       [world x so_safe_record][world x so_max_record][world x int64 candidate count][int64 status][int64 epoch]
 *   (records in rank order).  With result_h == NULL the launch stays asynchronous (e.g. inside a CUDA graph) and
 *   so_sets_fused_result fetches the records later.  Returns SO_ERR_TIMEOUT when a peer never published. */
#define SO_SETS_RESULT_BYTES(world) ((size_t)(world) * 136 + 16)
int so_sets_fused(so_handle* h, const double* Q_d, int n_gps, int64_t M, int64_t row0, const uint8_t* S_d,
                  const double* scaling_h, const double* thr_h, int with_candidates, uint8_t* Mmask_d,
                  double* cand_key_d, int64_t* cand_row_d, int64_t cap, void* result_h, void* stream);
int so_sets_fused_result(so_handle* h, void* result_h, void* stream);
/* Diagnostic: SM-clock stamps (block 0) at the phase boundaries of the last so_sets_fused; needs SO_FUSED_DEBUG_TIMES=1 in the
 * environment.  out_h: 16 int64 (scan A, barrier, publish, wait, scan B, ..., end). */
int so_debug_fused_times(so_handle* h, int64_t* out_h);

/* ------------------------------------------------------------------ cross-rank record exchange (multi-GPU, one process per GPU)
 * the candidate rows are sharded, and the only data that crosses GPUs are the 64..144-byte records of
 * the reductions (safe set, maximisers, candidate count, swarm best).  They travel through peer-mapped device
 * memory written from inside the kernels, not through a collective between kernels:
 *   so_xchg_export  : allocates this rank's exchange buffer (once) and returns its CUDA IPC handle (64 bytes);
 *   so_xchg_connect : given the handles of all `world` ranks in rank order (exchanged by the host runtime, e.g. one
 *                     torch.distributed all_gather at construction), maps the peers' buffers.  Collective in the sense
 *                     that every rank must call it before the first fused kernel runs; the caller places a host barrier
 *                     after it.  Without a connect the handle behaves as world = 1. */
#define SO_XCHG_HANDLE_BYTES 64
int so_xchg_export(so_handle* h, void* ipc_handle_h);
int so_xchg_connect(so_handle* h, int world, int rank, const void* ipc_handles_h);
int so_xchg_world(const so_handle* h);

/* ------------------------------------------------------------------ K4: batched expander test
 * Stands in for the refit/predict/refit loop of safeopt/gp_opt.py:579-606 for B
 * candidates at once, through the rank-1 identity:
 *   c(x) = k(x,x_c) - k_x^T Ky^-1 k_c ,  s = var(x_c) + noise + 1e-8
 *   mean2 = mean(x) + c(x) (u_c - mean(x_c)) / s ,  var2 = max(var(x) - c(x)^2 / s, 1e-15)
 *   flag[b] = any over rows with S==0 of (mean2 - beta sqrt(var2) >= fmin)
 * for GP `gp`.  Candidate b is described by its coordinates xc_d (B x d), its posterior
 * mean/var and the fake observation value u_c (all device arrays of length B).
 * Rows come from Xstar_d (M x d) or, if Xstar_d == NULL, from the defined grid.
 * flags_d (B, uint8) is OR-ed into (caller zeroes it), so ranks/launches can accumulate. */
int so_expander_check(so_handle* h, int gp, const double* Xstar_d, int64_t row0, int64_t M,
                      const uint8_t* S_d, const double* mean_d, const double* var_d,
                      const double* xc_d, const double* mean_c_d, const double* var_c_d,
                      const double* u_c_d, int B, double beta, double fmin,
                      uint8_t* flags_d, void* stream);

/* Lipschitz variant of the expander test (the original SafeOpt rule), safeopt/gp_opt.py:558-576:
 *   flag[b] |= any over rows with S==0 of (u_c[b] - lipschitz * ||x_c[b] - x||_2 >= fmin)
 * Rows come from Xstar_d (M x d) or, if Xstar_d == NULL, from the defined grid (d is then ignored). */
int so_expander_lipschitz(so_handle* h, const double* Xstar_d, int d, int64_t row0, int64_t M,
                          const uint8_t* S_d, const double* xc_d, const double* u_c_d, int B,
                          double lipschitz, double fmin, uint8_t* flags_d, void* stream);

/* ------------------------------------------------------------------ K5/K6: swarm
 * so_swarm_fitness stands in for SafeOptSwarm._compute_particle_fitness
 * (safeopt/gp_opt.py:901-1013) given per-GP posterior planes mean_d/var_d laid out
 * (G x P) (filled by so_posterior_rows with M = P).  Writes values_d (P) and safe_d (P).
 * so_swarm_step stands in for one iteration of SwarmOptimization.run_swarm
 * (safeopt/swarm.py:98-130: velocity/position update and clipping) with host-supplied
 * uniform randoms r_d (2P x d, the reference draws them from np.random). */
int so_swarm_fitness(so_handle* h, int kind, int n_gps, int64_t P, const double* mean_d,
                     const double* var_d, double beta, const double* fmin_h,
                     const double* scaling_h, double best_lower_bound,
                     double* values_d, uint8_t* safe_d, void* stream);
int so_swarm_step(so_handle* h, int64_t P, int d, double* pos_d, double* vel_d,
                  const double* best_pos_d, const double* global_best_d, const double* r_d,
                  double inertia, const double* velocity_scale_h, const double* bounds_h,
                  void* stream);
/* personal/global best update of safeopt/swarm.py:132-146; writes argmax of best_values
 * (first index) to *best_idx_d.  When the swarm is sharded over ranks this rank holds particles
 * [p0, p0+P); if rec_d != NULL the kernel also leaves the record
 *   rec_d[0] = best value, rec_d[1] = (double)(p0 + argmax), rec_d[2..2+d) = that best position
 * (SO_SWARM_REC_DOUBLES doubles) for the cross-rank exchange. */
#define SO_SWARM_REC_DOUBLES 18
int so_swarm_update_best(so_handle* h, int64_t P, int d, const double* pos_d, const double* values_d,
                         const uint8_t* safe_d, double* best_pos_d, double* best_values_d,
                         int64_t* best_idx_d, int64_t p0, double* rec_d, void* stream);
/* `global_best = best_positions[argmax(best_values)]` (safeopt/swarm.py:146) over the records of
 * all ranks (recs_d: n_ranks x SO_SWARM_REC_DOUBLES, as all-gathered): largest value, ties to the
 * lowest global particle index.  Writes global_best_d (d) and, if not NULL, global_rec_d[0..1] =
 * {value, index}.  Stream-ordered, no host involvement: a PSO iteration has no host sync. */
int so_swarm_combine_best(so_handle* h, const double* recs_d, int n_ranks, int d,
                          double* global_best_d, double* global_rec_d, void* stream);

/* Graph-replayable PSO iteration (device-resident swarms, BASELINE config 5): nothing in these launches changes between
 * iterations, so randoms -> update -> posterior -> fitness -> bests -> global best can be captured once and replayed.
 *   state_d : 4 doubles in device memory: [0] inertia, [1] inertia increment per iteration (swarm.py:94-96, :117),
 *             [2] iteration number (selects the random block), [3] reserved.  so_swarm_update_best_x advances [0] and [2].
 *   so_swarm_rand     : out_d[(p, j)] = U[0,1) of Philox4x32-10 keyed by `seed` at (counter, global particle p0 + p, j);
 *                       the value depends on the GLOBAL particle index only, so a sharded swarm draws what one GPU would.
 *   so_swarm_step_dev : so_swarm_step with the two random blocks of swarm.py:105 drawn in the kernel (same generator,
 *                       counter = state_d[2]) and the inertia read from state_d[0].
 *   so_swarm_update_best_x : so_swarm_update_best + the cross-rank exchange + so_swarm_combine_best in ONE launch: the
 *                       last block writes this rank's {value, global index, position} into every rank's exchange buffer
 *                       (so_xchg_*; peer-mapped stores, no collective), waits for the `world` records of this iteration
 *                       and leaves the global best in global_best_d (d) / global_rec_d = {value, index, status}.
 *                       state_d may be NULL (host-driven inertia). */
int so_swarm_rand(so_handle* h, int64_t P, int d, int64_t p0, uint64_t seed, uint64_t counter, double* out_d, void* stream);
int so_swarm_step_dev(so_handle* h, int64_t P, int d, int64_t p0, double* pos_d, double* vel_d, const double* best_pos_d,
                      const double* global_best_d, const double* state_d, uint64_t seed, const double* velocity_scale_h,
                      const double* bounds_h, void* stream);
int so_swarm_update_best_x(so_handle* h, int64_t P, int d, const double* pos_d, const double* values_d,
                           const uint8_t* safe_d, double* best_pos_d, double* best_values_d, int64_t* best_idx_d,
                           int64_t p0, double* global_best_d, double* global_rec_d, double* state_d, void* stream);

/* ------------------------------------------------------------------ f1: swarm safe-set maintenance
 * Stands in for the dense correlation test of safeopt/gp_opt.py:1088-1110
 *   covariance = gp.kern.K(best_positions, vstack(S, best_positions)) / scaling[0]**2
 *   for j in range(n): accept j iff all(covariance[j, already in the set] <= 0.95)
 * without the P x (|S| + P) matrix.  corr(x,x') = k(x,x') / scale2 with the kernel of GP `gp`.
 * so_safeset_filter: keep_d[j] = all_i corr(cand_j, ref_i) <= thresh   (n candidates x m references,
 *   both row-major x d; overwrites keep_d; rows may be any shard of the candidates).
 * so_safeset_insert: the sequential walk over ALL n candidates in index order, restricted to those
 *   with keep_d[j] != 0: accept_d[j] = keep_d[j] && corr(cand_j, cand_i) <= thresh for every accepted
 *   i < j.  keep_d is used as scratch (cleared for rejected candidates).  Accepted rows are appended
 *   in index order to accepted_pos_d (capacity n x d) and counted in *n_accept_d (device). */
int so_safeset_filter(so_handle* h, int gp, const double* cand_d, int64_t n, const double* ref_d, int64_t m,
                      double scale2, double thresh, uint8_t* keep_d, void* stream);
int so_safeset_insert(so_handle* h, int gp, const double* cand_d, int64_t n, uint8_t* keep_d,
                      double scale2, double thresh, uint8_t* accept_d, double* accepted_pos_d,
                      int64_t* n_accept_d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAFEOPT_B200_H */
