/* fcest_b200 -- C ABI of the B200-native (sm_100a) hot path of OnnoKampman/FCEst.
 *
 * Drop-in boundary.  The reference has no FFI of its own: its "plugin API" is the GPflow subclassing
 * contract.  Each entry point below names the reference code it replaces
 * (paths relative to the reference checkout); INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.  Conventions:
 *   - all pointers are DEVICE pointers to float64 row-major data owned by the caller; no entry point
 *     allocates, synchronises the device or touches the host (workspaces are caller supplied);
 *   - `stream` is a cudaStream_t passed as void*-compatible handle; work is enqueued on it;
 *   - return value: 0 on success, a cudaError_t value (> 0) if a launch failed, or a negative
 *     FCEST_ERR_* code for argument errors;
 *   - numerical failure of a Cholesky factorisation is reported LAPACK-style through a device
 *     `int info[]` (0 = ok, k = first non-positive pivot, 1-based), mirroring the
 *     InvalidArgumentError TensorFlow raises at fcest/models/likelihoods.py:184-189;
 *   - operands of fcest_dgemm must be 16-byte aligned with even leading dimensions.
 */
#ifndef FCEST_B200_H_
#define FCEST_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDACC__
typedef struct CUstream_st* cudaStream_t;
#endif

#define FCEST_ERR_ALIGNMENT (-1)
#define FCEST_ERR_UNSUPPORTED (-2)
#define FCEST_ERR_ARGUMENT (-3)

/* ---- likelihood: fcest/models/likelihoods.py:93-206 (variational_expectations + _log_prob) and the
 *      TF autodiff of it (fcest/helpers/inference.py:113-116).  One call = forward and, when g_fmean
 *      is non-null, the fused analytic backward.
 *   fmean, fvar : (N, R=D*nu) with row pitch ldf;  eps : (S, N, R) contiguous N(0,1) draws
 *   y : (N, D);  A : a_mode 0 -> (D, nu) matrix, 1 -> (nu,) vector; applied as a Hadamard product
 *   lam : (D,) unconstrained additive noise (softplus applied inside, likelihoods.py:253-264)
 *   ve : (N,) out, mean over samples of the Gaussian log density
 *   g_fmean, g_fvar : (N, R) pitch ldg, d sum(ve) / d fmean, fvar;  gA : (D,nu) or (nu,);  glam : (D,)
 *   ws_gA_part : (N, R) and ws_glam_part : (N, D) workspaces;  info : (N,)
 *   scratch : fcest_wishart_loglik_scratch_bytes(D, nu) bytes of global memory: 0 / NULL for D <= 55 (warp-per-
 *             matrix kernel); for 56 <= D <= 207 (shared-memory-tiled kernel) the per-CTA sample-sum slots plus G of one
 *             chunk of time steps (8 samples x one time step per SM); for 208 <= D <= 407 (cluster-tiled kernel) the
 *             per-cluster sample-sum slots plus G of one chunk (8 samples x one time step per pair of SMs).
 *   ws_gA_part / ws_glam_part hold per-CTA (or per-time-step) partial sums that the call reduces in fixed order. */
long long fcest_wishart_loglik_smem_bytes(int D, int nu);
long long fcest_wishart_loglik_scratch_bytes(int D, int nu);
int fcest_wishart_loglik(const double* fmean, const double* fvar, long long ldf, const double* eps,
                         const double* y, const double* A, int a_mode, const double* lam, int S,
                         int N, int D, int nu, double* ve, double* g_fmean, double* g_fvar,
                         long long ldg, double* gA, double* glam, double* ws_gA_part,
                         double* ws_glam_part, int* info, double* scratch, cudaStream_t stream);

/* ---- fp64 tensor-core GEMM (DMMA): C = alpha op(A) op(B) + beta C + bias_row[m] + bias_col[n].
 *      Replaces the tf.matmul calls inside gpflow.conditionals.util.base_conditional that
 *      SVGP.elbo / VGP.elbo / predict_f reach from fcest/models/wishart_process.py:109-114,381-386.
 *   a_kmajor: A[m*lda+k] else A[k*lda+m];  b_kmajor: B[n*ldb+k] else B[k*ldb+n];  batch via strides. */
long long fcest_dgemm_workspace_bytes(int M, int N, int K, long long ldc, int batch);
int fcest_dgemm(int a_kmajor, int b_kmajor, int M, int N, int K, double alpha, const double* A,
                long long lda, long long strideA, const double* B, long long ldb, long long strideB,
                double beta, double* C, long long ldc, long long strideC, const double* bias_row,
                const double* bias_col, int batch, double* workspace, long long workspace_bytes,
                cudaStream_t stream);

/* ---- the same GEMM on the int8 tensor cores (tcgen05.mma kind::i8 + TMEM + TMA): Ozaki-style emulation of
 *      C = alpha op(A) op(B) + bias_row[m]  (beta = 0, no batch).  Operand rows are cut into `slices` (2..7) balanced
 *      base-254 digits (int8); digit pairs with s + t < slices are exact integer GEMMs; error <= (slices + 1) K 254^-slices
 *      relative to max|A_m| max|B_n| (slices = 6: 47.9 bits, ~1e-12 of the largest entry at the C4 shapes; 7: 55.9 bits).  Used for the three projection GEMMs of SVGP.elbo / predict_f
 *      (fcest/models/wishart_process.py:381-386 -> gpflow base_conditional).  Deterministic (fixed-order sums).
 *      workspace: fcest_dgemm_ozaki_workspace_bytes(M, N, K, slices) bytes of device memory. */
long long fcest_dgemm_ozaki_workspace_bytes(int M, int N, int K, int slices);
int fcest_dgemm_ozaki(int a_kmajor, int b_kmajor, int M, int N, int K, double alpha, const double* A,
                      long long lda, const double* B, long long ldb, double* C, long long ldc,
                      const double* bias_row, int slices, void* workspace, long long workspace_bytes,
                      cudaStream_t stream);

/*      measurement hooks (bench.py; the only entry points that create CUDA events or synchronise): with profiling
 *      enabled every fcest_dgemm_ozaki call records an event pair around its GEMM kernel on the caller's stream;
 *      _profile_read waits for them and returns up to max_n kernel durations in milliseconds (count, or -1).
 *      fcest_int8_mma_peak launches back-to-back tcgen05.mma kind::i8 on every SM (the issue-rate ceiling of the
 *      int8 tensor pipe) and returns the number of int8 operations it issues: time it with events for the roofline. */
int fcest_dgemm_ozaki_profile(int enable);
int fcest_dgemm_ozaki_profile_read(float* ms, int max_n);
long long fcest_int8_mma_peak(int iters, cudaStream_t stream);
/*      diagnostics: while `counters` (8 x #SM long long, device) is non-null every launch records per-CTA wait cycles
 *      of the producer and MMA roles there. */
void fcest_dgemm_ozaki_debug(long long* counters);

/* ---- packed quadratic-form projection (replaces the (R,M,N) LTA tensor of base_conditional and its
 *      autodiff).  K = M(M+1)/2 packed lower-triangle entries, row pitch ldk >= K, even.
 *   pack_outer : Pk[n,(i,j)] = w_ij P[i,n] P[j,n];  bias[n] = use_prior ? kvar - sum_m P[m,n]^2 : 0
 *   tri_syrk_pack : Sk[r,(i,j)] = (tril(q_sqrt_r) tril(q_sqrt_r)^T)[i,j];  kl_parts[r] = {tr S_r, sum log L_mm^2}
 *   tri_grad : dq_sqrt_r = tril(2 W_r L_r - kl L_r) + kl diag(1/L_mm),  W_r from Gk = g_var^T Pk
 *   sym_matvec : dP[:,n] (+)= 2 T_n p_n - use_prior * 2 gs[n] p_n,  T_n from Tk = g_var Sk */
int fcest_pack_outer(const double* P, long long ldp, int M, int N, double* Pk, long long ldk,
                     double* bias, const double* kvar, int use_prior, cudaStream_t stream);
int fcest_tri_syrk_pack(const double* q_sqrt, int R, int M, double* Sk, long long ldk,
                        double* kl_parts, cudaStream_t stream);
int fcest_tri_grad(const double* Gk, long long ldk, const double* q_sqrt, int R, int M,
                   double* dq_sqrt, double kl, cudaStream_t stream);
/*   tri_grad_adam : tri_grad with the Keras-Adam update of run_adam (fcest/helpers/inference.py:72-74,111-116) fused
 *                   into its epilogue: q_sqrt (R,M,M) is updated in place on its lower triangle, m / v are its Adam
 *                   moments, alpha_t the bias-corrected step size in a DEVICE scalar (as fcest_adam_step_dev); the
 *                   dense gradient is never materialised.  M <= 200 and even, else FCEST_ERR_UNSUPPORTED. */
int fcest_tri_grad_adam(const double* Gk, long long ldk, double* q_sqrt, int R, int M, double kl, double* m,
                        double* v, const double* alpha_t, double b1, double b2, double eps, cudaStream_t stream);
int fcest_sym_matvec(const double* Tk, long long ldk, const double* P, long long ldp, int M, int N,
                     const double* g_var, long long ldg, int R, int use_prior, double* dP,
                     long long lddp, double* gs_out, int accumulate, cudaStream_t stream);

/* ---- gpflow.kernels.Matern52 (default kernel, fcest/models/wishart_process.py:99-100,326) and its
 *      adjoint; kvar / kls are device scalars holding the constrained variance and lengthscale. */
int fcest_matern52(const double* X1, int n1, const double* X2, int n2, int dim, const double* kvar,
                   const double* kls, double jitter, double* K, long long ldk, cudaStream_t stream);
int fcest_matern52_bwd(const double* X1, int n1, const double* X2, int n2, int dim,
                       const double* kvar, const double* kls, const double* gK, long long ldg,
                       int sym, double* ws_part, double* g_var, double* g_ls, double* gX1,
                       int accumulate, cudaStream_t stream);

/* ---- tf.linalg.cholesky / triangular_solve on the single M x M kernel matrix (base_conditional). */
int fcest_chol_single(double* A, int M, long long lda, int* info, cudaStream_t stream);
int fcest_tri_inv_single(const double* L, int M, long long ldl, double* Li, long long ldi,
                         cudaStream_t stream);
/*      fused, blocked variant used by the conditional: A -> chol(A) in place (lower), Li = chol(A)^-1. */
int fcest_chol_inv_single(double* A, int M, long long lda, double* Li, long long ldi, int* info,
                          cudaStream_t stream);

/* ---- glue: masks, transposes, axpby, deterministic reductions, softplus transforms. */
int fcest_tril(double* X, int M, long long ld, int mode, cudaStream_t stream);
int fcest_transpose(const double* X, int rows, int cols, long long ldx, double* Y, long long ldy,
                    cudaStream_t stream);
int fcest_axpby(long long n, double a, const double* x, double b, double* y, cudaStream_t stream);
int fcest_sum(const double* x, long long n, long long stride, double scale, double* out,
              int accumulate, cudaStream_t stream);
int fcest_sumsq(const double* x, int rows, int cols, long long ld, double scale, double* out,
                int accumulate, double* ws_296, cudaStream_t stream);
int fcest_softplus_fwd(const double* u, double* out, int n, cudaStream_t stream);
int fcest_softplus_bwd(const double* u, const double* g_c, double* g_u, int n, cudaStream_t stream);

/* ---- tf.optimizers.Adam(learning_rate=1e-3) as applied by run_adam
 *      (fcest/helpers/inference.py:72-74,111-116); t is the 1-based step, grad_sign = -1 for an ELBO. */
int fcest_adam_step(double* theta, const double* grad, double* m, double* v, long long n,
                    double grad_sign, long long t, double lr, double b1, double b2, double eps,
                    cudaStream_t stream);
/*      same update with the bias-corrected step size lr sqrt(1 - b2^t) / (1 - b1^t) read from a DEVICE scalar, so
 *      that a captured CUDA graph of the training step can be replayed while t advances. */
int fcest_adam_step_dev(double* theta, const double* grad, double* m, double* v, long long n,
                        double grad_sign, const double* alpha_t, double b1, double b2, double eps,
                        cudaStream_t stream);

/* ---- tf.random.normal stand-in (likelihoods.py:130-134, wishart_process.py:200-204,491-495):
 *      Philox4x32-10 + Box-Muller, counter based (element index, offset), keyed by seed. */
int fcest_randn(double* out, long long n, unsigned long long seed, unsigned long long offset,
                cudaStream_t stream);

/* ---- prediction: _get_cov_samples + predict_cov / predict_corr
 *      (fcest/models/wishart_process.py:120-212, :392-504; helpers/array_operations.py:56-103).
 *   cov_moments streams Sc samples Sigma = (A o F)(A o F)^T + diag(softplus(lam)) (optionally converted
 *   to correlations) into a running Welford state (mean, m2: (N,D,D), cnt0 samples already folded);
 *   `samples` (Sc,N,D,D) optionally receives the raw samples; finalize writes mean and population std.
 *   batched_chol_info: the are_all_positive_definite() assertion: -1 = not symmetric, k>0 = not PD;
 *   scratch = B*D*(D+1) doubles when D*(D+1)*8 exceeds shared memory (D > ~168), else NULL. */
long long fcest_cov_moments_scratch_bytes(int D, int nu);
int fcest_cov_moments(const double* fmean, const double* fvar, long long ldf, const double* eps,
                      const double* A, int a_mode, const double* lam, int Sc, int N, int D, int nu,
                      int corr, long long cnt0, double* mean, double* m2, double* samples,
                      int finalize, double* out_mean, double* out_std, double* scratch,
                      cudaStream_t stream);
int fcest_batched_chol_info(const double* A, int B, int D, double atol, int* info, double* scratch,
                            cudaStream_t stream);

/* ---- sliding windows: SlidingWindows.overlapping_windowed_cov_estimation
 *      (fcest/models/sliding_windows.py:109-174; np.cov per zero-padded window, optional cov -> corr of
 *      fcest/helpers/array_operations.py:37-53) and find_cross_validated_optimal_window_length (:199-274).
 *   window_cov : y (N,D) -> out (N,D,D); out[i] = cov(padded rows [i*step, i*step + w)), i < N/step, else 0.
 *                (step 1: blocks of consecutive windows per CTA, the centred Gram matrix moved from window to window
 *                by a rank-2 update and re-anchored by a direct two-pass accumulation at every block start; a scale
 *                guard falls back to the direct accumulation, so results hold to 1e-12 relative-to-max.)
 *   window_cv  : table[ti*ldt + wi] = logpdf(y_t; 0, cov(window of length w around t without its centre row)),
 *                w = w_min + wi*w_step, t = t_min + ti*t_step; -inf when the window covariance is singular
 *                (w - 2 < D, or a Cholesky pivot below rel_tol * max diagonal). */
int fcest_window_cov(const double* y, int N, int D, int window_length, int step_size, int corr,
                     double* out, cudaStream_t stream);
int fcest_window_cv(const double* y, int N, int D, int w_min, int w_step, int num_w, int t_min,
                    int t_step, int num_t, double rel_tol, double* table, long long ldt,
                    cudaStream_t stream);

/* ---- zero-phase high-pass filtering in front of the sliding-window estimator
 *      (fcest/helpers/filtering.py:12-67 -> scipy.signal.filtfilt(b, a, column), defaults padtype 'odd',
 *      padlen = 3 * (order + 1), method 'pad'; called from fcest/models/sliding_windows.py:147-154).
 *   x, out : (N, C) device arrays, every column filtered independently (C = D, or subjects * D);
 *   b, a   : HOST arrays of order + 1 transfer-function coefficients, a[0] == 1 (copied into the launch
 *            parameters); zi : HOST array of `order` steady-state initial conditions (scipy lfilter_zi);
 *   scratch: fcest_filtfilt_scratch_bytes(N, C, order) device bytes (the forward pass over the extended signal).
 *   Products and sums are rounded separately in scipy's order, so results are bit-identical to scipy's.
 *   Returns -2 when N <= padlen (scipy raises ValueError there), -3 for order > 8 or a[0] != 1. */
long long fcest_filtfilt_scratch_bytes(int N, long long C, int order);
int fcest_filtfilt(const double* x, int N, long long C, const double* b, const double* a, const double* zi,
                   int order, double* scratch, double* out, cudaStream_t stream);

/* ---- temporal summaries of a TVFC estimate (fcest/helpers/summary_measures.py:16-126):
 *   cov (N, D, D) -> out (D, D).  metric 0 = nanmean, 1 = nanvar, 2 = nanstd over the time axis (sums in time
 *   order: bit-identical to numpy), 3 = rate of change (mean_t |c_{t+1} / c_t - 1|, inf and > 10 count as 1),
 *   4 = AR(1) coefficient of every edge series (least squares of c_{t+1} on [1, c_t]; the (i, j), i > j series
 *   fills both triangles, the diagonal is 0). */
int fcest_tvfc_summary(const double* cov, int N, int D, int metric, double* out, cudaStream_t stream);

/* ---- instrumentation: number of kernels this library has launched since load (bench.py "gpu_launches"). */
long long fcest_launch_count(void);
long long fcest_launch_count_add(long long n);

#ifdef __cplusplus
}
#endif
#endif /* FCEST_B200_H_ */
