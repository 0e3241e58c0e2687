/* romcomma_b200.h - C ABI of the B200-native (sm_100a) dense-GP hot path of rom-comma.
 *
 * The reference (C-O-M-M-A/rom-comma) has no FFI of its own: its hot path is Python calling tf.* ops.  Every entry point
 * below therefore replaces a group of TensorFlow/GPflow call sites; the citation names them (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to row-major float64 unless it says "host"; the caller owns all buffers;
 *  - `stream` is a cudaStream_t (may be NULL = legacy default stream); calls are asynchronous on it;
 *  - return value: 0 ok; < 0 bad argument (-2) or CUDA error (-1000 - cudaError); text via rc_last_error();
 *    a failed Cholesky (the reference's tf InvalidArgumentError) is reported through the device-side `info` word:
 *    0, or the 1-based index of the first non-positive pivot (LAPACK convention);
 *  - workspaces are sized by the *_bufsize twins and provided by the caller.  What the library keeps itself, per DEVICE and created on
 *    first use: 32 KB of tile-scheduler counters for the GEMM kernel (its only device allocation), and - for factorisations of 32 blocks
 *    or more and the optional overlapped inverse - two internal streams (high / low priority) with their events per (device, caller stream),
 *    forked from and joined to `stream` by events, so the call is still ordered on `stream` and capturable.  Per-process: the launch counter
 *    and the diagnostic GEMM profile (rc_profile_begin/end, not thread-safe).  Calls are thread-safe per (device, stream); kernel attributes
 *    (dynamic shared memory opt-in) and SM counts are tracked per device, so one process may drive several devices;
 *  - factorisation matrices are stored padded: n_pad = rc_padded(n) (multiple of 128), identity in the padding,
 *    row stride `ld` (even, >= n_pad).  Only the lower triangle is meaningful.
 *  - multi-output index convention: row (l, n) -> l*N + n  (romcomma/gpf/kernels.py:103-104, gpf/models.py:130).
 *  - `batch` lays independent problems `stride*` doubles apart (variant path: one N x N problem per output).
 */
#ifndef ROMCOMMA_B200_H
#define ROMCOMMA_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* rc_stream_t;

int rc_version(void);
const char* rc_last_error(void);
int rc_padded(int n);
/* Number of CUDA kernels this library has launched in this process (bench.py's "gpu_launches"). */
long rc_launch_count(void);
/* Roofline denominators measured on the current device with register-resident loops (host-blocking, default stream):
 * FP64 tensor-core (DMMA.8x8x4) TFLOP/s and FP64 exp() evaluations per second (in 1e9/s). scratch: >= 8 bytes of device memory. */
int rc_measure_dmma_tflops(double* scratch, double* tflops_host);
int rc_measure_exp_gexps(double* scratch, double* gexps_host);
/* The same loop over the table form of the exp (csrc/common.cuh: exp_tab, 9 FP64 instructions + one shared-memory lookup), which the
 * register form of the Sobol sweep kernel runs. */
int rc_measure_exp_tab_gexps(double* scratch, double* gexps_host);
/* Test hook: y[i] = exp(x[i]) evaluated by the device exp of the pairwise kernels; form 0 = polynomial (exp_pairwise), 1 = table (exp_tab).
 * x, y: device pointers, n elements. */
int rc_debug_exp(const double* x, double* y, long n, int form, rc_stream_t stream);

/* Per-launch profile of the dominant kernel (gemm_dmma_kernel: FP64 DMMA tiles behind potrf / trtri / lauum / trsm), for bench.py's
 * roofline: between rc_profile_begin() and rc_profile_end() every GEMM launch is bracketed by CUDA events on its own stream.
 * rc_profile_end synchronises on them and returns the summed launch durations (ms), the flops those launches executed
 * (2*128*128*K per tile actually computed) and their count.  Diagnostic, process-global, not thread-safe. */
/* Test hook, host only (no GPU needed): the order in which the tile scheduler of the GEMM hands out the 128 x 128 tiles of an M x N x K
 * product - kmode 0 full K, 1 k >= n0, 2 k < m0+128, 3 k >= m0, 4 k < n0+128; lower_only: tiles on or below the diagonal; sel_block > 0:
 * tiles whose row block lies above the column block get k steps = -1.  out_host: 4 ints per tile { m0, n0, k begin, k steps of 16 },
 * M/128 * N/128 (or the triangular count) tiles.  Returns the tile count (< 0: bad argument).  The L2-blocked raster must visit every tile
 * exactly once (tests/test_capi_symbols.py). */
int rc_debug_tile_order(int M, int N, int K, int lower_only, int kmode, int sel_block, int* out_host);

int rc_profile_begin(void);
int rc_profile_end(double* gemm_ms_host, double* gemm_flops_host, long* gemm_launches_host);

/* ---- gram ---------------------------------------------------------------------------------------------------------
 * out[(l,n),(l',n')] = F[l,l'] * exp(-1/2 sum_m (X[n,m]/ls[l,m] - X2[n',m]/ls[l',m])^2) + E[l,l'] * [n == n']
 * Replaces MOStationary.K_unit_variance / K_d_apply_variance / K_d / __call__ (romcomma/gpf/kernels.py:74-116,153-154),
 * Variance.value_to_broadcast / value_times_eye (romcomma/gpf/base.py:57-69) and MOGaussian.add_to
 * (romcomma/gpf/likelihoods.py:64-67).  F == NULL: unit variance (K_unit_variance); E == NULL: no noise term.
 * X2 == NULL means X2 = X.  rows_pad/cols_pad: multiples of 64 >= L*N, L*N2; lower_only computes 64-tiles on/below the
 * diagonal only; pad_identity puts 1 on the padded diagonal (factorisation input) instead of 0.
 * batch > 1 (variant path, gpflow kernels.RBF per output, romcomma/gpr/kernels.py:176-177, gpr/models.py:340-342):
 * problem z uses ls + z*L*M, F + z*L*L, E + z*L*L and writes out + z*stride_out (X, X2 shared).
 * Limits: 1 <= M <= 80 (the scaled coordinates of a 64 x 256 strip are staged in shared memory), L*N <= 65535*64. */
int rc_gram(const double* X, int N, const double* X2, int N2, int M, const double* ls, int L, const double* F, const double* E,
            double* out, long ld_out, long stride_out, int rows_pad, int cols_pad, int lower_only, int pad_identity, int batch,
            rc_stream_t stream);

/* K = F (x) Kunit + E (x) I from a cached unit gram: MOGPR.KXX cached branch (romcomma/gpf/models.py:66-68,139). */
int rc_apply_variance_noise(const double* Kunit, long ldu, const double* F, const double* E, int L, int N, int n_pad, double* out,
                            long ld_out, int lower_only, rc_stream_t stream);

/* ---- Cholesky and solves -------------------------------------------------------------------------------------------
 * tf.linalg.cholesky at romcomma/gpf/models.py:81, romcomma/gpr/models.py:439 and inside gpflow base_conditional.
 * `work` (rc_potrf_bufsize bytes) receives the inverted 128x128 diagonal blocks and per-block log-determinant sums that
 * rc_logdet / rc_trsv / rc_trsm_fwd / rc_potri consume. */
size_t rc_potrf_bufsize(int n_pad, int batch);
int rc_potrf(double* A, int n_pad, long ld, long strideA, int batch, void* work, int* info, rc_stream_t stream);
/* out[z] = sum_i log L_ii  (gpflow multivariate_normal, used at romcomma/gpf/models.py:82) */
int rc_logdet(const void* work, int n_pad, int batch, double* out, rc_stream_t stream);
/* x = L^-1 w (transpose = 0) or L^-T w (transpose = 1); w is overwritten.  tf.linalg.triangular_solve in
 * multivariate_normal; the two solves of tf.linalg.cholesky_solve at romcomma/gpr/models.py:444. */
int rc_trsv(const double* A, int n_pad, long ld, long strideA, int batch, const void* work, double* w, double* x, long strideV,
            int transpose, rc_stream_t stream);
/* B <- L^-1 B, B is n_pad x nrhs_pad (multiple of 128): `A = Lm^-1 Kmn` of gpflow base_conditional (romcomma/gpf/models.py:97). */
int rc_trsm_fwd(const double* A, int n_pad, long ld, long strideA, int batch, const void* work, double* B, int nrhs_pad, long ldb,
                long strideB, rc_stream_t stream);
/* The same solve for ONE large factor (batch = 1) through inverted diagonal super-blocks of 1024 rows: rc_trsm_sbinv_prepare inverts them once
 * per factor into sbwork (rc_trsm_sbinv_bufsize(n_pad, nrhs_max) bytes, caller-owned, reusable until the factor changes); rc_trsm_fwd_sbinv
 * then spends one triangular product and one rank-1024 update per super-block instead of 2 n_pad/128 latency-sized launches (n = 16384:
 * 512 right-hand sides 9.8 -> 7.7 ms).  nrhs_pad <= the nrhs_max sbwork was sized for.  Results agree with rc_trsm_fwd to rounding. */
size_t rc_trsm_sbinv_bufsize(int n_pad, int nrhs_max);
int rc_trsm_sbinv_prepare(const double* A, int n_pad, long ld, const void* work, void* sbwork, rc_stream_t stream);
int rc_trsm_fwd_sbinv(const double* A, int n_pad, long ld, const void* work, const void* sbwork, int nrhs_max, double* B, int nrhs_pad, long ldb,
                      rc_stream_t stream);
/* A (holding L) <- L^-1 in place, Kinv (lower 128-tiles) <- L^-T L^-1.  Kinv doubles as scratch and must hold n_pad*ldk doubles
 * per matrix.  Provides the explicit inverse behind the analytic gradient that replaces tf.GradientTape through
 * CholeskyGrad (romcomma/gpr/models.py:359-361). */
int rc_potri(double* A, int n_pad, long ld, long strideA, int batch, const void* work, double* Kinv, long ldk, long strideK,
             rc_stream_t stream);
/* Layout helpers between caller matrices (n x n dense) and padded storage. */
int rc_pad_identity(const double* src, int n, long stride_src, double* dst, int n_pad, long ld, long stride_dst, int batch, rc_stream_t stream);
int rc_extract_lower(const double* src, long ld, long stride_src, double* dst, int n, long stride_dst, int batch, int symmetrize,
                     rc_stream_t stream);

/* ---- fused LML + gradient ------------------------------------------------------------------------------------------
 * One evaluation of MOGPR.log_marginal_likelihood (romcomma/gpf/models.py:73-82) or of gpflow GPR.log_marginal_likelihood
 * per output (variant path), plus its gradient with respect to the ENTRIES of F, E and the lengthscales - the quantity
 * tf.GradientTape delivers inside gf.optimizers.Scipy.minimize (romcomma/gpr/models.py:359-361) before the chain rule
 * through the Variance / softplus parametrisation, which stays on the host (L x L work).
 * Problem z of `batch` uses outputs Y[:, z*L:(z+1)*L] (Y is N x (batch*L) row-major), ls + z*L*M, F + z*L*L, E + z*L*L.
 * Covariant model: batch = 1; variant model: batch = number of outputs, L = 1.
 * Kunit (optional, may be NULL; batch == 1 only): cached unpadded (L*N)^2 unit gram, ld = L*N.
 * flags: RC_GRAD_NONE = value only, RC_GRAD_VARIANCE = dF and dE, RC_GRAD_LENGTHSCALES adds dls.
 * out[z*rc_lml_grad_stride(L,M) + ...] = { lml, dF[L*L], dE[L*L], dls[L*M] }   (entries not requested are zero). */
#define RC_GRAD_NONE 0
#define RC_GRAD_VARIANCE 1
#define RC_GRAD_LENGTHSCALES 2
/* Hint: dF is only wanted on its diagonal (kernel 'covariance' not trainable, romcomma/gpr/kernels.py:54-57 default).  Without
 * RC_GRAD_LENGTHSCALES and for L > 1, batch == 1 this lets K^-1 be formed on the diagonal (l,l) blocks only ("selected" LAUUM);
 * the off-diagonal entries of dF are then returned as zero.  dE is always complete. */
#define RC_GRAD_F_DIAGONAL 4
/* With RC_OVERLAP_PANELS=<2..16> in the environment the gradient path of a single matrix (batch == 1) runs the independent part of the
 * triangular inverse on an internal low-priority stream inside the idle phases of the factorisation (events fork from / join to
 * `stream`; capturable).  Off by default (measured gain <= 1 %).  This bit keeps every kernel on `stream` regardless - what per-kernel
 * timing (rc_profile_begin/end) needs.  Same results to rounding. */
#define RC_NO_OVERLAP 8
int rc_lml_grad_stride(int L, int M);
size_t rc_lml_grad_bufsize(int N, int M, int L, int batch, int flags);
int rc_lml_grad(const double* X, const double* Y, int N, int M, int L, int batch, const double* ls, const double* F, const double* E,
                const double* Kunit, int flags, void* work, size_t work_bytes, double* out, int* info, rc_stream_t stream);

/* The same evaluation for `batch` problems that do NOT share their data - the folds of a repository (romcomma/user/run.py:60-61 fits them one
 * after another), each with its own outputs fitted independently (romcomma/gpr/models.py:340-342,360-361): problem z reads inputs X + z*Nmax*M
 * ((Nmax, M) row-major, its first Ns[z] rows), outputs Y + z*Nmax*L ((Nmax, L)), ls + z*L*M, F + z*L*L, E + z*L*L.  Ns: DEVICE array of `batch`
 * sample counts, 1 <= Ns[z] <= Nmax; the rows beyond L*Ns[z] of a problem's padded matrix are identity padding.  Workspace and output layout as
 * rc_lml_grad with N = Nmax (rc_lml_grad_bufsize(Nmax, M, L, batch, flags)); no cached unit gram. */
int rc_lml_grad_multi(const double* X, const double* Y, const int* Ns, int Nmax, int M, int L, int batch, const double* ls, const double* F,
                      const double* E, int flags, void* work, size_t work_bytes, double* out, int* info, rc_stream_t stream);

/* ---- prediction reductions -----------------------------------------------------------------------------------------
 * With A = L^-1 Kmn (n_pad x c_pad, from rc_trsm_fwd; column c = l*nstar + i) and a = L^-1 y:
 *   mean[z][i][l] = sum_k A[k][c] a[k],   var[z][i][l] = kdiag[z][l] - sum_k A[k][c]^2 (+ noise[z][l] if noise != NULL)
 * i.e. fmean and diag(Knn - A^T A) of gpflow base_conditional as used by MOGPR.predict_f (romcomma/gpf/models.py:97-111), plus
 * the diag(E) of MOGaussian._predict_mean_and_var (romcomma/gpf/likelihoods.py:85-89), in the (n*, L) layout predict_f returns,
 * without forming the full (L n*)^2 covariance.  kdiag = diag F (L per problem).  parts: rc_predict_bufsize bytes. */
size_t rc_predict_bufsize(int c_pad, int batch);
int rc_predict_reduce(const double* A, long lda, long strideA, const double* a, long stride_a, int n_pad, int c_pad, int batch, int L, int nstar,
                      const double* kdiag, const double* noise, void* parts, double* mean, double* var, rc_stream_t stream);

/* C = alpha * A^T A + beta * C for an n_pad x c_pad block A (row-major; both multiples of 128), all of the c_pad x c_pad result, on the FP64
 * tensor-core tile kernel: the A^T A of the FULL predictive covariance Knn - A^T A of gpflow base_conditional (MOGPR.predict_f with full_cov /
 * full_output_cov, romcomma/gpf/models.py:94-109) and the -W^T W of MOGP.predict_gradient (romcomma/gpr/models.py:408-411). */
int rc_syrk_tn(const double* A, int n_pad, int c_pad, long lda, long strideA, int batch, double alpha, double beta, double* C, long ldc, long strideC,
               rc_stream_t stream);

/* ---- the gradient GP dy/dx (MOGP.predict_gradient, romcomma/gpr/models.py:386-415; variant GPs: batch = L problems) ----------------------
 * rc_predict_gradient_jacobian: B[z][n][j*M+m] = d k_z(X_n, x_j)/d x_jm (:395-398) - B is (batch, n_pad, ldb) and must be zeroed by the caller where
 * it is padding - and mean[j][z][m] = sum_n B[z][n][j*M+m] KinvY[z][n] (:399).  ls (batch, M), variance (batch), KinvY (batch, N), xs (o, M).
 * Then W = K_cho^-1 B (rc_trsm_fwd), C = -W^T W (rc_syrk_tn, alpha = -1, beta = 0) and
 * rc_predict_gradient_finish: var[O][j][z][Mi][m] = C[z][O*M+Mi][j*M+m] + [Mi == m] k_z(x_O, x_j)/ls[z][Mi]^2 (:400-414; the reference's own
 * expression, which fills only the Mi == m diagonal of the prior term). */
int rc_predict_gradient_jacobian(const double* X, int N, int M, const double* xs, int o, const double* ls, const double* variance, const double* KinvY,
                                 int batch, double* B, long ldb, long strideB, double* mean, rc_stream_t stream);
int rc_predict_gradient_finish(const double* C, long ldc, long strideC, const double* xs, int o, int M, const double* ls, const double* variance,
                               int batch, double* var, rc_stream_t stream);

/* ---- Sobol ---------------------------------------------------------------------------------------------------------
 * rc_sobol_prepare: Phi, g0, g0KY (mean-centred) of ClosedSobol._calibrate / _Lambda2 (romcomma/gsa/calibrators.py:82-92,99-109,
 * 134-138).  P = L (is_F_diagonal) or L*L.  Lam (L,M); F (L) if diagonal else (L,L); KinvY (L,N).  Outputs Phi (P,M), g0 (P,N),
 * g0KY (P,N).
 * rc_sobol_contract: V[s][l][j] for a list of marginal subsets - ClosedSobol._V / marginalize (romcomma/gsa/calibrators.py:49-80)
 * and the slice loop of GSA.calibrate (romcomma/gsa/models.py:127-134).  masks_host[s] has bit m set iff input m is in subset s
 * (the reference's slice [m0:m1] is bits m0..m1-1; M <= 64).  c is g0KY (P,N).  parts: rc_sobol_bufsize bytes. */
size_t rc_sobol_bufsize(int N, int P, int nslices);
int rc_sobol_prepare(const double* X, int N, int M, const double* Lam, const double* F, const double* KinvY, int L, int is_F_diagonal,
                     double* Phi, double* g0, double* g0KY, rc_stream_t stream);
int rc_sobol_contract(const double* X, int N, int M, const double* Phi, const double* c, int L, int is_F_diagonal,
                      const unsigned long long* masks_host, int nslices, void* parts, double* V, rc_stream_t stream);

/* Multi-GPU form of rc_sobol_contract: evaluates only the 64-row tiles ti of the (N, n) pair space with ti % nparts == part and returns
 * the PARTIAL sums V; the caller adds the parts (ncclAllReduce over the ranks).  The slice loop of romcomma/gsa/models.py:127-134 has no
 * exchange step, so this is the only collective a sharded sweep needs. */
int rc_sobol_contract_part(const double* X, int N, int M, const double* Phi, const double* c, int L, int is_F_diagonal,
                           const unsigned long long* masks_host, int nslices, int part, int nparts, void* parts, double* V, rc_stream_t stream);

/* rc_sobol_error: ClosedSobolWithError.marginalize / _calibrate (romcomma/gsa/calibrators.py:146-402) for a list of marginal subsets, with the
 * diagonal F the reference requires (:380-381), in the is_T_partial form (META, :149-157): evaluates the _psi_factor (:290-309),
 * _UpsilonGaussian (:244-257), _OmegaGaussian (:214-242) and _mu_phi_mu (:259-288) chains under the DIAGONAL rank equations (:167-168) as fused
 * pairwise kernels, runs the triangular solve of :307 as one TRSM over all (subset, l, i) right-hand sides, and returns
 *   V[s][l][i]  (ClosedSobol._V, a by-product)   and   W[s][l][i] = (mu_phi_mu - mu_psi_mu) + transpose  (:324-333).
 * Phi, g0, g0KY: outputs of rc_sobol_prepare (diagonal F, P = L); Lam (L,M); F (L).
 * Achol / potrf_work: a factor from rc_potrf - chol_batch = 1 for a covariant GP (n_pad = rc_padded(L*N), K_cho of
 * romcomma/gpr/models.py:427-432) or chol_batch = L for a variant GP (one rc_padded(N) factor per output, :433-439).
 *
 * rc_sobol_error_mixed: the same plus the MIXED rank equation (:169-170) that is_T_partial = False needs - the default of all three reference
 * scripts (installation_test.py:51, csv_script.py:48, benchmark_script.py:144):
 *   WMm[s][l][i] = (mu_phi_mu_MIXED - mu_psi_mu_MIXED) + transpose  (:358-372), the covariance between the FULL model and the marginal s
 * (one more family of pairwise kernels whose Upsilon factor comes from the full model, and the dots psi^FULL_ii . psi^s_li of the solved columns).
 * From W, WMm and V the host forms Q (:400-401) and T = sqrt(|W - 2 V WMm / V[1] + V^2 Q| / V[2]^2) (:335-346): L x L work. */
size_t rc_sobol_error_bufsize(int N, int M, int L, int nslices, int n_pad, int chol_batch);
int rc_sobol_error(const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0, const double* g0KY,
                   int L, const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const void* potrf_work,
                   const unsigned long long* masks_host, int nslices, void* work, size_t work_bytes, double* V, double* W, rc_stream_t stream);
size_t rc_sobol_error_mixed_bufsize(int N, int M, int L, int nslices, int n_pad, int chol_batch);
int rc_sobol_error_mixed(const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0, const double* g0KY,
                         int L, const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const void* potrf_work,
                         const unsigned long long* masks_host, int nslices, void* work, size_t work_bytes, double* V, double* W, double* WMm,
                         rc_stream_t stream);

/* ---- fold data on the device ------------------------------------------------------------------------------------------------
 * rc_column_stats: the statistics Normalization.__init__ stores in normalization.csv (romcomma/data/storage.py:544-558) for data (N, C)
 * row-major: stats (5, C) rows = mean, std (ddof = 1, as pandas), rng = 2 sqrt(3) std, min = mean - sqrt(3) std, max = mean + sqrt(3) std.
 * rc_normalize: Normalization.apply_to (direction = +1, :469-485: the first M columns -> norm.ppf(clip((x - min) / rng, margin, 1 - margin)),
 * the others -> (y - mean) / std) and undo_from (direction = -1, :487-503), with stats in the layout above (the caller may pass the statistics
 * of OTHER data, as Fold.from_dfs does with the fold's training rows, :430-435).  In place (out == data) is allowed.
 * rc_test_metrics: the columns GPR.test adds to test.csv and its test_summary.csv row (romcomma/gpr/models.py:235-272) from truth, predictive
 * mean and sd, each (n, L): reals (n, 2L) = [ Abs Error | Z Score ], flags (n, L + 2) = outlier per output, any, all (0/1),
 * summary (3L + 2) = RMSE (L), mean SD (L), outlier fractions (L + 2). */
int rc_column_stats(const double* data, int N, int C, double* stats, rc_stream_t stream);
int rc_normalize(const double* data, long N, int M, int C, const double* stats, double margin, int direction, double* out, rc_stream_t stream);
int rc_test_metrics(const double* truth, const double* mean, const double* sd, int n, int L, double* reals, double* flags, double* summary,
                    rc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
