// extern "C" boundary (include/romcomma_b200.h).  Plain pointers and sizes only; no torch types.
#include <cstdlib>
#include "../../include/romcomma_b200.h"
#include "chol.h"
#include "common.cuh"
#include "gp_ops.h"
#include "sobol.h"
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <vector>

namespace rc {
static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{0};
void count_launches(long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional GEMM launch profile --------------------------------------------------------------------------------
namespace {
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> pool;   // events are created once and re-used
  size_t used = 0;
  std::vector<double> flops;
} g_prof;
cudaEvent_t prof_event() {
  if (g_prof.used == g_prof.pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_prof.pool.push_back(e);
  }
  return g_prof.pool[g_prof.used++];
}
}  // namespace
bool profile_enabled() { return g_prof.on; }
void profile_gemm_begin(cudaStream_t st) { cudaEventRecord(prof_event(), st); }
void profile_gemm_end(cudaStream_t st, double flops) {
  cudaEventRecord(prof_event(), st);
  g_prof.flops.push_back(flops);
}

static inline size_t align256(size_t b) { return (b + 255) / 256 * 256; }

struct PotrfWork {
  double* dinv;
  double* logdet_parts;
};
static PotrfWork split_potrf_work(void* work, int n_pad, int batch) {
  PotrfWork w;
  w.dinv = static_cast<double*>(work);
  w.logdet_parts = w.dinv + (size_t)batch * (n_pad / TILE) * TILE * TILE;
  return w;
}

// y_z[l*Nz + n] = Yz[n][l], zero padded;  Yz = Y + z*strideY with row stride ldY (one (N, batch*L) matrix: strideY = L, ldY = batch*L;
// per-problem (Nmax, L) matrices: strideY = Nmax*L, ldY = L).  Nz: per-problem sample counts or nullptr.
__global__ void pack_y_kernel(const double* __restrict__ Y, long strideY, int ldY, int N, const int* __restrict__ Nz, int L, int n_pad,
                              double* __restrict__ y) {
  const int z = blockIdx.y;
  const int Nq = Nz ? Nz[z] : N;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (long)gridDim.x * blockDim.x) {
    double v = 0.0;
    if (i < (long)L * Nq) {
      const int l = (int)(i / Nq), n = (int)(i - (long)l * Nq);
      v = Y[(long)z * strideY + (long)n * ldY + l];
    }
    y[(long)z * n_pad + i] = v;
  }
}

// out_z = { lml, dF, dE, dls } from the raw sums { SF, SE, dls_row, dls_col }
__global__ void lml_finalize_kernel(const double* __restrict__ logdet, const double* __restrict__ quad, const double* __restrict__ raw, int L, int M,
                                    int n_all, const int* __restrict__ Nz, int flags, double* __restrict__ out) {
  const int z = blockIdx.x;
  const int n_real = Nz ? L * Nz[z] : n_all;
  const int stride = 1 + 2 * L * L + L * M, nvals = 2 * L * L + 2 * L * M;
  double* o = out + (long)z * stride;
  const double* r = raw + (long)z * nvals;
  for (int e = threadIdx.x; e < stride; e += blockDim.x) {
    double v = 0.0;
    if (e == 0) {
      v = -0.5 * quad[z] - 0.5 * (double)n_real * 1.8378770664093454835606594728112 - logdet[z];
    } else if (flags != RC_GRAD_NONE) {
      const int k = e - 1;
      if (k < 2 * L * L) {
        const int which = k / (L * L), ij = k - which * L * L, i = ij / L, j = ij - i * L;
        const int hi = max(i, j), lo = min(i, j);
        v = 0.5 * r[which * L * L + hi * L + lo];
      } else if (flags & RC_GRAD_LENGTHSCALES) {
        const int lm = k - 2 * L * L;
        v = r[2 * L * L + lm] + r[2 * L * L + L * M + lm];
      }
    }
    o[e] = v;
  }
}

// SE[l][l'] (l > l') = sum_i (a[l*N+i] a[l'*N+i] - Kinv[(l,i),(l',i)]),  a = K^-1 y, dots[pair][i] = Kinv[(l,i),(l',i)]
__global__ void offdiag_noise_sums_kernel(const double* __restrict__ a, const double* __restrict__ dots, int N, int L, double* __restrict__ SE) {
  __shared__ double red[32];
  const int pair = blockIdx.x;
  int l = 1;
  while ((l + 1) * l / 2 <= pair) ++l;
  const int lp = pair - l * (l - 1) / 2;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += a[(long)l * N + i] * a[(long)lp * N + i] - dots[(long)pair * N + i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) SE[l * L + lp] = s;
}
}  // namespace rc

using namespace rc;

extern "C" {

int rc_version(void) { return 100; }
long rc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* rc_last_error(void) { return g_err; }
int rc_padded(int n) { return round_up(n, TILE); }

int rc_profile_begin(void) {
  g_prof.on = true;
  g_prof.used = 0;
  g_prof.flops.clear();
  return 0;
}

int rc_profile_end(double* gemm_ms_host, double* gemm_flops_host, long* gemm_launches_host) {
  g_prof.on = false;
  double ms = 0.0, fl = 0.0;
  for (size_t i = 0; i < g_prof.flops.size(); ++i) {
    RC_CUDA_OK(cudaEventSynchronize(g_prof.pool[2 * i + 1]));
    float t = 0.f;
    RC_CUDA_OK(cudaEventElapsedTime(&t, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]));
    ms += t;
    fl += g_prof.flops[i];
  }
  if (gemm_ms_host) *gemm_ms_host = ms;
  if (gemm_flops_host) *gemm_flops_host = fl;
  if (gemm_launches_host) *gemm_launches_host = (long)g_prof.flops.size();
  return 0;
}

int rc_gram(const double* X, int N, const double* X2, int N2, int M, const double* ls, int L, const double* F, const double* E, double* out,
            long ld_out, long stride_out, int rows_pad, int cols_pad, int lower_only, int pad_identity, int batch, rc_stream_t stream) {
  RC_REQUIRE(X && ls && out && N > 0 && M > 0 && L > 0 && batch > 0, -2, "rc_gram: null pointer or non-positive size");
  GramArgs a{};
  a.X = X; a.N = N;
  a.X2 = X2 ? X2 : X; a.N2 = X2 ? N2 : N;
  a.M = M; a.L = L;
  a.ls = ls; a.stride_ls = (long)L * M;
  a.F = F; a.E = E; a.stride_FE = (long)L * L;
  a.out = out; a.ld_out = ld_out; a.stride_out = stride_out;
  a.rows_pad = rows_pad; a.cols_pad = cols_pad;
  a.lower_only = lower_only; a.pad_identity = pad_identity;
  RC_REQUIRE(rows_pad >= L * a.N && cols_pad >= L * a.N2 && ld_out >= cols_pad, -2, "rc_gram: padded sizes too small");
  return gram(a, batch, (cudaStream_t)stream);
}

int rc_apply_variance_noise(const double* Kunit, long ldu, const double* F, const double* E, int L, int N, int n_pad, double* out, long ld_out,
                            int lower_only, rc_stream_t stream) {
  RC_REQUIRE(Kunit && F && out, -2, "rc_apply_variance_noise: null pointer");
  RC_REQUIRE(L > 0 && N > 0 && n_pad >= L * N && ld_out >= n_pad && ldu >= (long)L * N, -2, "rc_apply_variance_noise: inconsistent sizes");
  return apply_variance_noise(Kunit, ldu, F, E, L, N, n_pad, out, ld_out, lower_only, (cudaStream_t)stream);
}

int rc_debug_tile_order(int M, int N, int K, int lower_only, int kmode, int sel_block, int* out_host) {
  return debug_tile_order(M, N, K, lower_only, kmode, sel_block, out_host);
}

size_t rc_potrf_bufsize(int n_pad, int batch) { return align256(potrf_workspace_bytes(n_pad, batch)); }

#define RC_CHECK_FACTOR_ARGS(who)                                                                                              \
  RC_REQUIRE(n_pad > 0 && n_pad % TILE == 0, -2, who ": n_pad=%d must be a positive multiple of 128 (rc_padded)", n_pad); \
  RC_REQUIRE(batch > 0, -2, who ": batch=%d must be positive", batch)

int rc_potrf(double* A, int n_pad, long ld, long strideA, int batch, void* work, int* info, rc_stream_t stream) {
  RC_REQUIRE(A && work && info, -2, "rc_potrf: null pointer");
  RC_CHECK_FACTOR_ARGS("rc_potrf");
  PotrfWork w = split_potrf_work(work, n_pad, batch);
  return potrf_lower(A, n_pad, ld, strideA, batch, w.dinv, w.logdet_parts, info, (cudaStream_t)stream);
}

int rc_logdet(const void* work, int n_pad, int batch, double* out, rc_stream_t stream) {
  RC_REQUIRE(work && out, -2, "rc_logdet: null pointer");
  RC_CHECK_FACTOR_ARGS("rc_logdet");
  PotrfWork w = split_potrf_work(const_cast<void*>(work), n_pad, batch);
  return sum_parts(w.logdet_parts, n_pad / TILE, batch, out, 1.0, (cudaStream_t)stream);
}

int rc_trsv(const double* A, int n_pad, long ld, long strideA, int batch, const void* work, double* wv, double* x, long strideV, int transpose,
            rc_stream_t stream) {
  RC_REQUIRE(A && work && wv && x, -2, "rc_trsv: null pointer");
  RC_CHECK_FACTOR_ARGS("rc_trsv");
  PotrfWork w = split_potrf_work(const_cast<void*>(work), n_pad, batch);
  return trsv_lower(A, n_pad, ld, strideA, batch, w.dinv, wv, x, strideV, transpose, (cudaStream_t)stream);
}

int rc_trsm_fwd(const double* A, int n_pad, long ld, long strideA, int batch, const void* work, double* B, int nrhs_pad, long ldb, long strideB,
                rc_stream_t stream) {
  RC_REQUIRE(A && work && B, -2, "rc_trsm_fwd: null pointer");
  RC_CHECK_FACTOR_ARGS("rc_trsm_fwd");
  RC_REQUIRE(nrhs_pad > 0 && nrhs_pad % TILE == 0 && ldb >= nrhs_pad && ldb % 2 == 0, -2, "rc_trsm_fwd: nrhs_pad=%d must be a positive multiple of 128 and ldb=%ld even and >= it", nrhs_pad, ldb);
  PotrfWork w = split_potrf_work(const_cast<void*>(work), n_pad, batch);
  return trsm_lower_fwd(A, n_pad, ld, strideA, batch, w.dinv, B, nrhs_pad, ldb, strideB, (cudaStream_t)stream);
}

size_t rc_trsm_sbinv_bufsize(int n_pad, int nrhs_max) { return align256(trsm_sbinv_workspace_doubles(n_pad, nrhs_max) * sizeof(double)); }

int rc_trsm_sbinv_prepare(const double* A, int n_pad, long ld, const void* work, void* sbwork, rc_stream_t stream) {
  RC_REQUIRE(A && work && sbwork, -2, "rc_trsm_sbinv_prepare: null pointer");
  RC_REQUIRE(n_pad > 0 && n_pad % TILE == 0 && ld >= n_pad && ld % 2 == 0, -2, "rc_trsm_sbinv_prepare: n_pad=%d must be a positive multiple of 128, ld=%ld even and >= it", n_pad, ld);
  PotrfWork w = split_potrf_work(const_cast<void*>(work), n_pad, 1);
  return trsm_sbinv_prepare(A, n_pad, ld, w.dinv, static_cast<double*>(sbwork), (cudaStream_t)stream);
}

int rc_trsm_fwd_sbinv(const double* A, int n_pad, long ld, const void* work, const void* sbwork, int nrhs_max, double* B, int nrhs_pad, long ldb,
                      rc_stream_t stream) {
  RC_REQUIRE(A && work && sbwork && B, -2, "rc_trsm_fwd_sbinv: null pointer");
  RC_REQUIRE(n_pad > 0 && n_pad % TILE == 0 && ld >= n_pad && ld % 2 == 0, -2, "rc_trsm_fwd_sbinv: n_pad=%d must be a positive multiple of 128, ld=%ld even and >= it", n_pad, ld);
  RC_REQUIRE(nrhs_pad > 0 && nrhs_pad % TILE == 0 && nrhs_pad <= nrhs_max && ldb >= nrhs_pad && ldb % 2 == 0, -2,
             "rc_trsm_fwd_sbinv: nrhs_pad=%d must be a positive multiple of 128 and <= nrhs_max=%d, ldb=%ld even and >= it", nrhs_pad, nrhs_max, ldb);
  PotrfWork w = split_potrf_work(const_cast<void*>(work), n_pad, 1);
  const double* sb = static_cast<const double*>(sbwork);
  double* Tbuf = const_cast<double*>(sb) + trsm_sbinv_workspace_doubles(n_pad, nrhs_max) - (size_t)1024 * nrhs_max;
  return trsm_lower_fwd_sbinv(A, n_pad, ld, w.dinv, sb, Tbuf, B, nrhs_pad, ldb, (cudaStream_t)stream);
}

int rc_potri(double* A, int n_pad, long ld, long strideA, int batch, const void* work, double* Kinv, long ldk, long strideK, rc_stream_t stream) {
  RC_REQUIRE(A && work && Kinv, -2, "rc_potri: null pointer");
  RC_CHECK_FACTOR_ARGS("rc_potri");
  RC_REQUIRE(ld >= n_pad && ldk >= n_pad && ld % 2 == 0 && ldk % 2 == 0, -2, "rc_potri: leading dimensions must be even and >= n_pad");
  PotrfWork w = split_potrf_work(const_cast<void*>(work), n_pad, batch);
  int rc = trtri_lower(A, n_pad, ld, strideA, batch, w.dinv, Kinv, strideK, (cudaStream_t)stream);
  if (rc) return rc;
  return lauum_lower(A, n_pad, ld, strideA, batch, Kinv, ldk, strideK, 0, (cudaStream_t)stream);
}

int rc_pad_identity(const double* src, int n, long stride_src, double* dst, int n_pad, long ld, long stride_dst, int batch, rc_stream_t stream) {
  RC_REQUIRE(src && dst && n > 0 && n_pad >= n && ld >= n_pad && batch > 0, -2, "rc_pad_identity: null pointer or inconsistent sizes");
  return pad_identity(src, n, stride_src, dst, n_pad, ld, stride_dst, batch, (cudaStream_t)stream);
}

int rc_extract_lower(const double* src, long ld, long stride_src, double* dst, int n, long stride_dst, int batch, int symmetrize,
                     rc_stream_t stream) {
  RC_REQUIRE(src && dst && n > 0 && ld >= n && batch > 0, -2, "rc_extract_lower: null pointer or inconsistent sizes");
  return extract_lower(src, ld, stride_src, dst, n, stride_dst, batch, symmetrize, (cudaStream_t)stream);
}

int rc_lml_grad_stride(int L, int M) { return 1 + 2 * L * L + L * M; }

namespace {
// "Selected inversion": when dF is only wanted on its diagonal and no lengthscale gradient is requested, the gradient needs just
// the diagonal (l,l) blocks of K^-1 and the diagonals of its off-diagonal blocks, so LAUUM is restricted to those tiles.
// Column panels of the overlapped potrf + trtri (chol.cu).  OFF unless RC_OVERLAP_PANELS >= 2 is set in the environment: measured on the B200 it
// gains 0.3-1 % (cfg3 119.8 -> 118.7 ms at 2-4 panels, cfg4 959.7 -> 954.1 ms), because only the one-CTA diagonal kernels leave SMs
// free - a panel-solve CTA owns its SM's whole register file - which is not worth two extra streams on the default path.
int overlap_panels() {
  const char* e = getenv("RC_OVERLAP_PANELS");
  const int v = e ? atoi(e) : 0;
  return v < 0 ? 0 : v;
}
bool lml_selected(int L, int batch, int flags) {
  return (flags & RC_GRAD_F_DIAGONAL) && (flags & RC_GRAD_VARIANCE) && !(flags & RC_GRAD_LENGTHSCALES) && L > 1 && batch == 1;
}
struct LmlLayout {
  size_t A, Kinv, potrf, vec, scal, gparts, graw, dots, tg, total;
  int n_pad;
};
LmlLayout lml_layout(int N, int M, int L, int batch, int flags) {
  LmlLayout o{};
  o.n_pad = round_up(L * N, TILE);
  const size_t mat = align256((size_t)batch * o.n_pad * o.n_pad * sizeof(double));
  size_t off = 0;
  o.A = off; off += mat;
  o.Kinv = off; off += (flags != RC_GRAD_NONE) ? mat : 0;
  o.potrf = off; off += align256(potrf_workspace_bytes(o.n_pad, batch));
  o.vec = off; off += align256((size_t)4 * batch * o.n_pad * sizeof(double));
  o.scal = off; off += align256((size_t)2 * batch * sizeof(double));
  o.gparts = off; off += (flags != RC_GRAD_NONE) ? align256(grad_workspace_bytes(o.n_pad, L, M, batch)) : 0;
  o.graw = off; off += align256((size_t)batch * grad_nvals(L, M) * sizeof(double));
  o.dots = off; off += lml_selected(L, batch, flags) ? align256(block_diag_dots_workspace_bytes(N, L) + (size_t)(L * (L - 1) / 2) * N * sizeof(double)) : 0;
  o.tg = off; off += (flags != RC_GRAD_NONE) ? align256(tri_gemv_workspace_bytes(o.n_pad, batch)) : 0;
  o.total = off;
  return o;
}
}  // namespace

size_t rc_lml_grad_bufsize(int N, int M, int L, int batch, int flags) { return lml_layout(N, M, L, batch, flags & ~RC_NO_OVERLAP).total; }

// One implementation behind rc_lml_grad (one design shared by the problems of the batch, Y an (N, batch*L) matrix) and rc_lml_grad_multi (every
// problem its own (Nmax, M) inputs, (Nmax, L) outputs and sample count Nz[z] <= N: the folds of a repository).
static int lml_grad_impl(const double* X, long strideX, const double* Y, long strideY, int ldY, int N, const int* Nz, int M, int L, int batch,
                         const double* ls, const double* F, const double* E, const double* Kunit, int flags, void* work, size_t work_bytes, double* out,
                         int* info, rc_stream_t stream) {
  const bool one_stream = (flags & RC_NO_OVERLAP) != 0;       // keep every kernel on `stream`: no look-ahead, no overlapped inverse
  const bool overlap = !one_stream && batch == 1 && (flags & ~RC_NO_OVERLAP) != RC_GRAD_NONE;
  flags &= ~RC_NO_OVERLAP;
  const LmlLayout lay = lml_layout(N, M, L, batch, flags);
  RC_REQUIRE(work_bytes >= lay.total, -2, "rc_lml_grad: workspace too small (%zu < %zu)", work_bytes, lay.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = static_cast<char*>(work);
  const int n_pad = lay.n_pad, n = L * N;
  const long mat = (long)n_pad * n_pad;
  double* A = reinterpret_cast<double*>(base + lay.A);
  double* Kinv = reinterpret_cast<double*>(base + lay.Kinv);
  PotrfWork pw = split_potrf_work(base + lay.potrf, n_pad, batch);
  double* yv = reinterpret_cast<double*>(base + lay.vec);
  double* alpha = yv + (size_t)batch * n_pad;
  double* wv = alpha + (size_t)batch * n_pad;
  double* kinvy = wv + (size_t)batch * n_pad;
  double* logdet = reinterpret_cast<double*>(base + lay.scal);
  double* quad = logdet + batch;
  double* gparts = reinterpret_cast<double*>(base + lay.gparts);
  double* graw = reinterpret_cast<double*>(base + lay.graw);
  int rc;
  // 1. K = F (x) Kunit + E (x) I, lower tiles, identity padding
  if (Kunit) {
    if ((rc = apply_variance_noise(Kunit, n, F, E, L, N, n_pad, A, n_pad, 1, st))) return rc;
  } else {
    GramArgs g{};
    g.X = X; g.N = N; g.X2 = X; g.N2 = N; g.M = M; g.L = L;
    g.ls = ls; g.stride_ls = (long)L * M; g.F = F; g.E = E; g.stride_FE = (long)L * L;
    g.out = A; g.ld_out = n_pad; g.stride_out = mat; g.rows_pad = n_pad; g.cols_pad = n_pad; g.lower_only = 1; g.pad_identity = 1;
    g.stride_X = strideX; g.Nz = Nz;
    if ((rc = gram(g, batch, st))) return rc;
  }
  // 2. factor, log-determinant (gradient of one matrix: Z = L^-1 is produced as well, its independent part inside the factorisation)
  if (overlap) {
    if ((rc = potrf_trtri_lower(A, n_pad, n_pad, pw.dinv, pw.logdet_parts, info, Kinv, (size_t)mat, overlap_panels(), st, true))) return rc;
  } else if ((rc = potrf_lower(A, n_pad, n_pad, mat, batch, pw.dinv, pw.logdet_parts, info, st, !one_stream))) {
    return rc;
  }
  if ((rc = sum_parts(pw.logdet_parts, n_pad / TILE, batch, logdet, 1.0, st))) return rc;
  // 3. alpha = L^-1 y, quad = alpha^T alpha
  pack_y_kernel<<<dim3((n_pad + 255) / 256, batch), 256, 0, st>>>(Y, strideY, ldY, N, Nz, L, n_pad, yv);
  RC_LAUNCH_OK();
  RC_CUDA_OK(cudaMemsetAsync(graw, 0, (size_t)batch * grad_nvals(L, M) * sizeof(double), st));
  if (flags == RC_GRAD_NONE) {
    if ((rc = trsv_lower(A, n_pad, n_pad, mat, batch, pw.dinv, yv, alpha, n_pad, 0, st))) return rc;
    if ((rc = dot_batched(alpha, alpha, n_pad, n_pad, batch, quad, st))) return rc;
  } else {
    // 4. Z = L^-1 in place first: alpha = Z y and K^-1 y = Z^T alpha are then two passes over the triangle instead of
    //    2 x n/128 dependent substitution steps; K^-1 = Z^T Z.
    double* tg = reinterpret_cast<double*>(base + lay.tg);
    if (!overlap && (rc = trtri_lower(A, n_pad, n_pad, mat, batch, pw.dinv, Kinv, mat, st))) return rc;
    if ((rc = tri_gemv_lower(A, n_pad, n_pad, mat, batch, yv, alpha, n_pad, 0, tg, st))) return rc;
    if ((rc = dot_batched(alpha, alpha, n_pad, n_pad, batch, quad, st))) return rc;
    if ((rc = tri_gemv_lower(A, n_pad, n_pad, mat, batch, alpha, kinvy, n_pad, 1, tg, st))) return rc;
    const bool selected = lml_selected(L, batch, flags);
    if ((rc = lauum_lower(A, n_pad, n_pad, mat, batch, Kinv, n_pad, mat, selected ? N : 0, st))) return rc;
    // 5. contractions
    GradArgs ga{};
    ga.X = X; ga.N = N; ga.M = M; ga.L = L; ga.ls = ls; ga.stride_ls = (long)L * M; ga.F = F; ga.stride_FE = (long)L * L;
    ga.Kinv = Kinv; ga.ldk = n_pad; ga.stride_K = mat; ga.alpha = kinvy; ga.stride_alpha = n_pad; ga.parts = gparts;
    ga.with_ls = (flags & RC_GRAD_LENGTHSCALES) ? 1 : 0;
    ga.diag_blocks_only = selected ? 1 : 0;
    ga.stride_X = strideX; ga.Nz = Nz;
    if ((rc = grad_reduce(ga, n_pad, batch, graw, st))) return rc;
    if (selected) {   // dE off-diagonal entries from the diagonals of the off-diagonal blocks of K^-1 = Z^T Z (A holds Z = L^-1 now)
      double* dparts = reinterpret_cast<double*>(base + lay.dots);
      double* dots = dparts + block_diag_dots_workspace_bytes(N, L) / sizeof(double);
      if ((rc = block_diag_dots(A, n_pad, n_pad, N, L, dparts, dots, st))) return rc;
      offdiag_noise_sums_kernel<<<L * (L - 1) / 2, 256, 0, st>>>(kinvy, dots, N, L, graw + (size_t)L * L);
      RC_LAUNCH_OK();
    }
  }
  lml_finalize_kernel<<<batch, 128, 0, st>>>(logdet, quad, graw, L, M, n, Nz, flags, out);
  RC_LAUNCH_OK();
  return 0;
}

int rc_lml_grad(const double* X, const double* Y, int N, int M, int L, int batch, const double* ls, const double* F, const double* E,
                const double* Kunit, int flags, void* work, size_t work_bytes, double* out, int* info, rc_stream_t stream) {
  RC_REQUIRE(X && Y && ls && F && E && work && out && info, -2, "rc_lml_grad: null pointer");
  RC_REQUIRE(N > 0 && M > 0 && L > 0 && batch > 0, -2, "rc_lml_grad: non-positive size");
  RC_REQUIRE(!Kunit || batch == 1, -2, "rc_lml_grad: a cached unit gram is only supported for batch == 1");
  return lml_grad_impl(X, 0, Y, L, batch * L, N, nullptr, M, L, batch, ls, F, E, Kunit, flags, work, work_bytes, out, info, stream);
}

int rc_lml_grad_multi(const double* X, const double* Y, const int* Ns, int Nmax, int M, int L, int batch, const double* ls, const double* F,
                      const double* E, int flags, void* work, size_t work_bytes, double* out, int* info, rc_stream_t stream) {
  RC_REQUIRE(X && Y && Ns && ls && F && E && work && out && info, -2, "rc_lml_grad_multi: null pointer");
  RC_REQUIRE(Nmax > 0 && M > 0 && L > 0 && batch > 0, -2, "rc_lml_grad_multi: non-positive size");
  return lml_grad_impl(X, (long)Nmax * M, Y, (long)Nmax * L, L, Nmax, Ns, M, L, batch, ls, F, E, nullptr, flags, work, work_bytes, out, info, stream);
}

size_t rc_predict_bufsize(int c_pad, int batch) { return align256(predict_workspace_bytes(c_pad, batch)); }

int rc_predict_reduce(const double* A, long lda, long strideA, const double* a, long stride_a, int n_pad, int c_pad, int batch, int L, int nstar,
                      const double* kdiag, const double* noise, void* parts, double* mean, double* var, rc_stream_t stream) {
  RC_REQUIRE(A && a && kdiag && parts && mean && var, -2, "rc_predict_reduce: null pointer");
  return predict_reduce(A, lda, strideA, a, stride_a, n_pad, c_pad, batch, L, nstar, kdiag, noise, static_cast<double*>(parts), mean, var,
                        (cudaStream_t)stream);
}

int rc_syrk_tn(const double* A, int n_pad, int c_pad, long lda, long strideA, int batch, double alpha, double beta, double* C, long ldc, long strideC,
               rc_stream_t stream) {
  RC_REQUIRE(A && C && batch > 0, -2, "rc_syrk_tn: null pointer or empty batch");
  RC_REQUIRE(lda >= c_pad && ldc >= c_pad && lda % 2 == 0 && ldc % 2 == 0, -2, "rc_syrk_tn: leading dimensions must be even and >= c_pad");
  return syrk_tn(A, n_pad, c_pad, lda, strideA, batch, alpha, beta, C, ldc, strideC, (cudaStream_t)stream);
}

int rc_predict_gradient_jacobian(const double* X, int N, int M, const double* xs, int o, const double* ls, const double* variance, const double* KinvY,
                                 int batch, double* B, long ldb, long strideB, double* mean, rc_stream_t stream) {
  RC_REQUIRE(X && xs && ls && variance && KinvY && B && mean && N > 0, -2, "rc_predict_gradient_jacobian: null pointer or non-positive size");
  RC_REQUIRE(ldb >= (long)o * M, -2, "rc_predict_gradient_jacobian: ldb too small");
  return predict_gradient_jacobian(X, N, M, xs, o, ls, variance, KinvY, batch, B, ldb, strideB, mean, (cudaStream_t)stream);
}

int rc_predict_gradient_finish(const double* C, long ldc, long strideC, const double* xs, int o, int M, const double* ls, const double* variance,
                               int batch, double* var, rc_stream_t stream) {
  RC_REQUIRE(C && xs && ls && variance && var && M > 0, -2, "rc_predict_gradient_finish: null pointer or non-positive size");
  return predict_gradient_finish(C, ldc, strideC, xs, o, M, ls, variance, batch, var, (cudaStream_t)stream);
}

size_t rc_sobol_bufsize(int N, int P, int nslices) { return align256(sobol_workspace_bytes(N, P, nslices)); }

int rc_sobol_prepare(const double* X, int N, int M, const double* Lam, const double* F, const double* KinvY, int L, int is_F_diagonal, double* Phi,
                     double* g0, double* g0KY, rc_stream_t stream) {
  RC_REQUIRE(X && Lam && F && KinvY && Phi && g0 && g0KY, -2, "rc_sobol_prepare: null pointer");
  return sobol_prepare(X, N, M, Lam, F, KinvY, L, is_F_diagonal, Phi, g0, g0KY, (cudaStream_t)stream);
}

int rc_sobol_contract(const double* X, int N, int M, const double* Phi, const double* c, int L, int is_F_diagonal,
                      const unsigned long long* masks_host, int nslices, void* parts, double* V, rc_stream_t stream) {
  RC_REQUIRE(X && Phi && c && masks_host && parts && V && nslices > 0, -2, "rc_sobol_contract: null pointer or empty subset list");
  return sobol_contract(X, N, M, Phi, c, L, is_F_diagonal ? 1 : L, masks_host, nslices, static_cast<double*>(parts), V, 0, 1, (cudaStream_t)stream);
}

int rc_sobol_contract_part(const double* X, int N, int M, const double* Phi, const double* c, int L, int is_F_diagonal,
                           const unsigned long long* masks_host, int nslices, int part, int nparts, void* parts, double* V, rc_stream_t stream) {
  RC_REQUIRE(X && Phi && c && masks_host && parts && V && nslices > 0, -2, "rc_sobol_contract_part: null pointer or empty subset list");
  return sobol_contract(X, N, M, Phi, c, L, is_F_diagonal ? 1 : L, masks_host, nslices, static_cast<double*>(parts), V, part, nparts,
                        (cudaStream_t)stream);
}

size_t rc_sobol_error_bufsize(int N, int M, int L, int nslices, int n_pad, int chol_batch) {
  return align256(sobol_error_workspace_bytes(N, M, L, nslices, n_pad, chol_batch, 0));
}
size_t rc_sobol_error_mixed_bufsize(int N, int M, int L, int nslices, int n_pad, int chol_batch) {
  return align256(sobol_error_workspace_bytes(N, M, L, nslices, n_pad, chol_batch, 1));
}

static int sobol_error_entry(const char* who, const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0,
                             const double* g0KY, int L, const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const void* potrf_work,
                             const unsigned long long* masks_host, int nslices, void* work, size_t work_bytes, double* V, double* W, double* WMm,
                             rc_stream_t stream) {
  RC_REQUIRE(X && Lam && F && Phi && g0 && g0KY && Achol && potrf_work && masks_host && work && V && W && nslices > 0, -2,
             "%s: null pointer or empty subset list", who);
  RC_REQUIRE(N > 0 && L > 0, -2, "%s: non-positive size", who);
  RC_REQUIRE(chol_batch == 1 || chol_batch == L, -2, "%s: chol_batch must be 1 (covariant) or L (variant)", who);
  RC_REQUIRE(n_pad > 0 && n_pad % TILE == 0, -2, "%s: n_pad must be a positive multiple of 128", who);
  RC_REQUIRE(work_bytes >= sobol_error_workspace_bytes(N, M, L, nslices, n_pad, chol_batch, WMm != nullptr), -2, "%s: workspace too small", who);
  PotrfWork w = split_potrf_work(const_cast<void*>(potrf_work), n_pad, chol_batch);
  return sobol_error(X, N, M, Lam, F, Phi, g0, g0KY, L, Achol, n_pad, ld, strideA, chol_batch, w.dinv, masks_host, nslices, work, V, W, WMm,
                     (cudaStream_t)stream);
}

int rc_sobol_error(const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0, const double* g0KY,
                   int L, const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const void* potrf_work,
                   const unsigned long long* masks_host, int nslices, void* work, size_t work_bytes, double* V, double* W, rc_stream_t stream) {
  return sobol_error_entry("rc_sobol_error", X, N, M, Lam, F, Phi, g0, g0KY, L, Achol, n_pad, ld, strideA, chol_batch, potrf_work, masks_host, nslices,
                           work, work_bytes, V, W, nullptr, stream);
}

int rc_sobol_error_mixed(const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0, const double* g0KY,
                         int L, const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const void* potrf_work,
                         const unsigned long long* masks_host, int nslices, void* work, size_t work_bytes, double* V, double* W, double* WMm,
                         rc_stream_t stream) {
  RC_REQUIRE(WMm, -2, "rc_sobol_error_mixed: null pointer");
  return sobol_error_entry("rc_sobol_error_mixed", X, N, M, Lam, F, Phi, g0, g0KY, L, Achol, n_pad, ld, strideA, chol_batch, potrf_work, masks_host,
                           nslices, work, work_bytes, V, W, WMm, stream);
}

}  // extern "C"
