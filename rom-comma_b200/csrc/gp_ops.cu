// Gram construction, LML-gradient contractions and predictive reductions (FP64, sm_100a).
//
// gram_kernel replaces romcomma/gpf/kernels.py:74-116,153-154 (+ gpflow scaled_difference_matrix), the variance
// broadcast of gpf/base.py:57-60 and the noise term of gpf/likelihoods.py:64-67 / gpf/base.py:62-69 in ONE pass: the
// (L,N,L,N,M) difference tensor and the dense eye(N) noise tensor of the reference are never formed.
// grad_reduce_kernel contracts W = a a^T - K^-1 with dK/dtheta, recomputing K_unit on the fly.
#include "gp_ops.h"
#include "common.cuh"

namespace rc {

constexpr int GT = 64;          // gram / reduction tile edge
constexpr int GTHREADS = 256;   // 16 x 16 threads, 4 x 4 outputs each

// Stage the scaled coordinates s[m][r] = X[n_r][m] / ls[l_r][m] of 64 consecutive global indices into shared memory
// (transposed so that the inner loop reads consecutive doubles). Indices >= L*Npts are padding (coordinate 0, l = -1).
__device__ __forceinline__ void stage_scaled(double* __restrict__ s, int* __restrict__ lidx, int* __restrict__ nidx, const double* __restrict__ X,
                                             int Npts, int M, const double* __restrict__ ls, int L, long g0) {
  const int total = L * Npts, base = (int)g0;                       // L * N < 2^31 (row tiles are limited to 65535 x 64)
  for (int e = threadIdx.x; e < GT * M; e += GTHREADS) {
    const int r = e / M, m = e - r * M;
    const int gi = base + r;
    double v = 0.0;
    if (gi < total) {
      const int l = gi / Npts, n = gi - l * Npts;
      v = X[(long)n * M + m] / ls[l * M + m];
    }
    s[m * GT + r] = v;
  }
  for (int r = threadIdx.x; r < GT; r += GTHREADS) {
    const int gi = base + r;
    if (gi < total) {
      lidx[r] = gi / Npts;
      nidx[r] = gi % Npts;
    } else {
      lidx[r] = -1;
      nidx[r] = -1;
    }
  }
}

// out[i][j] = F[l_i,l_j] * exp(-1/2 |s_i - s_j|^2) + E[l_i,l_j] * [n_i == n_j];  identity in the padding (square case).
// One CTA per 64-row x (GSTRIP * 64)-column strip: the scaled coordinates of the rows are staged once and the prologue (global loads,
// the divisions by the lengthscales, one barrier) is paid once per GSTRIP tiles; with one tile per CTA the prologue latency was ~60 % of
// the kernel (0.72 ms for the 1.07 GB lower triangle at n = 16384; FP64 issue and HBM both far from busy).
constexpr int GSTRIP = 4;   // upper limit; GramArgs::strip is what a launch uses
__global__ void __launch_bounds__(GTHREADS) gram_kernel(GramArgs p) {
  extern __shared__ __align__(16) double sm[];
  double* sr = sm;                              // [M][64]
  double* sc = sm + GT * p.M;                   // [strip][M][64]
  int* li = reinterpret_cast<int*>(sm + (1 + p.strip) * GT * p.M);
  int* ni = li + GT;
  int* lj = ni + GT;                            // [strip][64]
  int* nj = lj + p.strip * GT;                  // [strip][64]

  const int ti = blockIdx.y, tj0 = blockIdx.x * p.strip;
  int ntj = min(p.strip, p.cols_pad / GT - tj0);
  if (p.lower_only) ntj = min(ntj, ti - tj0 + 1);
  if (ntj <= 0) return;                         // strip entirely above the diagonal
  const int z = blockIdx.z;
  const double* ls = p.ls + (long)z * p.stride_ls;
  const double* F = p.F ? p.F + (long)z * p.stride_FE : nullptr;
  const double* E = p.E ? p.E + (long)z * p.stride_FE : nullptr;
  double* out = p.out + (long)z * p.stride_out;

  const int Nr = p.Nz ? p.Nz[z] : p.N, Nc = p.Nz ? p.Nz[z] : p.N2;       // per-problem sample counts (folds) only exist for the square training gram
  __shared__ double etab[32];
  exp_table_fill(etab);
  stage_scaled(sr, li, ni, p.X + (long)z * p.stride_X, Nr, p.M, ls, p.L, (long)ti * GT);
  for (int s = 0; s < ntj; ++s)
    stage_scaled(sc + s * GT * p.M, lj + s * GT, nj + s * GT, p.X2 + (long)z * p.stride_X, Nc, p.M, ls, p.L, (long)(tj0 + s) * GT);
  __syncthreads();

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (int s = 0; s < ntj; ++s) {
    const int tj = tj0 + s;
    const double* scs = sc + s * GT * p.M;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int m = 0; m < p.M; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = sr[m * GT + ty * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = scs[m * GT + tx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const double d = a[u] - b[v];
          acc[u][v] = fma(d, d, acc[u][v]);
        }
    }
    // Fast path (all but the tiles that straddle an output boundary, the padding or the noise diagonal): one (l_i, l_j) pair for the
    // whole tile and no sample shared between its rows and columns, so the epilogue is one multiply per element.
    const int li0 = li[0], lj0 = lj[s * GT];
    const bool uniform = li0 >= 0 && lj0 >= 0 && li[GT - 1] == li0 && lj[s * GT + GT - 1] == lj0 && (!E || abs(ni[0] - nj[s * GT]) >= GT);
    if (uniform) {
      const double f = F ? F[li0 * p.L + lj0] : 1.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long gi = (long)ti * GT + ty * 4 + u;
        double2* dst = reinterpret_cast<double2*>(out + gi * p.ld_out + (long)tj * GT + tx * 4);
        dst[0] = make_double2(f * exp_tab(-0.5 * acc[u][0], etab), f * exp_tab(-0.5 * acc[u][1], etab));
        dst[1] = make_double2(f * exp_tab(-0.5 * acc[u][2], etab), f * exp_tab(-0.5 * acc[u][3], etab));
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = ty * 4 + u;
      const long gi = (long)ti * GT + r;
      const int l_i = li[r], n_i = ni[r];
      double o[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int c = tx * 4 + v;
        const long gj = (long)tj * GT + c;
        const int l_j = lj[s * GT + c], n_j = nj[s * GT + c];
        double val;
        if (l_i >= 0 && l_j >= 0) {
          val = exp_tab(-0.5 * acc[u][v], etab);
          if (F) val *= F[l_i * p.L + l_j];
          if (E && n_i == n_j) val += E[l_i * p.L + l_j];
        } else {
          val = (p.pad_identity && gi == gj) ? 1.0 : 0.0;
        }
        o[v] = val;
      }
      double2* dst = reinterpret_cast<double2*>(out + gi * p.ld_out + (long)tj * GT + tx * 4);
      dst[0] = make_double2(o[0], o[1]);
      dst[1] = make_double2(o[2], o[3]);
    }
  }
}

int gram(const GramArgs& a_in, int batch, cudaStream_t st) {
  GramArgs a = a_in;
  RC_REQUIRE(a.M >= 1 && a.M <= 80, -2, "gram: M=%d out of range [1,80]", a.M);
  RC_REQUIRE(a.rows_pad % GT == 0 && a.cols_pad % GT == 0 && a.ld_out % 2 == 0, -2, "gram: padded sizes must be multiples of 64");
  RC_REQUIRE(!a.lower_only || a.rows_pad == a.cols_pad, -2, "gram: lower_only needs a square output");
  const long tr = a.rows_pad / GT, tc = a.cols_pad / GT;
  RC_REQUIRE(tr <= 65535 && batch <= 65535, -2, "gram: %ld row tiles / %d problems exceed the grid limits", tr, batch);
  // Strips amortise the prologue over several tiles, which only pays once there are more tiles than the GPU holds at a time
  // (~5 CTAs x 148 SMs); below that one tile per CTA keeps every tile in flight at once (N = 256: 20 -> 6 us).
  const long tiles = (a.lower_only ? tr * (tr + 1) / 2 : tr * tc) * batch;
  a.strip = tiles >= 3000 ? GSTRIP : 1;
  const size_t smem = (size_t)(1 + a.strip) * GT * a.M * sizeof(double) + (size_t)(2 + 2 * a.strip) * GT * sizeof(int);
  if (smem > 48 * 1024) RC_ENSURE_SMEM(gram_kernel, smem);
  gram_kernel<<<dim3((unsigned)((tc + a.strip - 1) / a.strip), (unsigned)tr, batch), GTHREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

// K = F (x) Kunit + E (x) I  from a cached unit-variance gram (gpf/models.py:66-68 cached branch + likelihoods.add_to).
__global__ void apply_variance_noise_kernel(const double* __restrict__ Ku, long ldu, const double* __restrict__ F, const double* __restrict__ E, int L,
                                            int N, int n_pad, double* __restrict__ out, long ldo, int lower_only) {
  for (long i = blockIdx.y; i < n_pad; i += gridDim.y) {        // rows in a grid-stride loop: gridDim.y is limited to 65535
    const int l_i = i < (long)L * N ? (int)(i / N) : -1;
    const int n_i = l_i >= 0 ? (int)(i - (long)l_i * N) : -1;
    const long jend = lower_only ? ((i / TILE + 1) * TILE) : n_pad;
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < jend; j += (long)gridDim.x * blockDim.x) {
      double v;
      if (l_i >= 0 && j < (long)L * N) {
        const int l_j = (int)(j / N), n_j = (int)(j - (long)l_j * N);
        v = F[l_i * L + l_j] * Ku[i * ldu + j];
        if (E && n_i == n_j) v += E[l_i * L + l_j];
      } else {
        v = (i == j) ? 1.0 : 0.0;
      }
      out[i * ldo + j] = v;
    }
  }
}

int apply_variance_noise(const double* Ku, long ldu, const double* F, const double* E, int L, int N, int n_pad, double* out, long ldo,
                         int lower_only, cudaStream_t st) {
  dim3 grid((n_pad + 1023) / 1024, n_pad < 65535 ? n_pad : 65535);
  apply_variance_noise_kernel<<<grid, 256, 0, st>>>(Ku, ldu, F, E, L, N, n_pad, out, ldo, lower_only);
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Gradient contractions over the lower triangle of K^-1.
// For every stored element (i >= j), W = a_i a_j - Kinv_ij, U = exp(-1/2 |s_i - s_j|^2), weight 2 for the strictly lower
// elements of the (symmetric) diagonal blocks l_i == l_j so that SF, SE hold full block sums for l_i >= l_j:
//   SF[l_i,l_j] += wgt * W * U      (-> dLML/dF = 1/2 SF, mirrored)        SE[l_i,l_j] += wgt * W  if n_i == n_j
//   dls[l_i,m]  += W F U d_m s_i[m] / ls[l_i,m]  ;  dls[l_j,m] -= W F U d_m s_j[m] / ls[l_j,m]   (i > j)
// Per-CTA partial sums are written out and reduced in a fixed order by grad_finish_kernel: no atomics, bitwise reproducible.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GTHREADS) grad_reduce_kernel(GradArgs p) {
  extern __shared__ __align__(16) double sm[];
  double* sr = sm;
  double* sc = sm + GT * p.M;
  double* ar = sm + 2 * GT * p.M;   // alpha rows
  double* ac = ar + GT;             // alpha cols
  double* red = ac + GT;            // 32
  double* wpart = red + 32;         // [8 warps][2 (row/col)][slots][M]
  int* li = reinterpret_cast<int*>(wpart + 8 * 2 * p.slots * p.M);
  int* ni = li + GT;
  int* lj = ni + GT;
  int* nj = lj + GT;

  const int idx = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= idx) ++ti;
  while ((long)ti * (ti + 1) / 2 > idx) --ti;
  const int tj = idx - ti * (ti + 1) / 2;
  const int z = blockIdx.z;
  const double* ls = p.ls + (long)z * p.stride_ls;
  const double* F = p.F + (long)z * p.stride_FE;
  const double* Kinv = p.Kinv + (long)z * p.stride_K;
  const double* alpha = p.alpha + (long)z * p.stride_alpha;
  double* parts = p.parts + ((long)z * gridDim.x + idx) * p.nvals;
  const int L = p.L, M = p.M;

  for (int e = threadIdx.x; e < p.nvals; e += GTHREADS) parts[e] = 0.0;
  const int Nz = p.Nz ? p.Nz[z] : p.N;
  const double* X = p.X + (long)z * p.stride_X;
  if (p.diag_blocks_only && ((long)ti * GT) / Nz > ((long)tj * GT + GT - 1) / Nz) return;   // tile lies wholly in off-diagonal blocks
  if ((long)tj * GT >= (long)L * Nz) return;                                                 // columns are all padding (smaller fold of a batch)
  __shared__ double etab[32];
  exp_table_fill(etab);
  stage_scaled(sr, li, ni, X, Nz, M, ls, L, (long)ti * GT);
  stage_scaled(sc, lj, nj, X, Nz, M, ls, L, (long)tj * GT);
  for (int r = threadIdx.x; r < GT; r += GTHREADS) {
    ar[r] = alpha[(long)ti * GT + r];
    ac[r] = alpha[(long)tj * GT + r];
  }
  __syncthreads();

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double d2[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) d2[u][v] = 0.0;
  for (int m = 0; m < M; ++m) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = sr[m * GT + ty * 4 + u];
#pragma unroll
    for (int v = 0; v < 4; ++v) b[v] = sc[m * GT + tx * 4 + v];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const double d = a[u] - b[v];
        d2[u][v] = fma(d, d, d2[u][v]);
      }
  }
  // per-element W*U (weighted), W (noise part) and W*F*U
  double wu[4][4], wn[4][4], wfu[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = ty * 4 + u;
    const long gi = (long)ti * GT + r;
    const double2* krow = reinterpret_cast<const double2*>(Kinv + gi * p.ldk + (long)tj * GT + tx * 4);
    const double2 k01 = krow[0], k23 = krow[1];
    const double kv[4] = {k01.x, k01.y, k23.x, k23.y};
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = tx * 4 + v;
      const long gj = (long)tj * GT + c;
      const bool valid = li[r] >= 0 && lj[c] >= 0 && gj <= gi && (!p.diag_blocks_only || li[r] == lj[c]);
      const double wgt = (gj < gi && li[r] == lj[c]) ? 2.0 : 1.0;   // diagonal (l,l) blocks are symmetric: count the mirror
      const double W = valid ? (ar[r] * ac[c] - kv[v]) : 0.0;
      const double U = exp_tab(-0.5 * d2[u][v], etab);
      wu[u][v] = wgt * W * U;
      wn[u][v] = (valid && ni[r] == nj[c]) ? wgt * W : 0.0;
      wfu[u][v] = (valid && gj < gi) ? W * U * F[li[r] * L + lj[c]] : 0.0;
    }
  }
  // (l_i, l_j) pairs present in this tile
  const int la0 = max(li[0], 0), lb0 = max(lj[0], 0);
  int la1 = la0, lb1 = lb0;
  for (int r = GT - 1; r >= 0; --r) if (li[r] >= 0) { la1 = li[r]; break; }
  for (int r = GT - 1; r >= 0; --r) if (lj[r] >= 0) { lb1 = lj[r]; break; }
  for (int la = la0; la <= la1; ++la)
    for (int lb = lb0; lb <= lb1; ++lb) {
      double sF = 0.0, sE = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (li[ty * 4 + u] == la && lj[tx * 4 + v] == lb) {
            sF += wu[u][v];
            sE += wn[u][v];
          }
      sF = block_sum(sF, red);
      sE = block_sum(sE, red);
      if (threadIdx.x == 0) {
        parts[la * L + lb] = sF;
        parts[L * L + la * L + lb] = sE;
      }
    }
  if (!p.with_ls) return;
  // lengthscale terms: warp partials -> fixed-order sum over the 8 warps
  for (int m = 0; m < M; ++m) {
    double a[4], b[4], rsum[4] = {0, 0, 0, 0}, csum[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = sr[m * GT + ty * 4 + u];
#pragma unroll
    for (int v = 0; v < 4; ++v) b[v] = sc[m * GT + tx * 4 + v];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const double t = wfu[u][v] * (a[u] - b[v]);
        rsum[u] += t;
        csum[v] += t;
      }
    for (int s = 0; s < p.slots; ++s) {
      const int la = la0 + s, lb = lb0 + s;
      double vr = 0.0, vc = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (li[ty * 4 + u] == la) vr += rsum[u] * a[u];
        if (lj[tx * 4 + u] == lb) vc += csum[u] * b[u];
      }
      vr = warp_sum(vr);
      vc = warp_sum(vc);
      if (lane == 0) {
        wpart[((warp * 2 + 0) * p.slots + s) * M + m] = vr;
        wpart[((warp * 2 + 1) * p.slots + s) * M + m] = vc;
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < p.slots * M; e += GTHREADS) {
    const int s = e / M, m = e - s * M;
    const int la = la0 + s, lb = lb0 + s;
    double vr = 0.0, vc = 0.0;
    for (int w = 0; w < 8; ++w) {
      vr += wpart[((w * 2 + 0) * p.slots + s) * M + m];
      vc += wpart[((w * 2 + 1) * p.slots + s) * M + m];
    }
    // rows of this tile with output la add +vr/ls[la,m]; columns with output lb add -vc/ls[lb,m].
    // A tile may have la == lb' for different slots, so row and column parts are kept in separate partial arrays.
    if (la <= la1) parts[2 * L * L + la * M + m] = vr / ls[la * M + m];
    if (lb <= lb1) parts[2 * L * L + L * M + lb * M + m] = -vc / ls[lb * M + m];
  }
}

// out[z][v] = sum over tiles (fixed order) of parts[z][tile][v]
__global__ void grad_finish_kernel(const double* __restrict__ parts, long ntiles, int nvals, double* __restrict__ out) {
  __shared__ double red[32];
  const int v = blockIdx.x, z = blockIdx.y;
  double s = 0.0;
  for (long t = threadIdx.x; t < ntiles; t += blockDim.x) s += parts[((long)z * ntiles + t) * nvals + v];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[(long)z * nvals + v] = s;
}

int grad_nvals(int L, int M) { return 2 * L * L + 2 * L * M; }

static int grad_slots(int L, int N) { return min(L, GT / (N > 0 ? N : 1) + 2); }

size_t grad_workspace_bytes(int n_pad, int L, int M, int batch) {
  const long t = n_pad / GT;
  return (size_t)batch * (size_t)(t * (t + 1) / 2) * grad_nvals(L, M) * sizeof(double);
}

int grad_reduce(GradArgs a, int n_pad, int batch, double* out, cudaStream_t st) {
  RC_REQUIRE(n_pad % GT == 0, -2, "grad_reduce: n_pad must be a multiple of 64");
  a.nvals = grad_nvals(a.L, a.M);
  a.slots = a.Nz ? a.L : grad_slots(a.L, a.N);     // per-problem sample counts (folds): a small problem packs more outputs into one 64-row tile
  const long t = n_pad / GT, tiles = t * (t + 1) / 2;
  const size_t smem = (size_t)(2 * GT * a.M + 2 * GT + 32 + 16 * a.slots * a.M) * sizeof(double) + 4 * GT * sizeof(int);
  RC_REQUIRE(smem <= 200 * 1024, -2, "grad_reduce: M=%d / slots=%d need %zu bytes of shared memory (> 200 KB)", a.M, a.slots, smem);
  if (smem > 48 * 1024) RC_ENSURE_SMEM(grad_reduce_kernel, 200 * 1024);     // M > ~36: opt in, as gram() does (round-1 advice)
  grad_reduce_kernel<<<dim3((unsigned)tiles, 1, batch), GTHREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  grad_finish_kernel<<<dim3(a.nvals, batch), 256, 0, st>>>(a.parts, tiles, a.nvals, out);
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Predictive reductions: for A = L^-1 Kmn (n x c) and a = L^-1 y:  mean_c = sum_k A[k][c] a[k],  ss_c = sum_k A[k][c]^2.
// ----------------------------------------------------------------------------------------------------------------
constexpr int PR_SPLIT = 32;

__global__ void predict_partial_kernel(const double* __restrict__ A, long lda, long strideA, const double* __restrict__ a, long stride_a, int n,
                                       int c_pad, double* __restrict__ parts) {
  const int z = blockIdx.z, split = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_pad) return;
  const int rows = (n + PR_SPLIT - 1) / PR_SPLIT;
  const int k0 = split * rows, k1 = min(n, k0 + rows);
  const double* Az = A + (long)z * strideA;
  const double* az = a + (long)z * stride_a;
  double m = 0.0, ss = 0.0;
  for (int k = k0; k < k1; ++k) {
    const double v = Az[(long)k * lda + c];
    m = fma(v, az[k], m);
    ss = fma(v, v, ss);
  }
  double* pz = parts + ((long)z * PR_SPLIT + split) * 2 * c_pad;
  pz[c] = m;
  pz[c_pad + c] = ss;
}

// mean[z][i][l] = sum of partials,  var[z][i][l] = kdiag[z][l] - sum of squares (+ noise[z][l]);  column c = l*nstar + i
__global__ void predict_finish_kernel(const double* __restrict__ parts, int c_pad, int L, int nstar, const double* __restrict__ kdiag,
                                      const double* __restrict__ noise, double* __restrict__ mean, double* __restrict__ var) {
  const int z = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L * nstar) return;
  double m = 0.0, s = 0.0;
  for (int sp = 0; sp < PR_SPLIT; ++sp) {
    const double* pz = parts + ((long)z * PR_SPLIT + sp) * 2 * c_pad;
    m += pz[c];
    s += pz[c_pad + c];
  }
  const int l = c / nstar, i = c - l * nstar;
  const long o = ((long)z * nstar + i) * L + l;
  mean[o] = m;
  var[o] = kdiag[(long)z * L + l] - s + (noise ? noise[(long)z * L + l] : 0.0);
}

size_t predict_workspace_bytes(int c_pad, int batch) { return (size_t)batch * PR_SPLIT * 2 * c_pad * sizeof(double); }

int predict_reduce(const double* A, long lda, long strideA, const double* a, long stride_a, int n, int c_pad, int batch, int L, int nstar,
                   const double* kdiag, const double* noise, double* parts, double* mean, double* var, cudaStream_t st) {
  RC_REQUIRE(L * nstar <= c_pad, -2, "predict_reduce: L*nstar=%d exceeds c_pad=%d", L * nstar, c_pad);
  predict_partial_kernel<<<dim3((c_pad + 127) / 128, PR_SPLIT, batch), 128, 0, st>>>(A, lda, strideA, a, stride_a, n, c_pad, parts);
  RC_LAUNCH_OK();
  predict_finish_kernel<<<dim3((c_pad + 127) / 128, batch), 128, 0, st>>>(parts, c_pad, L, nstar, kdiag, noise, mean, var);
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// The gradient GP dy/dx of a variant model (romcomma/gpr/models.py:386-415).
//   J_z[n][j*M+m] = d k_z(X_n, x_j) / d x_jm = v_z exp(-1/2 sum_m' ((X_nm' - x_jm')/ls_zm')^2) (X_nm - x_jm) / ls_zm^2      (:395-398)
//   mean[j][z][m]  = sum_n J_z[n][j*M+m] KiY[z][n]                                                                           (:399)
// then W = K_cho^-1 J (rc_trsm_fwd), C = -W^T W (rc_syrk_tn) and
//   var[O][j][z][Mi][m] = C_z[O*M+Mi][j*M+m] + [Mi == m] k_z(x_O, x_j) / ls_zMi^2                                            (:400-414)
// One CTA per (test point j, output z): the threads walk the N training points, write the M Jacobian entries of each and reduce the mean.
// ----------------------------------------------------------------------------------------------------------------
constexpr int PG_MAXM = 64;

__global__ void __launch_bounds__(256) predict_gradient_jacobian_kernel(const double* __restrict__ X, int N, int M, const double* __restrict__ xs, int o,
                                                                         const double* __restrict__ ls, const double* __restrict__ variance,
                                                                         const double* __restrict__ KiY, double* __restrict__ B, long ldb, long strideB,
                                                                         double* __restrict__ mean, int batch) {
  __shared__ double xj[PG_MAXM], il2[PG_MAXM], red[32];
  const int j = blockIdx.x, z = blockIdx.y;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double l = ls[(long)z * M + m];
    xj[m] = xs[(long)j * M + m];
    il2[m] = 1.0 / (l * l);
  }
  __syncthreads();
  const double v = variance[z];
  double* Bz = B + (long)z * strideB + (long)j * M;
  double acc[PG_MAXM / 8];                      // the mean is reduced 8 inputs at a time (registers)
  for (int m0 = 0; m0 < M; m0 += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const double* xn = X + (long)n * M;
      double r2 = 0.0;
      for (int m = 0; m < M; ++m) {
        const double d = xn[m] - xj[m];
        r2 = fma(d * d, il2[m], r2);
      }
      const double k = v * exp_pairwise(-0.5 * r2), a = KiY[(long)z * N + n];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (m0 + q < M) {
          const double jac = k * (xn[m0 + q] - xj[m0 + q]) * il2[m0 + q];
          Bz[(long)n * ldb + m0 + q] = jac;
          acc[q] = fma(jac, a, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double t = block_sum(acc[q], red);
      if (threadIdx.x == 0 && m0 + q < M) mean[((long)j * batch + z) * M + m0 + q] = t;
    }
  }
}

__global__ void predict_gradient_finish_kernel(const double* __restrict__ C, long ldc, long strideC, const double* __restrict__ xs, int o, int M,
                                               const double* __restrict__ ls, const double* __restrict__ variance, int batch,
                                               double* __restrict__ var) {
  const int O = blockIdx.x, j = blockIdx.y, z = blockIdx.z;
  __shared__ double kxx;
  if (threadIdx.x == 0) {
    double r2 = 0.0;
    for (int m = 0; m < M; ++m) {
      const double d = (xs[(long)O * M + m] - xs[(long)j * M + m]) / ls[(long)z * M + m];
      r2 = fma(d, d, r2);
    }
    kxx = variance[z] * exp_pairwise(-0.5 * r2);
  }
  __syncthreads();
  const double* Cz = C + (long)z * strideC;
  double* out = var + ((((long)O * o + j) * batch + z) * M) * M;
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int Mi = e / M, m = e - Mi * M;
    double val = Cz[((long)O * M + Mi) * ldc + (long)j * M + m];
    if (Mi == m) {
      const double l = ls[(long)z * M + Mi];
      val += kxx / (l * l);
    }
    out[e] = val;
  }
}

int predict_gradient_jacobian(const double* X, int N, int M, const double* xs, int o, const double* ls, const double* variance, const double* KiY,
                              int batch, double* B, long ldb, long strideB, double* mean, cudaStream_t st) {
  RC_REQUIRE(M >= 1 && M <= PG_MAXM, -2, "predict_gradient: M=%d out of range [1,%d]", M, PG_MAXM);
  RC_REQUIRE(o >= 1 && o <= 65535 && batch >= 1 && batch <= 65535, -2, "predict_gradient: %d test points / %d outputs exceed the grid limits", o, batch);
  predict_gradient_jacobian_kernel<<<dim3(o, batch), 256, 0, st>>>(X, N, M, xs, o, ls, variance, KiY, B, ldb, strideB, mean, batch);
  RC_LAUNCH_OK();
  return 0;
}

int predict_gradient_finish(const double* C, long ldc, long strideC, const double* xs, int o, int M, const double* ls, const double* variance, int batch,
                            double* var, cudaStream_t st) {
  RC_REQUIRE(o >= 1 && o <= 65535 && batch >= 1 && batch <= 65535, -2, "predict_gradient: %d test points / %d outputs exceed the grid limits", o, batch);
  predict_gradient_finish_kernel<<<dim3(o, o, batch), 64, 0, st>>>(C, ldc, strideC, xs, o, M, ls, variance, batch, var);
  RC_LAUNCH_OK();
  return 0;
}

}  // namespace rc
