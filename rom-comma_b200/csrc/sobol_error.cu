// Standard errors of the closed Sobol indices (FP64, sm_100a): ClosedSobolWithError.marginalize / _calibrate with diagonal F
// (romcomma/gsa/calibrators.py:146-402), for is_T_partial (W[mm] only) and for the non-partial form the reference's scripts run
// (`mixed`: the MIXED rank equation W[Mm] as well, :169-170,358-372).
//
// With the size-one axes of a diagonal F removed, the reference's rank-8 Gaussian chains collapse to two families of pairwise
// kernels over the (N, n) sample pairs, both of the form
//     K[N,n] = exp( sum_{m in s} ( cA_m x_Nm^2 + cB_m x_nm^2 + cC_m x_Nm x_nm + cK_m ) )
// job (kind 0; l,i):  H_li - the kernel of ClosedSobol._V - from which  u_li[n] = sum_N c_l[N] H_li[N,n]   (_psi_factor, :290-309)
// job (kind 1; l,i):  Q_li - the Omega/Upsilon/G Gaussian ratio of _mu_phi_mu (:259-288)               w_li[n] = sum_N c_l[N] Q_li[N,n]
// Then  V[l,i] = c_i . u_li,   psi_li = L_chol^-1 (g0_i * u_li)  (block i of an LN-vector for a covariant GP),
//       W_raw[l,i] = ( pre_i * c_l . w_li  -  |psi_li|^2 ) * (1 + [l == i]),   W = W_raw + W_raw^T.
// job (kind 2; i,l), `mixed` only:  Q'_il - the same ratio with the first output equated to i (MIXED) and its Upsilon factor, which the
//       reference takes from the FULL model, folded into the left weights ct_i[N] = c_i[N] wx_i[N];   w'_il[n] = sum_N ct_i[N] Q'_il[N,n]
//       WMm_raw[l,i] = ( pre_i * c_l . w'_il  -  psi^FULL_ii . psi_li ) * (1 + [l == i]),   W[Mm] = WMm_raw + WMm_raw^T.
// The coefficient algebra is derived in oracle/sobol_error.py (checked against vectors produced by running the reference's file).
// One fused mat-vec kernel evaluates every (job, subset) pair; nothing of size N^2 is ever stored.  Partial sums are combined in
// a fixed order: bitwise reproducible, no atomics.
#include "sobol.h"
#include "chol.h"
#include "common.cuh"
#include <cstdlib>
#include <algorithm>
#include <vector>

namespace rc {

constexpr int EC = 64;          // columns (n) per CTA
constexpr int ETHREADS = 256;   // 16 x 16 threads, 4 x 4 pairs each per 64 x 64 sub-tile

// ---- per-job coefficients ------------------------------------------------------------------------------------------
// coef layout: [4][J][M] = { cA, cB, cC, cK }, J = 2*L*L, job = kind*L*L + l*L + i.  pre[i] = F_i sqrt(prod_m Lam2/(Lam2+2)).
__global__ void sobol_error_coeff_kernel(const double* __restrict__ Phi, const double* __restrict__ Lam, const double* __restrict__ F, int L, int M,
                                         int kinds, double* __restrict__ coef, double* __restrict__ pre) {
  const int J = kinds * L * L;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < J * M; e += gridDim.x * blockDim.x) {
    const int job = e / M, m = e - job * M;
    const int kind = job / (L * L), li = job - kind * L * L, l = li / L, i = li - l * L;
    const double pl = Phi[l * M + m], pi = Phi[i * M + m];
    double A, B, C, K;
    if (kind == 0) {
      const double psi = 1.0 - pl * pi, g = pl * pi / psi;
      A = g * pl; B = g * pi; C = g; K = -0.5 * log(psi);
    } else if (kind == 1) {
      const double lam2 = Lam[i * M + m] * Lam[i * M + m];
      const double ups = 1.0 / (lam2 + 2.0);
      const double gl = 1.0 - pl, gi = 1.0 - pi;
      const double r = (1.0 - ups) / (1.0 - pl * ups);
      const double a = pi * pl * pl * r;
      const double v = gl * pl + pl * pl * gi + pi * pi * pl * pl * r * gl;
      const double b = ups * pl * pl / (1.0 - ups * pl);
      A = a * a / v + b; B = pl * pl / v - pl; C = a * pl / v; K = -0.5 * log(v * (1.0 - ups * pl) / pl);
    } else {
      // MIXED: the job's first index is the x-side output (the reference's i, with l equated to it), the second the y-side output
      const double px = pl, py = pi;
      const double lam2 = Lam[l * M + m] * Lam[l * M + m];
      const double ups = 1.0 / (lam2 + 2.0);
      const double gx = 1.0 - px, gy = 1.0 - py;
      const double r = (1.0 - ups) / (1.0 - px * ups);
      const double a = px * px * py * r;
      const double v = gy * py + py * py * gx + px * px * py * py * r * gx;
      A = a * a / v; B = py * py / v - py; C = a * py / v; K = -0.5 * log(v / py);
    }
    coef[(0L * J + job) * M + m] = -0.5 * A;
    coef[(1L * J + job) * M + m] = -0.5 * B;
    coef[(2L * J + job) * M + m] = C;
    coef[(3L * J + job) * M + m] = K;
  }
  if (blockIdx.x == 0 && threadIdx.x < L) {
    const int i = threadIdx.x;
    double prod = 1.0;
    for (int m = 0; m < M; ++m) {
      const double lam2 = Lam[i * M + m] * Lam[i * M + m];
      prod *= lam2 / (lam2 + 2.0);
    }
    pre[i] = sqrt(prod) * F[i];
  }
}

// MIXED left weights: ct[i][N] = c[i][N] * prod_{ALL m} (1 - ups phi)^-1/2 exp(-1/2 b x^2), b = ups phi^2 / (1 - ups phi): the Upsilon Gaussian of
// the FULL model (calibrators.py:369 pairs the marginal Omega Gaussian with self.UpsilonGaussians.MIXED) as a per-sample factor.
__global__ void sobol_error_mixed_weight_kernel(const double* __restrict__ X, int N, int M, const double* __restrict__ Phi, const double* __restrict__ Lam,
                                                const double* __restrict__ c, double* __restrict__ ct) {
  const int i = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double e = 0.0;
  for (int m = 0; m < M; ++m) {
    const double lam2 = Lam[i * M + m] * Lam[i * M + m], ups = 1.0 / (lam2 + 2.0), ph = Phi[i * M + m];
    const double q = 1.0 - ups * ph, x = X[(long)n * M + m];
    e -= 0.5 * (ups * ph * ph / q * x * x + log(q));
  }
  ct[(long)i * N + n] = c[(long)i * N + n] * exp(e);
}

// ---- fused pairwise kernel + mat-vec -----------------------------------------------------------------------------------
struct ErrMatvecArgs {
  const double* X; int N, M;
  const double* coef;     // [4][J][M]
  const double* c;        // [L][N] left weights g0KY; job (kind,l,i) uses row l
  const double* c2;       // [L][N] left weights of the kind-2 (MIXED) jobs
  int L, J, T, RC, RCH;   // T column tiles of 64, RC row chunks of RCH rows
  int ns;
  double* parts;          // [J][RC][ns][T*64]
  unsigned long long masks[SOBOL_MAX_SLICES];
};

__global__ void __launch_bounds__(ETHREADS) sobol_error_matvec_kernel(ErrMatvecArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int M = p.M, RCH = p.RCH;
  double* cA = sm;                  // [M]
  double* cB = cA + M;
  double* cC = cB + M;
  double* cK = cC + M;
  double* xs = cK + M;              // [M][RCH]   x of this row chunk
  double* wl = xs + (long)M * RCH;  // [RCH]      left weights
  double* yc = wl + RCH;            // [M][64]    cC_m * y
  double* y2 = yc + M * EC;         // [M][64]    cB_m * y^2
  double* su = y2 + M * EC;         // [RCH]
  double* sv = su + RCH;            // [64]
  double* red = sv + EC;            // [16][64]

  const int job = blockIdx.y, J = p.J;
  const int tj = blockIdx.x / p.RC, rc = blockIdx.x - tj * p.RC;
  const int li = job % (p.L * p.L), l = li / p.L;
  const int tid = threadIdx.x;
  const int row0 = rc * RCH, col0 = tj * EC;

  __shared__ double etab[32];
  exp_table_fill(etab);
  for (int m = tid; m < M; m += ETHREADS) {
    cA[m] = p.coef[(0L * J + job) * M + m];
    cB[m] = p.coef[(1L * J + job) * M + m];
    cC[m] = p.coef[(2L * J + job) * M + m];
    cK[m] = p.coef[(3L * J + job) * M + m];
  }
  __syncthreads();
  for (long e = tid; e < (long)RCH * M; e += ETHREADS) {
    const int r = (int)(e / M), m = (int)(e - (long)r * M);
    const int gi = row0 + r;
    xs[(long)m * RCH + r] = gi < p.N ? p.X[(long)gi * M + m] : 0.0;
  }
  for (int r = tid; r < RCH; r += ETHREADS) {
    const int gi = row0 + r;
    wl[r] = gi < p.N ? (job >= 2 * p.L * p.L ? p.c2 : p.c)[(long)l * p.N + gi] : 0.0;      // rows beyond N carry zero weight
  }
  for (int e = tid; e < EC * M; e += ETHREADS) {
    const int r = e / M, m = e - r * M;
    const int gj = col0 + r;
    const double y = gj < p.N ? p.X[(long)gj * M + m] : 0.0;
    yc[m * EC + r] = cC[m] * y;
    y2[m * EC + r] = cB[m] * y * y;
  }
  __syncthreads();

  const int ty = tid >> 4, tx = tid & 15;
  const int nsub = RCH / 64;
  for (int s = 0; s < p.ns; ++s) {
    const unsigned long long mask = p.masks[s];
    for (int r = tid; r < RCH; r += ETHREADS) {
      double a = 0.0;
      for (int m = 0; m < M; ++m)
        if ((mask >> m) & 1ull) {
          const double x = xs[(long)m * RCH + r];
          a += fma(cA[m] * x, x, cK[m]);
        }
      su[r] = a;
    }
    if (tid < EC) {
      double b = 0.0;
      for (int m = 0; m < M; ++m)
        if ((mask >> m) & 1ull) b += y2[m * EC + tid];
      sv[tid] = b;
    }
    __syncthreads();
    double colacc[4] = {0.0, 0.0, 0.0, 0.0};
    double svv[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) svv[v] = sv[tx * 4 + v];
    for (int sub = 0; sub < nsub; ++sub) {
      const int rbase = sub * 64 + ty * 4;
      double e[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) e[u][v] = su[rbase + u] + svv[v];
      for (int m = 0; m < M; ++m) {
        if (!((mask >> m) & 1ull)) continue;
        double av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) av[u] = xs[(long)m * RCH + rbase + u];
#pragma unroll
        for (int v = 0; v < 4; ++v) bv[v] = yc[m * EC + tx * 4 + v];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) e[u][v] = fma(av[u], bv[v], e[u][v]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double w = wl[rbase + u];
#pragma unroll
        for (int v = 0; v < 4; ++v) colacc[v] = fma(w, exp_tab(fmin(e[u][v], 708.0), etab), colacc[v]);
      }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) red[ty * EC + tx * 4 + v] = colacc[v];
    __syncthreads();
    if (tid < EC) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += red[k * EC + tid];
      p.parts[(((long)job * p.RC + rc) * p.ns + s) * ((long)p.T * EC) + col0 + tid] = t;
    }
    __syncthreads();
  }
}


// ---- sweep form of the mat-vec ----------------------------------------------------------------------------------------------
// Like the closed indices (sobol.cu, sobol_sweep_kernel) the pairwise kernels factorise over the inputs, K_s = prod_{m in s} k_m with
// k_m[N,n] = exp( cA_m x^2 + cK_m + cB_m y^2 + cC_m x y ), so the slices GSA sweeps (singles, prefixes, suffixes, full, empty) cost M exps per
// (sample pair, job) instead of 3M+1.  CTA = (job, 16 columns n, chunk of SW_RCH rows N); thread (ty,tx) owns one column and walks the chunk
// 4 rows at a time; its 3M column sums stay in registers, the k_m of a step are parked in thread-private shared-memory columns for the suffix scan.
// Output layout per (job, row chunk):  [3M][T*32]  with the value order { F[m] | P[k] k=1..M | S[k] k=1..M-1 | E } of sobol_sweep_kernel.
constexpr int SW_EC = 16;       // columns per CTA
constexpr int SW_RCH = 512;     // rows per chunk

template <int MAXM>
__global__ void __launch_bounds__(ETHREADS, 2) sobol_error_sweep_kernel(ErrMatvecArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int M = p.M, RCH = p.RCH, nv = 3 * M;
  double* cA = sm;                  // [M]
  double* cB = cA + M;
  double* cC = cB + M;
  double* cK = cC + M;
  double* xs = cK + M;              // [M][RCH]
  double* wl = xs + (long)M * RCH;  // [RCH]
  double* yc = wl + RCH;            // [M][16]   cC_m * y
  double* y2 = yc + M * SW_EC;      // [M][16]   cB_m * y^2
  double* red = y2 + M * SW_EC;     // [16][16]
  double* hs = red + 16 * SW_EC;    // [M][4][256]

  const int job = blockIdx.y, J = p.J;
  const int tj = blockIdx.x / p.RC, rc = blockIdx.x - tj * p.RC;
  const int li = job % (p.L * p.L), l = li / p.L;
  const int tid = threadIdx.x;
  const int row0 = rc * RCH, col0 = tj * SW_EC;
  __shared__ double etab[32];
  exp_table_fill(etab);
  for (int m = tid; m < M; m += ETHREADS) {
    cA[m] = p.coef[(0L * J + job) * M + m];
    cB[m] = p.coef[(1L * J + job) * M + m];
    cC[m] = p.coef[(2L * J + job) * M + m];
    cK[m] = p.coef[(3L * J + job) * M + m];
  }
  __syncthreads();
  for (long e = tid; e < (long)RCH * M; e += ETHREADS) {
    const int r = (int)(e / M), m = (int)(e - (long)r * M);
    const int gi = row0 + r;
    xs[(long)m * RCH + r] = gi < p.N ? p.X[(long)gi * M + m] : 0.0;
  }
  for (int r = tid; r < RCH; r += ETHREADS) {
    const int gi = row0 + r;
    wl[r] = gi < p.N ? (job >= 2 * p.L * p.L ? p.c2 : p.c)[(long)l * p.N + gi] : 0.0;
  }
  for (int e = tid; e < SW_EC * M; e += ETHREADS) {
    const int r = e / M, m = e - r * M;
    const int gj = col0 + r;
    const double y = gj < p.N ? p.X[(long)gj * M + m] : 0.0;
    yc[m * SW_EC + r] = cC[m] * y;
    y2[m * SW_EC + r] = cB[m] * y * y;
  }
  __syncthreads();

  const int ty = tid >> 4, tx = tid & 15;       // 4 rows per step (ty), one column (tx)
  double accF[MAXM], accP[MAXM], accS[MAXM], accE = 0.0;
#pragma unroll
  for (int m = 0; m < MAXM; ++m) accF[m] = accP[m] = accS[m] = 0.0;
  double* hme = hs + tid;
#pragma unroll 1
  for (int sub = 0; sub < RCH / 64; ++sub) {
    const int r0 = sub * 64 + 4 * ty;
    double w[4], run[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      w[q] = wl[r0 + q];
      run[q] = w[q];
      accE += w[q];
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < M) {
        const double b0 = yc[m * SW_EC + tx], v0 = y2[m * SW_EC + tx], ca = cA[m], ck = cK[m];
        double f = 0.0, pr = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double x = xs[(long)m * RCH + r0 + q];
          const double h = exp_tab(fma(x, b0, fma(ca * x, x, ck) + v0), etab);
          if (m >= 1) hme[(m * 4 + q) * ETHREADS] = h;
          f = fma(w[q], h, f);
          run[q] *= h;
          pr += run[q];
        }
        accF[m] += f;
        accP[m] += pr;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) run[q] = w[q];
#pragma unroll
    for (int m = MAXM - 1; m >= 1; --m) {
      if (m < M) {
        double sf = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          run[q] *= hme[(m * 4 + q) * ETHREADS];
          sf += run[q];
        }
        accS[m] += sf;
      }
    }
  }
  // column sums over the 16 thread rows, one output value at a time (fixed order)
  double* outbase = p.parts + (((long)job * p.RC + rc) * nv) * ((long)p.T * SW_EC) + col0;
  auto reduce_store = [&](double v, int slot) {
    red[ty * SW_EC + tx] = v;
    __syncthreads();
    if (tid < SW_EC) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += red[k * SW_EC + tid];
      outbase[(long)slot * ((long)p.T * SW_EC) + tid] = t;
    }
    __syncthreads();
  };
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    if (m < M) {
      reduce_store(accF[m], m);
      reduce_store(accP[m], M + m);                       // P[k], k = m+1
      if (m >= 1) reduce_store(accS[m], 2 * M + m - 1);   // S[k], k = m
    }
  }
  reduce_store(accE, 3 * M - 1);
}

template <int MAXM>
static int launch_error_sweep(const ErrMatvecArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(4 * a.M + (long)a.M * a.RCH + a.RCH + 2 * a.M * SW_EC + 16 * SW_EC + (long)a.M * 4 * ETHREADS) * sizeof(double);
  RC_ENSURE_SMEM(sobol_error_sweep_kernel<MAXM>, 200 * 1024);
  RC_REQUIRE(smem <= 200 * 1024, -2, "sobol_error: shared memory %zu too large", smem);
  sobol_error_sweep_kernel<MAXM><<<dim3(a.T * a.RC, a.J), ETHREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

// Register form of the sweep mat-vec (round 2; same changes as sobol.cu: sobol_sweep_reg_kernel): the k_m of a step stay in registers, one
// instantiation per M, row operands { x, cA x^2 + cK } and column operands { cC y, cB y^2 } interleaved for 16-byte loads, 128 threads and
// three CTAs per SM.  A CTA owns 16 columns and walks ALL rows in stages of SWR_ROWS (accumulators persist), so there is ONE partial per
// (job, column) instead of one per 512-row chunk: the gather kernel reads 1/8 of the bytes.  Thread (ty, tx): column tx, RQ rows per step.
constexpr int SWR_THREADS = 128;
constexpr int SWR_ROWS = 128;    // rows per stage

template <int M, int RQ>
__global__ void __launch_bounds__(SWR_THREADS, (M <= 12 ? 3 : 2)) sobol_error_sweep_reg_kernel(ErrMatvecArgs p) {
  extern __shared__ __align__(16) double sm[];
  constexpr int nv = 3 * M, MP = (M + 1) & ~1;
  double* cA = sm;                                          // [M]
  double* cB = cA + MP;
  double* cC = cB + MP;
  double* cK = cC + MP;
  double2* rowd = reinterpret_cast<double2*>(cK + MP);      // [M][SWR_ROWS]  { x, cA_m x^2 + cK_m }
  double2* cold = rowd + M * SWR_ROWS;                      // [M][16]        { cC_m y, cB_m y^2 }
  double* wl = reinterpret_cast<double*>(cold + M * SW_EC); // [SWR_ROWS]     left weights
  double* red = wl + SWR_ROWS;                              // [8][16]
  __shared__ double etab[32];

  const int job = blockIdx.y, J = p.J;
  const int li = job % (p.L * p.L), l = li / p.L;
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * SW_EC;
  const double* wsrc = (job >= 2 * p.L * p.L ? p.c2 : p.c) + (long)l * p.N;
  exp_table_fill(etab);
  for (int m = tid; m < M; m += SWR_THREADS) {
    cA[m] = p.coef[(0L * J + job) * M + m];
    cB[m] = p.coef[(1L * J + job) * M + m];
    cC[m] = p.coef[(2L * J + job) * M + m];
    cK[m] = p.coef[(3L * J + job) * M + m];
  }
  __syncthreads();
  for (int e = tid; e < SW_EC * M; e += SWR_THREADS) {
    const int r = e / M, m = e - r * M;
    const int gj = col0 + r;
    const double y = gj < p.N ? p.X[(long)gj * M + m] : 0.0;
    cold[m * SW_EC + r] = make_double2(cC[m] * y, cB[m] * y * y);
  }

  const int ty = tid >> 4, tx = tid & 15;       // RQ rows per step (ty), one column (tx)
  // accF[m] = F[m] for m >= 1 (F[0] is P[1]);  accP[k-1] = P[k];  accS[k] = S[k] for 1 <= k <= M-2 (S[M-1] is F[M-1]);  accS[0] = E
  double accF[M], accP[M], accS[M];
#pragma unroll
  for (int m = 0; m < M; ++m) accF[m] = accP[m] = accS[m] = 0.0;
#pragma unroll 1
  for (int row0 = 0; row0 < p.N; row0 += SWR_ROWS) {
    if (row0 > 0) __syncthreads();               // everybody has finished with the previous stage
    for (int e = tid; e < SWR_ROWS * M; e += SWR_THREADS) {
      const int r = e / M, m = e - r * M;
      const int gi = row0 + r;
      const double x = gi < p.N ? p.X[(long)gi * M + m] : 0.0;
      rowd[m * SWR_ROWS + r] = make_double2(x, fma(cA[m] * x, x, cK[m]));
    }
    for (int r = tid; r < SWR_ROWS; r += SWR_THREADS) {
      const int gi = row0 + r;
      wl[r] = gi < p.N ? wsrc[gi] : 0.0;         // rows beyond N carry zero weight
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < SWR_ROWS / (8 * RQ); ++sub) {
      const int r0 = sub * 8 * RQ + RQ * ty;
      double w[RQ], run[RQ], h[M][RQ];
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        w[q] = wl[r0 + q];
        run[q] = w[q];
        accS[0] += w[q];
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double2 cd = cold[m * SW_EC + tx];
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
          const double2 rw = rowd[m * SWR_ROWS + r0 + q];
          h[m][q] = exp_tab(fma(rw.x, cd.x, rw.y + cd.y), etab);
          if (m >= 1) accF[m] = fma(w[q], h[m][q], accF[m]);
          run[q] *= h[m][q];
          accP[m] += run[q];
        }
      }
      if constexpr (M >= 3) {
#pragma unroll
        for (int q = 0; q < RQ; ++q) run[q] = w[q] * h[M - 1][q];
#pragma unroll
        for (int m = M - 2; m >= 1; --m) {
#pragma unroll
          for (int q = 0; q < RQ; ++q) {
            run[q] *= h[m][q];
            accS[m] += run[q];
          }
        }
      }
    }
  }
  // column sums over the 8 thread rows, one output value at a time (fixed order).  Output slots: F[m] -> m; P[k] -> M+k-1; S[k] -> 2M+k-1; E -> 3M-1
  double* outbase = p.parts + ((long)job * nv) * ((long)p.T * SW_EC) + col0;
  auto reduce_store = [&](double v, int slot) {
    red[ty * SW_EC + tx] = v;
    __syncthreads();
    if (tid < SW_EC) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < SWR_THREADS / 16; ++k) t += red[k * SW_EC + tid];
      outbase[(long)slot * ((long)p.T * SW_EC) + tid] = t;
    }
    __syncthreads();
  };
#pragma unroll
  for (int m = 0; m < M; ++m) {
    reduce_store(accP[m], M + m);                                 // P[k], k = m + 1
    reduce_store(m >= 1 ? accF[m] : accP[0], m);                  // F[0] = P[1]
    if (m >= 1 && m <= M - 2) reduce_store(accS[m], 2 * M + m - 1);
    if (m == M - 1 && M >= 2) reduce_store(accF[m], 2 * M + m - 1);   // S[M-1] = F[M-1]
  }
  reduce_store(accS[0], 3 * M - 1);
}

template <int M, int RQ>
static int launch_error_sweep_reg(const ErrMatvecArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(4 * ((M + 1) & ~1) + 2 * M * SWR_ROWS + 2 * M * SW_EC + SWR_ROWS + (SWR_THREADS / 16) * SW_EC) * sizeof(double);
  static_assert((4 * 20 + 2 * 20 * SWR_ROWS + 2 * 20 * SW_EC + SWR_ROWS + 8 * SW_EC) * sizeof(double) <= 48 * 1024, "default dynamic shared memory limit");
  sobol_error_sweep_reg_kernel<M, RQ><<<dim3(a.T, a.J), SWR_THREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

// slice -> (column of the mat-vec output, destination row of V / W); passed by value
struct ErrMap {
  int col[32];
  int dest[32];
};

// ---- gather: V, the Omega bilinear forms R (and Rm, MIXED), and the right-hand sides g0_i * u_li of the triangular solve --------------
// grid (ns, L*L).  Covariant GP (chol_batch == 1): B is n_pad x ncol, column (s*L + l)*L + i, rows i*N + n.
// Variant GP (chol_batch == L): problem i has its own N_pad x ncol block at B + i*strideB, column s*L + l, rows n.
__global__ void sobol_error_gather_kernel(const double* __restrict__ parts, int L, int N, int RC, int nv, long ncols_u, const double* __restrict__ c,
                                          const double* __restrict__ g0, const double* __restrict__ pre, int chol_batch, double* __restrict__ B,
                                          long ldb, long strideB, ErrMap map, double* __restrict__ V, double* __restrict__ R, double* __restrict__ Rm) {
  __shared__ double red[32];
  const int s = blockIdx.x, li = blockIdx.y, l = li / L, i = li - l * L;
  const double* pu = parts + (((long)(0 * L * L + li) * RC) * nv + map.col[s]) * ncols_u;     // nv values per (job, row chunk)
  const double* pw = parts + (((long)(1 * L * L + li) * RC) * nv + map.col[s]) * ncols_u;
  const double* pm = Rm ? parts + (((long)(2 * L * L + i * L + l) * RC) * nv + map.col[s]) * ncols_u : nullptr;   // MIXED job (x side i, y side l)
  const long rc_stride = (long)nv * ncols_u;
  double vacc = 0.0, racc = 0.0, macc = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double u = 0.0, w = 0.0, wm = 0.0;
    for (int r = 0; r < RC; ++r) {
      u += pu[r * rc_stride + n];
      w += pw[r * rc_stride + n];
      if (pm) wm += pm[r * rc_stride + n];
    }
    vacc = fma(c[(long)i * N + n], u, vacc);
    racc = fma(c[(long)l * N + n], w, racc);
    macc = fma(c[(long)l * N + n], wm, macc);
    const double f = g0[(long)i * N + n] * u;
    if (chol_batch == 1) B[((long)i * N + n) * ldb + ((long)s * L + l) * L + i] = f;
    else B[(long)i * strideB + (long)n * ldb + (long)s * L + l] = f;
  }
  vacc = block_sum(vacc, red);
  racc = block_sum(racc, red);
  macc = block_sum(macc, red);
  if (threadIdx.x == 0) {
    if (V) V[((long)map.dest[s] * L + l) * L + i] = vacc;
    R[((long)s * L + l) * L + i] = pre[i] * racc;
    if (Rm) Rm[((long)s * L + l) * L + i] = pre[i] * macc;
  }
}

// partial[chunk][col] = sum over 128 rows of B[row][col]^2;  with psifull (MIXED) also partial2[chunk][col] = sum B[row][col] * psifull[i(col)][row],
// i(col) = col % L for a covariant GP (column (s,l,i)), the problem index z for a variant one.
__global__ void colnorm_partial_kernel(const double* __restrict__ B, long ldb, long strideB, int ncols, double* __restrict__ partial, long stride_partial,
                                       const double* __restrict__ psifull, long n_pad, int L, int chol_batch, double* __restrict__ partial2) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x, chunk = blockIdx.y, z = blockIdx.z;
  if (col >= ncols) return;
  const double* b = B + (long)z * strideB + (long)chunk * 128 * ldb + col;
  const double* pf = psifull ? psifull + (long)(chol_batch == 1 ? col % L : z) * n_pad + (long)chunk * 128 : nullptr;
  double acc = 0.0, acc2 = 0.0;
#pragma unroll 8
  for (int r = 0; r < 128; ++r) {
    const double v = b[(long)r * ldb];
    acc = fma(v, v, acc);
    if (pf) acc2 = fma(v, pf[r], acc2);
  }
  partial[(long)z * stride_partial + (long)chunk * ncols + col] = acc;
  if (pf) partial2[(long)z * stride_partial + (long)chunk * ncols + col] = acc2;
}

// psifull[i][row] = solved column (s = 0, l = i, i) of B: the psi factor of the FULL model, kept for the MIXED dots of every later chunk
__global__ void sobol_error_keep_psifull_kernel(const double* __restrict__ B, long ldb, long strideB, int L, long n_pad, int chol_batch,
                                                double* __restrict__ psifull) {
  const int i = blockIdx.y;
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_pad) return;
  psifull[(long)i * n_pad + row] = chol_batch == 1 ? B[row * ldb + (long)i * L + i] : B[(long)i * strideB + row * ldb + i];
}

// W[s][l][i] = Wraw[l][i] + Wraw[i][l],  Wraw[l][i] = (R[s][l][i] - |psi_li|^2) * (1 + [l == i]);  WMm likewise from Rm and the psi^FULL . psi dots
__global__ void sobol_error_W_kernel(const double* __restrict__ R, const double* __restrict__ partial, long stride_partial, int chunks, int ncols,
                                     int L, int ns, int chol_batch, ErrMap map, double* __restrict__ W, const double* __restrict__ Rm,
                                     const double* __restrict__ partial2, double* __restrict__ WMm) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ns * L * L) return;
  const int s = e / (L * L), li = e - s * L * L, l = li / L, i = li - l * L;
  auto raw = [&](const double* Rv, const double* part, int a, int b) {
    double psi2 = 0.0;
    const long col = chol_batch == 1 ? ((long)s * L + a) * L + b : (long)s * L + a;
    const double* pp = part + (chol_batch == 1 ? 0 : (long)b * stride_partial) + col;
    for (int k = 0; k < chunks; ++k) psi2 += pp[(long)k * ncols];
    return (Rv[((long)s * L + a) * L + b] - psi2) * (a == b ? 2.0 : 1.0);
  };
  if (W) W[((long)map.dest[s] * L + l) * L + i] = raw(R, partial, l, i) + raw(R, partial, i, l);
  if (WMm) WMm[((long)map.dest[s] * L + l) * L + i] = raw(Rm, partial2, l, i) + raw(Rm, partial2, i, l);
}

namespace {
struct ErrLayout {
  size_t coef, pre, parts, B, partial, partial2, R, Rm, ct, psifull, sbinv, total;
  int RCH, RC, T, chunk_slices, ncols, kinds;
  bool use_sbinv;
  long ldb, strideB;
};
inline size_t al(size_t b) { return (b + 255) / 256 * 256; }
int err_row_chunk(int N, int M) {
  int rch = (int)((144 * 1024) / (8L * M)) / 64 * 64;
  if (rch > 1024) rch = 1024;
  if (rch < 64) rch = 64;
  const int n64 = (N + 63) / 64 * 64;
  return rch > n64 ? n64 : rch;
}
ErrLayout err_layout(int N, int M, int L, int nslices, int n_pad, int chol_batch, bool mixed) {
  ErrLayout o{};
  o.RCH = err_row_chunk(N, M);
  o.RC = (N + o.RCH - 1) / o.RCH;
  o.T = (N + EC - 1) / EC;
  o.chunk_slices = nslices < 32 ? nslices : 32;
  o.kinds = mixed ? 3 : 2;
  const int J = o.kinds * L * L;
  const int per_problem_cols = chol_batch == 1 ? o.chunk_slices * L * L : o.chunk_slices * L;
  o.ncols = round_up(per_problem_cols, TILE);
  o.ldb = o.ncols;
  o.strideB = (long)n_pad * o.ldb;
  size_t off = 0;
  o.coef = off; off += al((size_t)4 * J * M * sizeof(double));
  o.pre = off; off += al((size_t)L * sizeof(double));
  {
    const size_t general = (size_t)J * o.RC * o.chunk_slices * o.T * EC;                                        // [J][RC][chunk][T*64]
    const size_t sweep = M <= 20 ? (size_t)J * ((N + SW_RCH - 1) / SW_RCH) * 3 * M * ((N + SW_EC - 1) / SW_EC) * SW_EC : 0;   // [J][RC'][3M][T'*32]
    o.parts = off; off += al((general > sweep ? general : sweep) * sizeof(double));
  }
  o.B = off; off += al((size_t)chol_batch * o.strideB * sizeof(double));
  o.partial = off; off += al((size_t)chol_batch * (n_pad / 128) * o.ncols * sizeof(double));
  o.R = off; off += al((size_t)o.chunk_slices * L * L * sizeof(double));
  if (mixed) {
    o.partial2 = off; off += al((size_t)chol_batch * (n_pad / 128) * o.ncols * sizeof(double));
    o.Rm = off; off += al((size_t)o.chunk_slices * L * L * sizeof(double));
    o.ct = off; off += al((size_t)L * N * sizeof(double));
    o.psifull = off; off += al((size_t)L * n_pad * sizeof(double));
  }
  // one large factor, few columns: triangular solves through inverted diagonal super-blocks (chol.cu: trsm_lower_fwd_sbinv); RC_TRSM_SBINV=0 disables
  static const bool sbinv_on = [] { const char* e = getenv("RC_TRSM_SBINV"); return !e || atoi(e) != 0; }();
  o.use_sbinv = sbinv_on && chol_batch == 1 && n_pad >= 4096 && o.ncols <= 1024;
  o.sbinv = off;
  if (o.use_sbinv) off += al(trsm_sbinv_workspace_doubles(n_pad, o.ncols) * sizeof(double));
  o.total = off;
  return o;
}
}  // namespace

size_t sobol_error_workspace_bytes(int N, int M, int L, int nslices, int n_pad, int chol_batch, int mixed) {
  return err_layout(N, M, L, nslices, n_pad, chol_batch, mixed != 0).total;
}

int sobol_error(const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0, const double* g0KY, int L,
                const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const double* dinv, const unsigned long long* masks,
                int nslices, void* work, double* V, double* W, double* WMm, cudaStream_t st) {
  RC_REQUIRE(M >= 1 && M <= 64, -2, "sobol_error: M=%d out of range [1,64]", M);
  RC_REQUIRE(chol_batch == 1 || chol_batch == L, -2, "sobol_error: chol_batch must be 1 (covariant) or L (variant)");
  RC_REQUIRE(n_pad >= (chol_batch == 1 ? L * N : N), -2, "sobol_error: n_pad too small");
  const bool mixed = WMm != nullptr;
  const ErrLayout lay = err_layout(N, M, L, nslices, n_pad, chol_batch, mixed);
  char* base = static_cast<char*>(work);
  double* coef = reinterpret_cast<double*>(base + lay.coef);
  double* pre = reinterpret_cast<double*>(base + lay.pre);
  double* parts = reinterpret_cast<double*>(base + lay.parts);
  double* B = reinterpret_cast<double*>(base + lay.B);
  double* partial = reinterpret_cast<double*>(base + lay.partial);
  double* R = reinterpret_cast<double*>(base + lay.R);
  double* partial2 = mixed ? reinterpret_cast<double*>(base + lay.partial2) : nullptr;
  double* Rm = mixed ? reinterpret_cast<double*>(base + lay.Rm) : nullptr;
  double* ct = mixed ? reinterpret_cast<double*>(base + lay.ct) : nullptr;
  double* psifull = mixed ? reinterpret_cast<double*>(base + lay.psifull) : nullptr;
  const int J = lay.kinds * L * L;
  double* sbwork = lay.use_sbinv ? reinterpret_cast<double*>(base + lay.sbinv) : nullptr;
  double* sbT = lay.use_sbinv ? sbwork + trsm_sbinv_workspace_doubles(n_pad, lay.ncols) - (size_t)1024 * lay.ncols : nullptr;
  if (lay.use_sbinv) {
    int rc0 = trsm_sbinv_prepare(Achol, n_pad, ld, dinv, sbwork, st);
    if (rc0) return rc0;
  }
  RC_ENSURE_SMEM(sobol_error_matvec_kernel, 220 * 1024);
  sobol_error_coeff_kernel<<<(J * M + 255) / 256, 256, 0, st>>>(Phi, Lam, F, L, M, lay.kinds, coef, pre);
  RC_LAUNCH_OK();
  if (mixed) {
    sobol_error_mixed_weight_kernel<<<dim3((N + 255) / 256, L), 256, 0, st>>>(X, N, M, Phi, Lam, g0KY, ct);
    RC_LAUNCH_OK();
  }
  const size_t smem = (size_t)(4 * M + (long)M * lay.RCH + lay.RCH + 2 * M * EC + lay.RCH + EC + 16 * EC) * sizeof(double);
  RC_REQUIRE(smem <= 220 * 1024, -2, "sobol_error: shared memory %zu too large", smem);
  const int chunks = n_pad / 128;
  const long stride_partial = (long)chunks * lay.ncols;
  const unsigned long long full_mask = (M >= 64) ? ~0ull : ((1ull << M) - 1ull);
  // Structured slices (singles, prefixes, suffixes, full, empty) share ONE sweep-form launch; general subsets take the summed-exponent kernel.
  // RC_SOBOL_SWEEP=park selects the round-1 form of the sweep kernel (k_m parked in shared memory, one partial per 512-row chunk, M <= 12)
  static const bool reg_form = [] { const char* e = getenv("RC_SOBOL_SWEEP"); return !e || e[0] != 'p'; }();
  const int sweep_max_M = reg_form ? 20 : 12;
  std::vector<int> structured, scol, general;
  for (int s = 0; s < nslices; ++s) {
    const int k = M <= sweep_max_M ? sobol_sweep_index(masks[s], M) : -1;
    if (k >= 0) {
      structured.push_back(s);
      scol.push_back(k);
    } else {
      general.push_back(s);
    }
  }
  // keep_full: this chunk is the internal one holding only the full model (MIXED): no V / W output, its psi_ii columns are kept
  // mode 0: V / W (/ WMm) of the chunk;  1: the internal chunk that holds only the full model (MIXED): no output, its psi_ii columns are kept;
  // 2: a chunk whose FIRST slice is the full model: its psi_ii columns are kept, then everything of mode 0 (saves the separate solve of mode 1)
  auto finish_chunk = [&](const ErrMap& map, int ns, int nv, int RC, long ncols_u, int mode) -> int {
    const bool keep_full = mode == 1;
    RC_CUDA_OK(cudaMemsetAsync(B, 0, (size_t)chol_batch * lay.strideB * sizeof(double), st));
    sobol_error_gather_kernel<<<dim3(ns, L * L), 256, 0, st>>>(parts, L, N, RC, nv, ncols_u, g0KY, g0, pre, chol_batch, B, lay.ldb, lay.strideB, map,
                                                               keep_full ? nullptr : V, R, keep_full ? nullptr : Rm);
    RC_LAUNCH_OK();
    int rc = lay.use_sbinv ? trsm_lower_fwd_sbinv(Achol, n_pad, ld, dinv, sbwork, sbT, B, lay.ncols, lay.ldb, st)
                           : trsm_lower_fwd(Achol, n_pad, ld, strideA, chol_batch, dinv, B, lay.ncols, lay.ldb, lay.strideB, st);
    if (rc) return rc;
    if (mode != 0) {
      sobol_error_keep_psifull_kernel<<<dim3((n_pad + 255) / 256, L), 256, 0, st>>>(B, lay.ldb, lay.strideB, L, n_pad, chol_batch, psifull);
      RC_LAUNCH_OK();
      if (keep_full) return 0;
    }
    colnorm_partial_kernel<<<dim3((lay.ncols + 127) / 128, chunks, chol_batch), 128, 0, st>>>(B, lay.ldb, lay.strideB, lay.ncols, partial, stride_partial,
                                                                                                 psifull, n_pad, L, chol_batch, partial2);
    RC_LAUNCH_OK();
    sobol_error_W_kernel<<<(ns * L * L + 127) / 128, 128, 0, st>>>(R, partial, stride_partial, chunks, lay.ncols, L, ns, chol_batch, map, W, Rm, partial2,
                                                                   WMm);
    RC_LAUNCH_OK();
    return 0;
  };
  auto general_args = [&](int ns) {
    ErrMatvecArgs a{};
    a.X = X; a.N = N; a.M = M; a.coef = coef; a.c = g0KY; a.c2 = ct; a.L = L; a.J = J; a.T = lay.T; a.RC = lay.RC; a.RCH = lay.RCH; a.ns = ns; a.parts = parts;
    return a;
  };
  const bool sweep_form = M <= sweep_max_M && (!structured.empty() || mixed);
  ErrMatvecArgs sw{};
  if (sweep_form) {
    sw.X = X; sw.N = N; sw.M = M; sw.coef = coef; sw.c = g0KY; sw.c2 = ct; sw.L = L; sw.J = J;
    sw.RCH = SW_RCH; sw.RC = (N + SW_RCH - 1) / SW_RCH; sw.T = (N + SW_EC - 1) / SW_EC; sw.ns = 3 * M; sw.parts = parts;
    int rc;
    if (reg_form) {
      sw.RC = 1; sw.RCH = SWR_ROWS;
      switch (M) {
        case 1: rc = launch_error_sweep_reg<1, 4>(sw, st); break;
        case 2: rc = launch_error_sweep_reg<2, 4>(sw, st); break;
        case 3: rc = launch_error_sweep_reg<3, 4>(sw, st); break;
        case 4: rc = launch_error_sweep_reg<4, 4>(sw, st); break;
        case 5: rc = launch_error_sweep_reg<5, 4>(sw, st); break;
        case 6: rc = launch_error_sweep_reg<6, 4>(sw, st); break;
        case 7: rc = launch_error_sweep_reg<7, 4>(sw, st); break;
        case 8: rc = launch_error_sweep_reg<8, 4>(sw, st); break;
        case 9: rc = launch_error_sweep_reg<9, 2>(sw, st); break;
        case 10: rc = launch_error_sweep_reg<10, 2>(sw, st); break;
        case 11: rc = launch_error_sweep_reg<11, 2>(sw, st); break;
        case 12: rc = launch_error_sweep_reg<12, 2>(sw, st); break;
        case 13: rc = launch_error_sweep_reg<13, 2>(sw, st); break;
        case 14: rc = launch_error_sweep_reg<14, 2>(sw, st); break;
        case 15: rc = launch_error_sweep_reg<15, 2>(sw, st); break;
        case 16: rc = launch_error_sweep_reg<16, 2>(sw, st); break;
        case 17: rc = launch_error_sweep_reg<17, 2>(sw, st); break;
        case 18: rc = launch_error_sweep_reg<18, 2>(sw, st); break;
        case 19: rc = launch_error_sweep_reg<19, 2>(sw, st); break;
        default: rc = launch_error_sweep_reg<20, 2>(sw, st); break;
      }
    } else {
      rc = M <= 4 ? launch_error_sweep<4>(sw, st) : M <= 8 ? launch_error_sweep<8>(sw, st) : launch_error_sweep<12>(sw, st);
    }
    if (rc) return rc;
  }
  // MIXED needs psi^FULL_ii for the dots of every chunk.  If the full model is itself one of the structured slices asked for (gsa.models.GSA always
  // appends it), it is moved to the front of the list and its columns are kept from the first chunk's solve; otherwise it gets a chunk of its own.
  bool full_in_first_chunk = false;
  if (mixed && sweep_form) {
    const int full_col = sobol_sweep_index(full_mask, M);
    for (size_t q = 0; q < structured.size(); ++q)
      if (scol[q] == full_col) {
        std::swap(structured[0], structured[q]);
        std::swap(scol[0], scol[q]);
        full_in_first_chunk = true;
        break;
      }
  }
  if (mixed && !full_in_first_chunk) {   // the full model first
    ErrMap map{};
    map.dest[0] = 0;
    int rc;
    if (sweep_form) {
      map.col[0] = sobol_sweep_index(full_mask, M);
      if ((rc = finish_chunk(map, 1, 3 * M, sw.RC, (long)sw.T * SW_EC, 1))) return rc;
    } else {
      ErrMatvecArgs a = general_args(1);
      a.masks[0] = full_mask;
      map.col[0] = 0;
      sobol_error_matvec_kernel<<<dim3(lay.T * lay.RC, J), ETHREADS, smem, st>>>(a);
      RC_LAUNCH_OK();
      if ((rc = finish_chunk(map, 1, 1, lay.RC, (long)lay.T * EC, 1))) return rc;
    }
  }
  for (size_t i0 = 0; i0 < structured.size(); i0 += lay.chunk_slices) {
    const int ns = (int)std::min<size_t>(lay.chunk_slices, structured.size() - i0);
    ErrMap map{};
    for (int s = 0; s < ns; ++s) {
      map.col[s] = scol[i0 + s];
      map.dest[s] = structured[i0 + s];
    }
    int rc = finish_chunk(map, ns, 3 * M, sw.RC, (long)sw.T * SW_EC, (i0 == 0 && full_in_first_chunk) ? 2 : 0);
    if (rc) return rc;
  }
  for (size_t i0 = 0; i0 < general.size(); i0 += lay.chunk_slices) {
    const int ns = (int)std::min<size_t>(lay.chunk_slices, general.size() - i0);
    ErrMatvecArgs a = general_args(ns);
    ErrMap map{};
    for (int s = 0; s < ns; ++s) {
      a.masks[s] = masks[general[i0 + s]];
      map.col[s] = s;
      map.dest[s] = general[i0 + s];
    }
    sobol_error_matvec_kernel<<<dim3(lay.T * lay.RC, J), ETHREADS, smem, st>>>(a);
    RC_LAUNCH_OK();
    int rc = finish_chunk(map, ns, ns, lay.RC, (long)lay.T * EC, 0);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace rc
