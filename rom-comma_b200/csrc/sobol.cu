// Closed-form Gaussian-integral Sobol contractions (FP64, sm_100a).
// Replaces ClosedSobol._calibrate / _V / marginalize (romcomma/gsa/calibrators.py:49-97) and the Gaussian-ratio chain of
// romcomma/gsa/base.py:92-126 with one fused integrand+contraction kernel: for every pair of "output rows" a = (l,L'),
// b = (j,J') and every 64x64 tile of the (N,n) pair space it evaluates, for a whole list of marginal subsets at once,
//
//   H = exp( SU_s[N] + SV_s[n] + sum_{m in s} gamma_m x_Nm y_nm ),   V_s[l,j] += c_aN * H * c_bn
//
// with p = Phi[a,m], q = Phi[b,m], psi = 1 - p q, gamma = p q / psi,
//   SU_s[N] = sum_{m in s} ( -1/2 gamma p x_Nm^2 - 1/2 log psi ),  SV_s[n] = sum_{m in s} -1/2 gamma q y_nm^2
// (everything is summed before the single exp because the exponent may be positive).  The
// (l,L',N,j,J',n,m) tensor the reference materialises (17 GB per slice at N=4096, L=4, M=8) is never formed; HBM traffic is
// O(N*M).  The symmetry H[(a,N),(b,n)] = H[(b,n),(a,N)] halves the work.  Subsets are bit masks over the M inputs, so
// non-contiguous subsets (the all-subsets sweep) cost the same as the reference's contiguous slices.
// Partial sums are written per CTA and reduced in a fixed order: bitwise reproducible, no atomics.
#include "sobol.h"
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <map>
#include <vector>

namespace rc {

constexpr int ST = 64;
constexpr int STHREADS = 256;

// ---- prepare: Phi, g0, g0KY (gsa/calibrators.py:82-92,99-109,134-138) ---------------------------------------------
// a = l*Lp + k.  diag F: Lp = 1, Lambda2 = Lam[l]^2, Fa = F[l].  full F: Lp = L, Lambda2 = Lam[l]*Lam[k], Fa = F[l,k], KinvY row k.
__global__ void sobol_prepare_kernel(const double* __restrict__ X, int N, int M, const double* __restrict__ Lam, const double* __restrict__ F,
                                     const double* __restrict__ KinvY, int L, int Lp, double* __restrict__ Phi, double* __restrict__ g0,
                                     double* __restrict__ g0KY) {
  extern __shared__ double sm[];
  double* phi = sm;   // [M]
  __shared__ double pref;
  const int a = blockIdx.y, l = a / Lp, k = a - l * Lp;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double lam2 = (Lp == 1) ? Lam[l * M + m] * Lam[l * M + m] : Lam[l * M + m] * Lam[k * M + m];
    phi[m] = 1.0 / (lam2 + 1.0);
    if (blockIdx.x == 0) Phi[a * M + m] = phi[m];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double prod = 1.0;
    for (int m = 0; m < M; ++m) {
      const double lam2 = (Lp == 1) ? Lam[l * M + m] * Lam[l * M + m] : Lam[l * M + m] * Lam[k * M + m];
      prod *= lam2 * phi[m];
    }
    pref = sqrt(prod) * ((Lp == 1) ? F[l] : F[l * L + k]);
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double e = 0.0;
  for (int m = 0; m < M; ++m) {
    const double x = X[(long)n * M + m];
    e = fma(x * x, phi[m], e);
  }
  const double g = exp(-0.5 * e) * pref;
  g0[(long)a * N + n] = g;
  g0KY[(long)a * N + n] = g * KinvY[(long)((Lp == 1) ? l : k) * N + n];
}

// g0KY[l, :, :] -= mean over (L', N)   (calibrators.py:90)
__global__ void sobol_centre_kernel(double* __restrict__ g0KY, int N, int Lp) {
  __shared__ double red[32];
  __shared__ double mean;
  const long cnt = (long)Lp * N;
  double* row = g0KY + (long)blockIdx.x * cnt;
  double s = 0.0;
  for (long i = threadIdx.x; i < cnt; i += blockDim.x) s += row[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) mean = s / (double)cnt;
  __syncthreads();
  for (long i = threadIdx.x; i < cnt; i += blockDim.x) row[i] -= mean;
}

int sobol_prepare(const double* X, int N, int M, const double* Lam, const double* F, const double* KinvY, int L, int is_F_diagonal, double* Phi,
                  double* g0, double* g0KY, cudaStream_t st) {
  const int Lp = is_F_diagonal ? 1 : L;
  sobol_prepare_kernel<<<dim3((N + 255) / 256, L * Lp), 256, M * sizeof(double), st>>>(X, N, M, Lam, F, KinvY, L, Lp, Phi, g0, g0KY);
  RC_LAUNCH_OK();
  sobol_centre_kernel<<<L, 1024, 0, st>>>(g0KY, N, Lp);
  RC_LAUNCH_OK();
  return 0;
}

// ---- the contraction ----------------------------------------------------------------------------------------------
struct SobolPairArgs {
  const double* X; int N, M;
  const double* Phi;   // [P][M]
  const double* c;     // [P][N] coefficients (g0KY); the same on both sides, which is what makes H's symmetry usable
  int P, T;            // T = ceil(N / 64)
  int ns;
  double* parts;       // [npairs][T*T][ns]
  int part, nparts;    // this call evaluates the row tiles ti with ti % nparts == part (multi-GPU: partial V, summed by the caller)
  int chunk;           // register form of the sweep kernel: column tiles per CTA
  unsigned long long masks[SOBOL_MAX_SLICES];
};

__global__ void __launch_bounds__(STHREADS) sobol_pair_kernel(SobolPairArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int M = p.M, ns = p.ns;
  double* gam = sm;                 // [M]
  double* cu = gam + M;             // -1/2 gamma p
  double* cv = cu + M;              // -1/2 gamma q
  double* lp = cv + M;              // -1/2 log psi
  double* xs = lp + M;              // [M][64]  x (rows)
  double* gx = xs + M * ST;         // [M][64]  gamma_m * x
  double* yy = gx + M * ST;         // [M][64]  y (columns)
  double* SU = yy + M * ST;         // [ns][64]
  double* SV = SU + ns * ST;        // [ns][64]
  double* cr = SV + ns * ST;        // [64]
  double* cc = cr + ST;             // [64]
  double* wpart = cc + ST;          // [8][ns]

  // pair (a >= b)
  const int pidx = blockIdx.y;
  int a = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= pidx) ++a;
  while (a * (a + 1) / 2 > pidx) --a;
  const int b = pidx - a * (a + 1) / 2;
  // only this rank's row tiles are launched: local row k of the grid is the global row tile part + k * nparts
  const int ti = p.part + (int)(blockIdx.x / p.T) * p.nparts, tj = blockIdx.x % p.T;
  double* out = p.parts + ((long)pidx * gridDim.x + blockIdx.x) * ns;
  if (a == b && tj > ti) {   // covered by the mirrored tile (weight 2)
    for (int s = threadIdx.x; s < ns; s += STHREADS) out[s] = 0.0;
    return;
  }
  __shared__ double etab[32];
  exp_table_fill(etab);
  for (int m = threadIdx.x; m < M; m += STHREADS) {
    const double pp = p.Phi[a * M + m], qq = p.Phi[b * M + m];
    const double psi = 1.0 - pp * qq, g = pp * qq / psi;
    gam[m] = g;
    cu[m] = -0.5 * g * pp;
    cv[m] = -0.5 * g * qq;
    lp[m] = -0.5 * log(psi);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ST * M; e += STHREADS) {
    const int r = e / M, m = e - r * M;
    const int gi = ti * ST + r, gj = tj * ST + r;
    const double x = gi < p.N ? p.X[(long)gi * M + m] : 0.0;
    const double y = gj < p.N ? p.X[(long)gj * M + m] : 0.0;
    xs[m * ST + r] = x;
    gx[m * ST + r] = gam[m] * x;
    yy[m * ST + r] = y;
  }
  for (int r = threadIdx.x; r < ST; r += STHREADS) {
    const int gi = ti * ST + r, gj = tj * ST + r;
    cr[r] = gi < p.N ? p.c[(long)a * p.N + gi] : 0.0;
    cc[r] = gj < p.N ? p.c[(long)b * p.N + gj] : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ns * ST; e += STHREADS) {
    const int s = e / ST, r = e - s * ST;
    const unsigned long long mask = p.masks[s];
    double su = 0.0, sv = 0.0;
    for (int m = 0; m < M; ++m)
      if ((mask >> m) & 1ull) {
        const double x = xs[m * ST + r];
        const double y = yy[m * ST + r];
        su += fma(cu[m] * x, x, lp[m]);
        sv = fma(cv[m] * y, y, sv);
      }
    SU[s * ST + r] = su;
    SV[s * ST + r] = sv;
  }
  __syncthreads();

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double crr[4], ccc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    crr[u] = cr[ty * 4 + u];
    ccc[u] = cc[tx * 4 + u];
  }
  for (int s = 0; s < ns; ++s) {
    const unsigned long long mask = p.masks[s];
    double e[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) e[u][v] = SU[s * ST + ty * 4 + u] + SV[s * ST + tx * 4 + v];
    for (int m = 0; m < M; ++m) {
      if (!((mask >> m) & 1ull)) continue;
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = gx[m * ST + ty * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) bv[v] = yy[m * ST + tx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) e[u][v] = fma(av[u], bv[v], e[u][v]);
    }
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double rowacc = 0.0;
#pragma unroll
      for (int v = 0; v < 4; ++v) rowacc = fma(ccc[v], exp_tab(fmin(e[u][v], 708.0), etab), rowacc);
      acc = fma(crr[u], rowacc, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) wpart[warp * ns + s] = acc;
  }
  __syncthreads();
  const double wgt = (a == b && ti != tj) ? 2.0 : 1.0;
  for (int s = threadIdx.x; s < ns; s += STHREADS) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += wpart[w * ns + s];
    out[s] = v * wgt;
  }
}


// ---- sweep form -----------------------------------------------------------------------------------------------------
// The integrand factorises over the inputs:  H_s[N,n] = prod_{m in s} h_m[N,n],  h_m = exp( su_m[N] + sv_m[n] + gamma_m x_Nm y_nm ).
// For the slice families the reference sweeps (gsa/models.py:77-90: first order [m:m+1], closed [0:m+1], total-complement [m+1:M], plus the
// full and the empty slice) that means M exps per sample pair instead of 3M+1: singles are the h_m themselves, closed slices their
// running prefix products and complements the suffix products.  (Single-input exponents are bounded by gamma (1-p) x^2 with |x| <= 7.04 after
// the probit normalisation, far from the FP64 range limits; a factor that underflows gives 0 where the summed form gives < 1e-300.)
// Output per (pair of output rows, tile): 3M values  { F[m] m<M | P[k] k=1..M | S[k] k=1..M-1 | E },
//   F[m] = sum w h_m,  P[k] = sum w prod_{j<k} h_j,  S[k] = sum w prod_{j>=k} h_j,  E = sum w   (w = c_aN c_bn).
// Thread (ty,tx) of the 16x16 block evaluates RU x 2 pairs per pass.  The 3M accumulators stay in registers; the h values of a pass are
// parked in shared memory (thread-private columns, conflict-free) between the prefix and the suffix scan so that two CTAs fit on an SM:
// the kernel is bound by the latency of the exp dependency chains, not by issue slots.
template <int MAXM, int RU>
__global__ void __launch_bounds__(STHREADS, 2) sobol_sweep_kernel(SobolPairArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int M = p.M, nv = 3 * M;
  double* gam = sm;                 // [M]
  double* cu = gam + M;             // -1/2 gamma p
  double* cv = cu + M;              // -1/2 gamma q
  double* lp = cv + M;              // -1/2 log psi
  double* gx = lp + M;              // [M][64]  gamma_m * x (rows)
  double* yy = gx + M * ST;         // [M][64]  y (columns)
  double* su = yy + M * ST;         // [M][64]  cu_m x^2 + lp_m
  double* sv = su + M * ST;         // [M][64]  cv_m y^2
  double* cr = sv + M * ST;         // [64]
  double* cc = cr + ST;             // [64]
  double* wpart = cc + ST;          // [8][3*MAXM]
  double* hs = wpart + 8 * 3 * MAXM;   // [M][2*RU][256]  h of the current pass

  const int pidx = blockIdx.y;
  int a = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= pidx) ++a;
  while (a * (a + 1) / 2 > pidx) --a;
  const int b = pidx - a * (a + 1) / 2;
  const int ti = p.part + (int)(blockIdx.x / p.T) * p.nparts, tj = blockIdx.x % p.T;   // only this rank's row tiles are launched
  double* out = p.parts + ((long)pidx * gridDim.x + blockIdx.x) * nv;
  if (a == b && tj > ti) {   // covered by the mirrored tile (weight 2)
    for (int s = threadIdx.x; s < nv; s += STHREADS) out[s] = 0.0;
    return;
  }
  for (int m = threadIdx.x; m < M; m += STHREADS) {
    const double pp = p.Phi[a * M + m], qq = p.Phi[b * M + m];
    const double psi = 1.0 - pp * qq, g = pp * qq / psi;
    gam[m] = g;
    cu[m] = -0.5 * g * pp;
    cv[m] = -0.5 * g * qq;
    lp[m] = -0.5 * log(psi);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ST * M; e += STHREADS) {
    const int r = e / M, m = e - r * M;
    const int gi = ti * ST + r, gj = tj * ST + r;
    const double x = gi < p.N ? p.X[(long)gi * M + m] : 0.0;
    const double y = gj < p.N ? p.X[(long)gj * M + m] : 0.0;
    gx[m * ST + r] = gam[m] * x;
    yy[m * ST + r] = y;
    su[m * ST + r] = fma(cu[m] * x, x, lp[m]);
    sv[m * ST + r] = cv[m] * y * y;
  }
  for (int r = threadIdx.x; r < ST; r += STHREADS) {
    const int gi = ti * ST + r, gj = tj * ST + r;
    cr[r] = gi < p.N ? p.c[(long)a * p.N + gi] : 0.0;     // rows/columns beyond N carry zero weight
    cc[r] = gj < p.N ? p.c[(long)b * p.N + gj] : 0.0;
  }
  __syncthreads();

  constexpr int NP = 2 * RU;   // pairs per thread per pass
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double accF[MAXM], accP[MAXM], accS[MAXM], accE = 0.0;   // accP[k-1] = P[k]; accS[k] = S[k] (accS[0] unused)
#pragma unroll
  for (int m = 0; m < MAXM; ++m) accF[m] = accP[m] = accS[m] = 0.0;
#pragma unroll 1
  for (int pass = 0; pass < (ST / (16 * RU)) * 2; ++pass) {
    const int r0 = (pass >> 1) * 16 * RU + ty * RU, c0 = (pass & 1) * 32 + tx * 2;
    double w[NP], run[NP];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      w[2 * u] = cr[r0 + u] * cc[c0];
      w[2 * u + 1] = cr[r0 + u] * cc[c0 + 1];
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      run[q] = w[q];
      accE += w[q];
    }
    double* hme = hs + threadIdx.x;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {     // exps, singles and prefixes
      if (m < M) {
        const double b0 = yy[m * ST + c0], b1 = yy[m * ST + c0 + 1], v0 = sv[m * ST + c0], v1 = sv[m * ST + c0 + 1];
        double h[NP];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const double a0 = gx[m * ST + r0 + u], u0 = su[m * ST + r0 + u];
          h[2 * u] = exp_pairwise(fma(a0, b0, u0 + v0));
          h[2 * u + 1] = exp_pairwise(fma(a0, b1, u0 + v1));
        }
        double f = 0.0, pr = 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          if (m >= 1) hme[(m * NP + q) * STHREADS] = h[q];
          f = fma(w[q], h[q], f);
          run[q] *= h[q];
          pr += run[q];
        }
        accF[m] += f;
        accP[m] += pr;
      }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) run[q] = w[q];
#pragma unroll
    for (int m = MAXM - 1; m >= 1; --m) {   // suffixes S[m] = prod_{j >= m} h_j
      if (m < M) {
        double sf = 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          run[q] *= hme[(m * NP + q) * STHREADS];
          sf += run[q];
        }
        accS[m] += sf;
      }
    }
  }
  // warp reductions in a fixed order, then one partial per CTA
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    if (m < M) {
      const double f = warp_sum(accF[m]), pr = warp_sum(accP[m]);
      if (lane == 0) {
        wpart[warp * 3 * MAXM + m] = f;
        wpart[warp * 3 * MAXM + MAXM + m] = pr;
      }
      if (m >= 1) {
        const double sf = warp_sum(accS[m]);
        if (lane == 0) wpart[warp * 3 * MAXM + 2 * MAXM + m] = sf;
      }
    }
  }
  {
    const double e = warp_sum(accE);
    if (lane == 0) wpart[warp * 3 * MAXM + 2 * MAXM] = e;   // slot of the unused S[0]
  }
  __syncthreads();
  const double wgt = (a == b && ti != tj) ? 2.0 : 1.0;
  for (int s = threadIdx.x; s < nv; s += STHREADS) {
    // output index -> wpart slot:  F[m] -> m ; P[k] (s = M+k-1) -> MAXM+k-1 ; S[k] (s = 2M+k-1) -> 2MAXM+k ; E (s = 3M-1) -> 2MAXM
    int slot;
    if (s < M) slot = s;
    else if (s < 2 * M) slot = MAXM + (s - M);
    else if (s < 3 * M - 1) slot = 2 * MAXM + (s - 2 * M + 1);
    else slot = 2 * MAXM;
    double v = 0.0;
    for (int wdx = 0; wdx < 8; ++wdx) v += wpart[wdx * 3 * MAXM + slot];
    out[s] = v * wgt;
  }
}


// Round 2, second form of the sweep kernel (M <= 20): the h values of a pass stay in REGISTERS and the exp is the table form (exp_tab).
// The kernel above keeps two 256-thread CTAs per SM by parking h in shared memory; by ncu's wavefront count that parking, the 8-byte operand
// loads and the bank conflicts of the stride-2 column reads cost 8 shared-memory cycles per warp-exp against 10.6 cycles of the FP64 pipe - the
// two pipes were nearly co-critical, so a cheaper exp alone would only have moved the bound.  Here: 128 threads per CTA, three CTAs per SM
// (168 registers per thread: 3M accumulators + M x 2RU live h), row operands {gamma x, su} and column operands {y, sv} interleaved so that one
// 16-byte load fetches both (1.5 wavefronts per warp-exp), one table lookup per exp (2-4 wavefronts), 9 + 7 FP64 instructions per exp.
// Same tiles, same outputs, same per-CTA partial sums as sobol_sweep_kernel (the finish kernels do not know the difference).
// (Measured at M = 8: one row per thread, 128 registers and four CTAs per SM instead of two rows, 168 registers and three: 1.49 -> 1.67 ms.)
constexpr int SRTHREADS = 128;

// CTAs per SM the register allocation aims for: three (168 registers) up to M = 12, two (255) beyond - and at M = 8, where the looser allocation
// schedules better than the third CTA helps (1.48 -> 1.42 ms at cfg3; at M = 6, 7, 9, 10, 12 the same choice costs 3-9 %, tools/time_sweep_M.py).
constexpr int sweep_reg_min_blocks(int M) { return (M == 8 || M > 12) ? 2 : 3; }

template <int M, int RU>
__global__ void __launch_bounds__(SRTHREADS, sweep_reg_min_blocks(M)) sobol_sweep_reg_kernel(SobolPairArgs p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double etab[32];
  constexpr int nv = 3 * M, MP = (M + 1) & ~1;             // MP: M rounded up to even (16-byte alignment of what follows)
  double* gam = sm;                                        // [M]
  double* cu = gam + MP;                                   // -1/2 gamma p
  double* cv = cu + MP;                                    // -1/2 gamma q
  double* lp = cv + MP;                                    // -1/2 log psi
  double2* rowd = reinterpret_cast<double2*>(lp + MP);     // [M][64]  { gamma_m x, cu_m x^2 + lp_m }
  double2* cold = rowd + M * ST;                           // [M][64]  { y, cv_m y^2 }   of the current column tile
  double* cr = reinterpret_cast<double*>(cold + M * ST);   // [64]
  double* cc = cr + ST;                                    // [64]    column weights of the current tile (x 2 for a mirrored tile)
  double* wpart = cc + ST;                                 // [4][3*M]

  const int pidx = blockIdx.y;
  int a = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= pidx) ++a;
  while (a * (a + 1) / 2 > pidx) --a;
  const int b = pidx - a * (a + 1) / 2;
  // CTA = (pair of output rows, row tile, chunk of column tiles): the row data, the accumulators and the reduction are paid once per chunk
  const int nchunks = (p.T + p.chunk - 1) / p.chunk;
  const int ti = p.part + (int)(blockIdx.x / nchunks) * p.nparts, tj0 = (int)(blockIdx.x % nchunks) * p.chunk;
  const int tj1 = min(tj0 + p.chunk, a == b ? ti + 1 : p.T);        // a == b: tiles above the diagonal are covered by their mirror (weight 2)
  double* out = p.parts + ((long)pidx * gridDim.x + blockIdx.x) * nv;
  if (tj0 >= tj1) {
    for (int s = threadIdx.x; s < nv; s += SRTHREADS) out[s] = 0.0;
    return;
  }
  exp_table_fill(etab);
  for (int m = threadIdx.x; m < M; m += SRTHREADS) {
    const double pp = p.Phi[a * M + m], qq = p.Phi[b * M + m];
    const double psi = 1.0 - pp * qq, g = pp * qq / psi;
    gam[m] = g;
    cu[m] = -0.5 * g * pp;
    cv[m] = -0.5 * g * qq;
    lp[m] = -0.5 * log(psi);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ST * M; e += SRTHREADS) {
    const int r = e / M, m = e - r * M;
    const int gi = ti * ST + r;
    const double x = gi < p.N ? p.X[(long)gi * M + m] : 0.0;
    rowd[m * ST + r] = make_double2(gam[m] * x, fma(cu[m] * x, x, lp[m]));
  }
  for (int r = threadIdx.x; r < ST; r += SRTHREADS) {
    const int gi = ti * ST + r;
    cr[r] = gi < p.N ? p.c[(long)a * p.N + gi] : 0.0;     // rows/columns beyond N carry zero weight
  }

  constexpr int NP = 2 * RU;   // pairs per thread per pass
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // accF[m] = F[m] for m >= 1 (F[0] is P[1]);  accP[k-1] = P[k];  accS[k] = S[k] for 1 <= k <= M-2 (S[M-1] is F[M-1]);  accS[0] = E
  double accF[M], accP[M], accS[M];
#pragma unroll
  for (int m = 0; m < M; ++m) accF[m] = accP[m] = accS[m] = 0.0;
#pragma unroll 1
  for (int tj = tj0; tj < tj1; ++tj) {
    if (tj > tj0) __syncthreads();                         // everybody has finished with the previous column tile
    for (int e = threadIdx.x; e < ST * M; e += SRTHREADS) {
      const int r = e / M, m = e - r * M;
      const int gj = tj * ST + r;
      const double y = gj < p.N ? p.X[(long)gj * M + m] : 0.0;
      cold[m * ST + r] = make_double2(y, cv[m] * y * y);
    }
    for (int r = threadIdx.x; r < ST; r += SRTHREADS) {
      const int gj = tj * ST + r;
      cc[r] = gj < p.N ? p.c[(long)b * p.N + gj] * ((a == b && tj != ti) ? 2.0 : 1.0) : 0.0;
    }
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < (ST / (8 * RU)) * 2; ++pass) {
      const int r0 = (pass >> 1) * 8 * RU + ty * RU, c0 = (pass & 1) * 32 + tx * 2;
      double w[NP], run[NP], h[M][NP];
      {
        const double2 ccv = *reinterpret_cast<const double2*>(cc + c0);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const double cru = cr[r0 + u];
          w[2 * u] = cru * ccv.x;
          w[2 * u + 1] = cru * ccv.y;
        }
      }
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        run[q] = w[q];
        accS[0] += w[q];
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {     // exps, singles and prefixes
        const double2 d0 = cold[m * ST + c0], d1 = cold[m * ST + c0 + 1];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const double2 rw = rowd[m * ST + r0 + u];
          h[m][2 * u] = exp_tab(fma(rw.x, d0.x, rw.y + d0.y), etab);
          h[m][2 * u + 1] = exp_tab(fma(rw.x, d1.x, rw.y + d1.y), etab);
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          if (m >= 1) accF[m] = fma(w[q], h[m][q], accF[m]);
          run[q] *= h[m][q];
          accP[m] += run[q];
        }
      }
      if constexpr (M >= 3) {
#pragma unroll
        for (int q = 0; q < NP; ++q) run[q] = w[q] * h[M - 1][q];
#pragma unroll
        for (int m = M - 2; m >= 1; --m) {   // suffixes S[m] = prod_{j >= m} h_j
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            run[q] *= h[m][q];
            accS[m] += run[q];
          }
        }
      }
    }
  }
  // warp reductions in a fixed order, then one partial per CTA.  wpart slots: [0,M) F, [M,2M) P[k] at M+k-1, [2M,3M) S[k] at 2M+k, E at 2M
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const double pr = warp_sum(accP[m]);
    const double f = m >= 1 ? warp_sum(accF[m]) : pr;
    if (lane == 0) {
      wpart[warp * nv + m] = f;
      wpart[warp * nv + M + m] = pr;
    }
    if (m == 0 || (m >= 1 && m <= M - 2)) {
      const double sf = warp_sum(accS[m]);
      if (lane == 0) wpart[warp * nv + 2 * M + m] = sf;
    } else if (m == M - 1 && M >= 2) {
      if (lane == 0) wpart[warp * nv + 2 * M + m] = f;
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < nv; s += SRTHREADS) {
    // output index -> wpart slot:  F[m] -> m ; P[k] (s = M+k-1) -> M+k-1 ; S[k] (s = 2M+k-1) -> 2M+k ; E (s = 3M-1) -> 2M
    int slot;
    if (s < 2 * M) slot = s;
    else if (s < 3 * M - 1) slot = s + 1;
    else slot = 2 * M;
    double v = 0.0;
    for (int wdx = 0; wdx < SRTHREADS / 32; ++wdx) v += wpart[wdx * nv + slot];
    out[s] = v;
  }
}


// ---- all-subsets ("lattice") form -----------------------------------------------------------------------------------------
// The product form holds for EVERY subset:  H_S = prod_{m in S} h_m.  A block of the subset lattice = the 2^KL subsets that share their
// high mask bits hi (inputs KL..M-1) and run through all patterns lo of the KL low inputs:  H_(hi,lo) = H_hi * prod_{m in lo} h_m.  Per sample
// pair a thread evaluates ONE exp for the summed exponents of the hi inputs (times the weight c_aN c_bn), KL exps for the low inputs, and walks
// the binomial tree of the low patterns depth first - node = parent * h_b, accumulated as it is formed; the leaves (every pattern that contains
// input KL-1) are a single FMA - 2^KL accumulators in registers: (KL + 1) exps + ~1.5 FP64 instructions per subset, against one exp + 2|S| FMAs
// per (pair, subset) in sobol_pair_kernel (~33 instructions at M = 12).  CTA = (64-row tile ti, pair of output rows, lattice block); it walks
// all column tiles, so one partial per (pair, ti, block, lo).  Fixed reduction order, no atomics.
struct SobolLatticeArgs {
  const double* X; int N, M;
  const double* Phi;
  const double* c;
  int P, T;
  int nhi;                // lattice blocks of this launch (gridDim.z)
  double* parts;          // [npairs][T][nhi][2^KL]
  int part, nparts;
  unsigned hi[SOBOL_MAX_SLICES];   // high-bit pattern of each block (mask >> KL)
};

template <int KL, int S, int B>
__device__ __forceinline__ void lattice_walk(double (&acc)[1 << KL], const double (&h)[KL], double P) {
  if constexpr (B < KL) {
    constexpr int child = S | (1 << B);
    if constexpr (B == KL - 1) {
      acc[child] = fma(P, h[B], acc[child]);
    } else {
      const double Pc = P * h[B];
      acc[child] += Pc;
      lattice_walk<KL, child, B + 1>(acc, h, Pc);
      lattice_walk<KL, S, B + 1>(acc, h, P);
    }
  }
}

// The same sums with the last three inputs flattened (KL >= 4): their seven products q[t] are formed once per sample pair (4 multiplications),
// the binary tree runs over the first U = KL - 3 inputs only, and every node S of it (product P, the root included) feeds its seven
// descendants S | t << U with ONE FMA each.  A binary tree spends two instructions on a node that accumulates AND propagates and one on a leaf
// (1.5 per subset); flattened, KL = 6 costs 4 + 14 + 56 = 74 FP64 instructions per sample pair for the 63 subsets instead of 94.
template <int KL, int S, int B>
__device__ __forceinline__ void lattice_flat_children(double (&acc)[1 << KL], const double (&h)[KL], const double (&q)[8], double P);
template <int KL, int S, int B>
__device__ __forceinline__ void lattice_flat_node(double (&acc)[1 << KL], const double (&h)[KL], const double (&q)[8], double P) {
  constexpr int U = KL - 3;
#pragma unroll
  for (int t = 1; t < 8; ++t) acc[S | (t << U)] = fma(P, q[t], acc[S | (t << U)]);
  lattice_flat_children<KL, S, B>(acc, h, q, P);
}
template <int KL, int S, int B>
__device__ __forceinline__ void lattice_flat_children(double (&acc)[1 << KL], const double (&h)[KL], const double (&q)[8], double P) {
  constexpr int U = KL - 3;
  if constexpr (B < U) {
    constexpr int child = S | (1 << B);
    const double Pc = P * h[B];
    acc[child] += Pc;
    lattice_flat_node<KL, child, B + 1>(acc, h, q, Pc);
    lattice_flat_children<KL, S, B + 1>(acc, h, q, P);
  }
}
template <int KL>
__device__ __forceinline__ void lattice_accumulate(double (&acc)[1 << KL], const double (&h)[KL], double H) {
  acc[0] += H;
  if constexpr (KL >= 4) {
    constexpr int U = KL - 3;
    double q[8];
    q[1] = h[U];
    q[2] = h[U + 1];
    q[4] = h[U + 2];
    q[3] = q[1] * q[2];
    q[5] = q[1] * q[4];
    q[6] = q[2] * q[4];
    q[7] = q[3] * q[4];
    q[0] = 1.0;
    lattice_flat_node<KL, 0, 0>(acc, h, q, H);
  } else {
    lattice_walk<KL, 0, 0>(acc, h, H);
  }
}

template <int KL>
__global__ void __launch_bounds__(STHREADS, 1) sobol_lattice_kernel(SobolLatticeArgs p) {
  constexpr int NLO = 1 << KL;
  extern __shared__ __align__(16) double sm[];
  const int M = p.M;
  const unsigned hi = p.hi[blockIdx.z];
  int hidx[64], nh = 0;                       // the inputs of the hi pattern (registers; M <= 64)
  for (int m = KL; m < M; ++m)
    if ((hi >> (m - KL)) & 1u) hidx[nh++] = m;
  double* gam = sm;                 // [M]
  double* cu = gam + M;
  double* cv = cu + M;
  double* lp = cv + M;
  double* gxl = lp + M;             // [KL][64]   gamma_m x, low inputs (rows)
  double* sul = gxl + KL * ST;      // [KL][64]   cu_m x^2 + lp_m
  double* gxh = sul + KL * ST;      // [nh][64]   gamma_m x, hi inputs
  double* suh = gxh + (M - KL) * ST;   // [64]    sum over the hi inputs of cu_m x^2 + lp_m
  double* cr = suh + ST;            // [64]
  // column-side data of a tile, double buffered: the 64 loader threads fetch tile tj + 1 while everybody computes tile tj (one barrier per tile)
  constexpr int COLS = 2 * KL + 2;  // per buffer: yyl [KL][64], svl [KL][64], svh [64], cc [64], then yyh [M - KL][64]
  double* colbuf = cr + ST;
  const int col_stride = (COLS + (M - KL)) * ST;
  double* wpart = colbuf + 2 * col_stride;   // [8][NLO]

  const int pidx = blockIdx.y;
  int a = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= pidx) ++a;
  while (a * (a + 1) / 2 > pidx) --a;
  const int b = pidx - a * (a + 1) / 2;
  const int ti = p.part + (int)blockIdx.x * p.nparts, tid = threadIdx.x;       // only this rank's row tiles are launched
  double* out = p.parts + (((long)pidx * gridDim.x + blockIdx.x) * p.nhi + blockIdx.z) * NLO;
  __shared__ double etab[32];
  exp_table_fill(etab);
  for (int m = tid; m < M; m += STHREADS) {
    const double pp = p.Phi[a * M + m], qq = p.Phi[b * M + m];
    const double psi = 1.0 - pp * qq, g = pp * qq / psi;
    gam[m] = g;
    cu[m] = -0.5 * g * pp;
    cv[m] = -0.5 * g * qq;
    lp[m] = -0.5 * log(psi);
  }
  __syncthreads();
  for (int r = tid; r < ST; r += STHREADS) {
    const int gi = ti * ST + r;
    const bool live = gi < p.N;
    cr[r] = live ? p.c[(long)a * p.N + gi] : 0.0;
    double sh = 0.0;
    for (int k = 0; k < nh; ++k) {
      const int m = hidx[k];
      const double x = live ? p.X[(long)gi * M + m] : 0.0;
      gxh[k * ST + r] = gam[m] * x;
      sh += fma(cu[m] * x, x, lp[m]);
    }
    suh[r] = sh;
#pragma unroll
    for (int m = 0; m < KL; ++m) {
      const double x = (live && m < M) ? p.X[(long)gi * M + m] : 0.0;
      gxl[m * ST + r] = m < M ? gam[m] * x : 0.0;
      sul[m * ST + r] = m < M ? fma(cu[m] * x, x, lp[m]) : 0.0;
    }
  }
  double acc[NLO];
#pragma unroll
  for (int s = 0; s < NLO; ++s) acc[s] = 0.0;
  const int ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
  const int tj_end = (a == b) ? ti + 1 : p.T;   // a == b: the mirrored tile carries weight 2 instead
  auto load_columns = [&](int tj, double* buf) {
    double* yyl = buf;                 // [KL][64]   y, low inputs (columns)
    double* svl = yyl + KL * ST;       // [KL][64]   cv_m y^2
    double* svh = svl + KL * ST;       // [64]
    double* cc = svh + ST;             // [64]
    double* yyh = cc + ST;             // [nh][64]
    for (int r = tid; r < ST; r += STHREADS) {
      const int gj = tj * ST + r;
      const bool live = gj < p.N;
      cc[r] = live ? p.c[(long)b * p.N + gj] * ((a == b && tj != ti) ? 2.0 : 1.0) : 0.0;
      double sh = 0.0;
      for (int k = 0; k < nh; ++k) {
        const int m = hidx[k];
        const double y = live ? p.X[(long)gj * M + m] : 0.0;
        yyh[k * ST + r] = y;
        sh = fma(cv[m] * y, y, sh);
      }
      svh[r] = sh;
#pragma unroll
      for (int m = 0; m < KL; ++m) {
        const double y = (live && m < M) ? p.X[(long)gj * M + m] : 0.0;
        yyl[m * ST + r] = y;
        svl[m * ST + r] = m < M ? cv[m] * y * y : 0.0;
      }
    }
  };
  load_columns(0, colbuf);
#pragma unroll 1
  for (int tj = 0; tj < tj_end; ++tj) {
    __syncthreads();                             // buffer tj & 1 is complete (and the row data); buffer (tj + 1) & 1 is no longer read
    if (tj + 1 < tj_end) load_columns(tj + 1, colbuf + ((tj + 1) & 1) * col_stride);
    const double* buf = colbuf + (tj & 1) * col_stride;
    const double* yyl = buf;
    const double* svl = yyl + KL * ST;
    const double* svh = svl + KL * ST;
    const double* cc = svh + ST;
    const double* yyh = cc + ST;
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
      const int r = ty + 16 * u;
      const double wr = cr[r], shr = suh[r];
#pragma unroll 1
      for (int v = 0; v < 4; ++v) {
        const int cidx = tx + 16 * v;
        double e = shr + svh[cidx];
        for (int k = 0; k < nh; ++k) e = fma(gxh[k * ST + r], yyh[k * ST + cidx], e);
        const double H = wr * cc[cidx] * exp_tab(fmin(e, 708.0), etab);
        double h[KL];
#pragma unroll
        for (int m = 0; m < KL; ++m) h[m] = exp_tab(fma(gxl[m * ST + r], yyl[m * ST + cidx], sul[m * ST + r] + svl[m * ST + cidx]), etab);
        lattice_accumulate<KL>(acc, h, H);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NLO; ++s) {
    const double v = warp_sum(acc[s]);
    if (lane == 0) wpart[warp * NLO + s] = v;
  }
  __syncthreads();
  for (int s = tid; s < NLO; s += STHREADS) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += wpart[w * NLO + s];
    out[s] = v;
  }
}

struct LatticeDest {
  int dest[8][64];        // output row of V for (block of this launch chunk, low pattern), -1 = not requested
};

// V[dest][l][j] = sum_{a in l, b in j} sum_ti parts[pair][ti][block][lo]      grid (2^KL, L*L, blocks of the chunk)
__global__ void sobol_lattice_finish_kernel(const double* __restrict__ parts, int Lp, int L, int T, int nhi, int NLO, int hz0, LatticeDest map,
                                            double* __restrict__ V) {
  __shared__ double red[32];
  const int lo = blockIdx.x, lj = blockIdx.y, l = lj / L, j = lj - l * L, hz = hz0 + blockIdx.z;
  const int dest = map.dest[blockIdx.z][lo];
  if (dest < 0) return;
  double tot = 0.0;
  for (int ka = 0; ka < Lp; ++ka)
    for (int kb = 0; kb < Lp; ++kb) {
      const int a = l * Lp + ka, b = j * Lp + kb;
      const int hi_ = max(a, b), lo_ = min(a, b);
      const long pidx = (long)hi_ * (hi_ + 1) / 2 + lo_;
      const double* pp = parts + ((pidx * T) * nhi + hz) * NLO + lo;
      double acc = 0.0;
      for (int t = threadIdx.x; t < T; t += blockDim.x) acc += pp[(long)t * nhi * NLO];
      acc = block_sum(acc, red);
      tot += acc;   // meaningful in thread 0
    }
  if (threadIdx.x == 0) V[((long)dest * L + l) * L + j] = tot;
}

template <int KL>
static int launch_lattice(const SobolLatticeArgs& a, int npairs, cudaStream_t st) {
  const int hi_inputs = a.M - KL > 0 ? a.M - KL : 0;     // rows: gxl, sul [KL] + gxh [hi] + suh + cr; columns: two buffers of yyl, svl [KL] + yyh [hi] + svh + cc
  const size_t smem = (size_t)(4 * a.M + 3 * (2 * KL + hi_inputs + 2) * ST + 8 * (1 << KL)) * sizeof(double);
  RC_REQUIRE(smem <= 200 * 1024, -2, "sobol_contract: shared memory %zu too large for the lattice form", smem);
  if (smem > 48 * 1024) RC_ENSURE_SMEM(sobol_lattice_kernel<KL>, 200 * 1024);
  sobol_lattice_kernel<KL><<<dim3((a.T - a.part + a.nparts - 1) / a.nparts, npairs, a.nhi), STHREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

template <int MAXM, int RU>
static int launch_sweep(const SobolPairArgs& a, int npairs, cudaStream_t st) {
  const size_t smem = (size_t)(4 * a.M + 4 * a.M * ST + 2 * ST + 8 * 3 * MAXM + (size_t)a.M * 2 * RU * STHREADS) * sizeof(double);
  RC_ENSURE_SMEM((sobol_sweep_kernel<MAXM, RU>), 160 * 1024);
  RC_REQUIRE(smem <= 160 * 1024, -2, "sobol_contract: shared memory %zu too large", smem);
  sobol_sweep_kernel<MAXM, RU><<<dim3(a.T * ((a.T - a.part + a.nparts - 1) / a.nparts), npairs), STHREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

// Column tiles per CTA of the register form: long enough to amortise the row staging and the reduction, short enough to leave a few
// waves of CTAs per SM (RC_SOBOL_CHUNK overrides).
static int sweep_chunk(int T, int own_rows, int npairs) {
  static const int forced = [] { const char* e = getenv("RC_SOBOL_CHUNK"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced < T ? forced : T;
  const long slots = 3L * device_sm_count();
  int chunk = 8;
  while (chunk > 1 && (long)own_rows * ((T + chunk - 1) / chunk) * npairs < 5 * slots) chunk /= 2;   // tools/time_sweep_part.py: within 2 % of the best chunk for 1, 2, 4 and 8 ranks
  return chunk < T ? chunk : T;
}

template <int M, int RU>
static int launch_sweep_reg(SobolPairArgs& a, int npairs, cudaStream_t st) {
  const size_t smem = (size_t)(4 * ((M + 1) & ~1) + 4 * M * ST + 2 * ST + (SRTHREADS / 32) * 3 * M) * sizeof(double);
  static_assert((4 * 20 + 4 * 20 * ST + 2 * ST + 4 * 60) * sizeof(double) <= 48 * 1024, "fits the default dynamic shared memory limit");
  const int own_rows = (a.T - a.part + a.nparts - 1) / a.nparts;
  a.chunk = sweep_chunk(a.T, own_rows, npairs);
  sobol_sweep_reg_kernel<M, RU><<<dim3(own_rows * ((a.T + a.chunk - 1) / a.chunk), npairs), SRTHREADS, smem, st>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

// index of a structured subset in the sweep kernel's output, or -1 for a general subset
int sobol_sweep_index(unsigned long long mask, int M) {
  const unsigned long long full = (M >= 64) ? ~0ull : ((1ull << M) - 1ull);
  if (mask == 0) return 3 * M - 1;
  if ((mask & (mask - 1)) == 0) {            // single input
    int m = 0;
    while (!((mask >> m) & 1ull)) ++m;
    return m;
  }
  if (((mask + 1) & mask) == 0) {            // prefix [0:k], k >= 2
    int k = 0;
    while ((mask >> k) & 1ull) ++k;
    return M + k - 1;
  }
  const unsigned long long comp = full & ~mask;   // suffix [k:M]  <=>  complement is the prefix [0:k]
  if (comp != 0 && ((comp + 1) & comp) == 0) {
    int k = 0;
    while ((comp >> k) & 1ull) ++k;
    if (k >= 1 && k <= M - 1) return 2 * M + k - 1;
  }
  return -1;
}

struct SweepMap {
  int idx[SOBOL_MAX_SLICES];
};

// V[s0+s][l][j] = sum over (a in l, b in j) and tiles of parts[pair][tile][map.idx[s]]
__global__ void sobol_finish_map_kernel(const double* __restrict__ parts, int P, int Lp, int L, long tiles, int nv, SweepMap map,
                                        double* __restrict__ V, const int* __restrict__ dest) {
  __shared__ double red[32];
  const int s = blockIdx.x, lj = blockIdx.y, l = lj / L, j = lj - l * L;
  const int col = map.idx[s];
  double tot = 0.0;
  for (int ka = 0; ka < Lp; ++ka)
    for (int kb = 0; kb < Lp; ++kb) {
      const int a = l * Lp + ka, b = j * Lp + kb;
      const int hi = max(a, b), lo = min(a, b);
      const long pidx = (long)hi * (hi + 1) / 2 + lo;
      const double* pp = parts + pidx * tiles * nv + col;
      double acc = 0.0;
      for (long t = threadIdx.x; t < tiles; t += blockDim.x) acc += pp[t * nv];
      acc = block_sum(acc, red);
      tot += acc;   // meaningful in thread 0
    }
  (void)dest;
  if (threadIdx.x == 0) V[((long)s * L + l) * L + j] = tot;
}

// V[s][l][j] = sum_{a in l, b in j} sum_tiles parts[pair(max(a,b),min(a,b))][tile][s]
__global__ void sobol_finish_kernel(const double* __restrict__ parts, int P, int Lp, int L, long tiles, int ns, double* __restrict__ V) {
  __shared__ double red[32];
  const int s = blockIdx.x, lj = blockIdx.y, l = lj / L, j = lj - l * L;
  double tot = 0.0;
  for (int ka = 0; ka < Lp; ++ka)
    for (int kb = 0; kb < Lp; ++kb) {
      const int a = l * Lp + ka, b = j * Lp + kb;
      const int hi = max(a, b), lo = min(a, b);
      const long pidx = (long)hi * (hi + 1) / 2 + lo;
      const double* pp = parts + pidx * tiles * ns + s;
      double acc = 0.0;
      for (long t = threadIdx.x; t < tiles; t += blockDim.x) acc += pp[t * ns];
      acc = block_sum(acc, red);
      tot += acc;   // meaningful in thread 0
    }
  if (threadIdx.x == 0) V[((long)s * L + l) * L + j] = tot;
}

size_t sobol_workspace_bytes(int N, int P, int ns) {
  const long T = (N + ST - 1) / ST;
  (void)ns;   // 64 values per (pair, tile): a chunk of general subsets, or the 3M <= 36 outputs of the sweep form ...
  const long per_pair = T * T > T * SOBOL_MAX_SLICES ? T * T : T * SOBOL_MAX_SLICES;   // ... or 64 lattice blocks of 64 subsets per (pair, row tile)
  return (size_t)((long)P * (P + 1) / 2) * per_pair * SOBOL_MAX_SLICES * sizeof(double);
}

int sobol_contract(const double* X, int N, int M, const double* Phi, const double* c, int L, int Lp,
                   const unsigned long long* masks, int nslices, double* parts, double* V, int part, int nparts, cudaStream_t st) {
  RC_REQUIRE(M >= 1 && M <= 64, -2, "sobol_contract: M=%d out of range [1,64]", M);
  RC_REQUIRE(nparts >= 1 && part >= 0 && part < nparts, -2, "sobol_contract: part %d of %d", part, nparts);
  const int P = L * Lp;
  RC_ENSURE_SMEM(sobol_pair_kernel, 200 * 1024);
  const int T = (N + ST - 1) / ST;
  const int npairs = P * (P + 1) / 2;
  const int own_rows = (T - part + nparts - 1) / nparts;      // row tiles part, part + nparts, ... : the only ones this call launches
  if (own_rows <= 0) {                                         // more ranks than row tiles: this one has nothing to add
    RC_CUDA_OK(cudaMemsetAsync(V, 0, (size_t)nslices * L * L * sizeof(double), st));
    return 0;
  }
  // Blocks of the subset lattice (the 2^KL subsets that share their inputs >= KL) of which at least half is asked for - the all-subsets sweep -
  // take the lattice form: (KL + 1) exps and ~1.5 FP64 instructions per subset for the whole block.  RC_SOBOL_LATTICE=0 switches it off.
  std::vector<char> taken(nslices, 0);
  {
    static const bool lattice_on = [] { const char* e = getenv("RC_SOBOL_LATTICE"); return !e || atoi(e) != 0; }();
    const int KL = M < 6 ? M : 6, NLO = 1 << KL;
    if (lattice_on && KL >= 3 && M - KL <= 32) {
      std::map<unsigned long long, std::vector<int>> blocks;
      for (int s = 0; s < nslices; ++s) blocks[masks[s] >> KL].push_back(s);
      std::vector<unsigned> his;
      std::vector<std::vector<int>> dests;
      for (auto& kv : blocks) {
        // Only subsets the sweep form cannot serve count towards the threshold: a list of structured slices (everything gsa.models.GSA asks for) costs
        // M exps per pair there, less than a lattice block's KL + 1 exps and walk (M = 5, cfg3 sizes: 2.25 ms through the lattice, 0.9 through
        // the sweep).  Once a block qualifies, the structured subsets inside it ride along (the all-subsets sweep launches no sweep kernel).
        int unstructured = 0;
        for (int s : kv.second)
          if (M > 20 || sobol_sweep_index(masks[s], M) < 0) ++unstructured;
        if (unstructured * 2 < NLO) continue;
        std::vector<int> dest(64, -1);
        for (int s : kv.second) {
          const int lo = (int)(masks[s] & (unsigned long long)(NLO - 1));
          if (dest[lo] < 0) {             // a subset listed twice: the second copy goes the general way
            dest[lo] = s;
            taken[s] = 1;
          }
        }
        his.push_back((unsigned)kv.first);
        dests.push_back(dest);
      }
      for (size_t h0 = 0; h0 < his.size(); h0 += SOBOL_MAX_SLICES) {
        const int nhi = (int)std::min<size_t>(SOBOL_MAX_SLICES, his.size() - h0);
        SobolLatticeArgs a{};
        a.X = X; a.N = N; a.M = M; a.Phi = Phi; a.c = c; a.P = P; a.T = T; a.nhi = nhi; a.parts = parts; a.part = part; a.nparts = nparts;
        for (int z = 0; z < nhi; ++z) a.hi[z] = his[h0 + z];
        int rc = KL == 6 ? launch_lattice<6>(a, npairs, st) : KL == 5 ? launch_lattice<5>(a, npairs, st) : KL == 4 ? launch_lattice<4>(a, npairs, st)
                                                                                                                  : launch_lattice<3>(a, npairs, st);
        if (rc) return rc;
        for (int z0 = 0; z0 < nhi; z0 += 8) {
          const int nz = std::min(8, nhi - z0);
          LatticeDest map;
          for (int z = 0; z < 8; ++z)
            for (int lo = 0; lo < 64; ++lo) map.dest[z][lo] = z < nz ? dests[h0 + z0 + z][lo] : -1;
          sobol_lattice_finish_kernel<<<dim3(NLO, L * L, nz), 256, 0, st>>>(parts, Lp, L, own_rows, nhi, NLO, z0, map, V);
          RC_LAUNCH_OK();
        }
      }
    }
  }
  // Structured subsets (single inputs, prefixes, suffixes, full, empty: everything gsa.models.GSA asks for) go through the sweep form:
  // ONE launch whatever their number; only general subsets that are not part of a lattice block pay one exp per (pair, subset) below.
  std::vector<int> general;
  if (M <= 20) {
    std::vector<int> structured, sidx;
    for (int s = 0; s < nslices; ++s) {
      if (taken[s]) continue;
      const int k = sobol_sweep_index(masks[s], M);
      if (k >= 0) {
        structured.push_back(s);
        sidx.push_back(k);
      } else {
        general.push_back(s);
      }
    }
    if (!structured.empty()) {
      SobolPairArgs a{};
      a.X = X; a.N = N; a.M = M; a.Phi = Phi; a.c = c; a.P = P; a.T = T; a.ns = 3 * M; a.parts = parts; a.part = part; a.nparts = nparts;
      // RC_SOBOL_SWEEP=park selects the round-1 form (h parked in shared memory, polynomial exp) for every M; default: registers + table exp up to M = 12
      static const bool reg_form = [] { const char* e = getenv("RC_SOBOL_SWEEP"); return !e || e[0] != 'p'; }();
      int rc;
      long partials = (long)T * own_rows;        // per pair of output rows
      if (reg_form) {                   // M <= 20 here: three CTAs per SM up to M = 12, two (up to 255 registers) beyond
#define RC_SWEEP_CASE(MM, RR) case MM: rc = launch_sweep_reg<MM, RR>(a, npairs, st); break;
        switch (M) {
          RC_SWEEP_CASE(1, 2) RC_SWEEP_CASE(2, 2) RC_SWEEP_CASE(3, 2) RC_SWEEP_CASE(4, 2) RC_SWEEP_CASE(5, 2) RC_SWEEP_CASE(6, 2) RC_SWEEP_CASE(7, 2)
          RC_SWEEP_CASE(8, 2) RC_SWEEP_CASE(9, 1) RC_SWEEP_CASE(10, 1) RC_SWEEP_CASE(11, 1) RC_SWEEP_CASE(12, 1) RC_SWEEP_CASE(13, 1) RC_SWEEP_CASE(14, 1)
          RC_SWEEP_CASE(15, 1) RC_SWEEP_CASE(16, 1) RC_SWEEP_CASE(17, 1) RC_SWEEP_CASE(18, 1) RC_SWEEP_CASE(19, 1)
          default: rc = launch_sweep_reg<20, 1>(a, npairs, st); break;
        }
#undef RC_SWEEP_CASE
        partials = (long)own_rows * ((T + a.chunk - 1) / a.chunk);
      } else
        rc = M <= 4 ? launch_sweep<4, 2>(a, npairs, st) : M <= 8 ? launch_sweep<8, 2>(a, npairs, st)
             : M <= 12 ? launch_sweep<12, 2>(a, npairs, st) : launch_sweep<20, 1>(a, npairs, st);
      if (rc) return rc;
      // contiguous runs of output slots are finished together (<= SOBOL_MAX_SLICES per launch)
      size_t i = 0;
      while (i < structured.size()) {
        size_t j = i;
        SweepMap map{};
        while (j < structured.size() && j - i < (size_t)SOBOL_MAX_SLICES && structured[j] == structured[i] + (int)(j - i)) {
          map.idx[j - i] = sidx[j];
          ++j;
        }
        sobol_finish_map_kernel<<<dim3((unsigned)(j - i), L * L), 256, 0, st>>>(parts, P, Lp, L, partials, 3 * M, map,
                                                                                V + (long)structured[i] * L * L, nullptr);
        RC_LAUNCH_OK();
        i = j;
      }
    }
    if (general.empty()) return 0;
  } else {
    for (int s = 0; s < nslices; ++s)
      if (!taken[s]) general.push_back(s);
  }
  // general subsets: gather them into chunks (the output rows of a chunk need not be contiguous, so each chunk is finished per run)
  for (size_t g0 = 0; g0 < general.size();) {
    size_t g1 = g0;
    while (g1 < general.size() && g1 - g0 < (size_t)SOBOL_MAX_SLICES && general[g1] == general[g0] + (int)(g1 - g0)) ++g1;
    const int ns = (int)(g1 - g0), s0 = general[g0];
    SobolPairArgs a{};
    a.X = X; a.N = N; a.M = M; a.Phi = Phi; a.c = c; a.P = P; a.T = T; a.ns = ns; a.parts = parts; a.part = part; a.nparts = nparts;
    for (int s = 0; s < ns; ++s) a.masks[s] = masks[s0 + s];
    const size_t smem = (size_t)(4 * M + 3 * M * ST + 2 * ns * ST + 2 * ST + 8 * ns) * sizeof(double);
    RC_REQUIRE(smem <= 200 * 1024, -2, "sobol_contract: shared memory %zu too large", smem);
    sobol_pair_kernel<<<dim3(T * own_rows, npairs), STHREADS, smem, st>>>(a);
    RC_LAUNCH_OK();
    sobol_finish_kernel<<<dim3(ns, L * L), 256, 0, st>>>(parts, P, Lp, L, (long)T * own_rows, ns, V + (long)s0 * L * L);
    RC_LAUNCH_OK();
    g0 = g1;
  }
  return 0;
}

}  // namespace rc
