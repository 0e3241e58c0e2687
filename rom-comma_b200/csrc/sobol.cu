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
  const int ti = blockIdx.x / p.T, tj = blockIdx.x - ti * p.T;
  double* out = p.parts + ((long)pidx * p.T * p.T + blockIdx.x) * ns;
  if (a == b && tj > ti) {   // covered by the mirrored tile (weight 2)
    for (int s = threadIdx.x; s < ns; s += STHREADS) out[s] = 0.0;
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
      for (int v = 0; v < 4; ++v) rowacc = fma(ccc[v], exp(e[u][v]), rowacc);
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
  if (ns > SOBOL_MAX_SLICES) ns = SOBOL_MAX_SLICES;
  return (size_t)((long)P * (P + 1) / 2) * T * T * ns * sizeof(double);
}

int sobol_contract(const double* X, int N, int M, const double* Phi, const double* c, int L, int Lp,
                   const unsigned long long* masks, int nslices, double* parts, double* V, cudaStream_t st) {
  RC_REQUIRE(M >= 1 && M <= 64, -2, "sobol_contract: M=%d out of range [1,64]", M);
  const int P = L * Lp;
  static bool configured = false;
  if (!configured) {
    RC_CUDA_OK(cudaFuncSetAttribute(sobol_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const int T = (N + ST - 1) / ST;
  const int npairs = P * (P + 1) / 2;
  for (int s0 = 0; s0 < nslices; s0 += SOBOL_MAX_SLICES) {
    const int ns = min(SOBOL_MAX_SLICES, nslices - s0);
    SobolPairArgs a{};
    a.X = X; a.N = N; a.M = M; a.Phi = Phi; a.c = c; a.P = P; a.T = T; a.ns = ns; a.parts = parts;
    for (int s = 0; s < ns; ++s) a.masks[s] = masks[s0 + s];
    const size_t smem = (size_t)(4 * M + 3 * M * ST + 2 * ns * ST + 2 * ST + 8 * ns) * sizeof(double);
    RC_REQUIRE(smem <= 200 * 1024, -2, "sobol_contract: shared memory %zu too large", smem);
    sobol_pair_kernel<<<dim3(T * T, npairs), STHREADS, smem, st>>>(a);
    RC_LAUNCH_OK();
    sobol_finish_kernel<<<dim3(ns, L * L), 256, 0, st>>>(parts, P, Lp, L, (long)T * T, ns, V + (long)s0 * L * L);
    RC_LAUNCH_OK();
  }
  return 0;
}

}  // namespace rc
