// FP64 tensor-core (DMMA.8x8x4) tile GEMM used by the Cholesky, triangular-inverse, LAUUM and TRSM drivers.
//
//   C[m][n] = alpha * sum_{k in range(tile)} Aop[m][k] * Bop[n][k] + beta * C[m][n]
//   Aop[m][k] = TA ? A[k*lda + m] : A[m*lda + k]        (row-major storage everywhere)
//   Bop[n][k] = TB ? B[k*ldb + n] : B[n*ldb + k]
//
// CTA tile 128x128x16, 8 warps (2x4), warp tile 64x32 = 8x4 DMMA tiles, 4-stage cp.async pipeline (160 KB smem, 1 CTA/SM).
// Shared-memory rows are padded by 4 doubles so that every half-warp fragment read (4 rows x 4 consecutive doubles, or its
// transpose) hits 16 distinct 8-byte banks.  M, N are multiples of 128 and K of 16 (callers pad matrices with identity).
// Triangular structure is exploited at tile granularity through `kmode` (per-tile K range) and `lower_only` (tile list);
// inside diagonal 128-blocks the operands carry explicit zeros.
#pragma once
#include "common.cuh"

namespace rc {

enum KMode : int {
  K_FULL = 0,
  K_GE_N0 = 1,   // k >= n0           (B lower-triangular, stored [k][n])
  K_LT_M1 = 2,   // k <  m0 + 128     (A lower-triangular, stored [m][k])
  K_GE_M0 = 3,   // k >= m0           (A lower-triangular, stored [k][m])
  K_LE_N1 = 4    // k <  n0 + 128     (B lower-triangular, stored [n][k])
};

struct GemmArgs {
  const double* A; long lda; long strideA;
  const double* B; long ldb; long strideB;
  double* C;       long ldc; long strideC;
  int M, N, K;
  double alpha, beta;
  int lower_only;   // 1: only tiles with tile_m >= tile_n (requires M == N)
  int kmode;
};

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_STAGES = 4, G_PAD = 4, G_THREADS = 256;

template <bool TA, bool TB>
struct GemmSmem {
  static constexpr int A_STAGE = TA ? G_BK * (G_BM + G_PAD) : G_BM * (G_BK + G_PAD);
  static constexpr int B_STAGE = TB ? G_BK * (G_BN + G_PAD) : G_BN * (G_BK + G_PAD);
  static constexpr size_t BYTES = (size_t)G_STAGES * (A_STAGE + B_STAGE) * sizeof(double);
};

template <bool T>
__device__ __forceinline__ void gemm_load_operand(double* __restrict__ dst, const double* __restrict__ src, long ld, int mn0, int k0, int tid) {
  // 128 x 16 doubles = 1024 16-byte chunks; 4 per thread.
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = tid + r * G_THREADS;
    if (!T) {
      const int row = c >> 3, kc = (c & 7) * 2;
      cp_async16(dst + row * (G_BK + G_PAD) + kc, src + (long)(mn0 + row) * ld + k0 + kc);
    } else {
      const int krow = c >> 6, mc = (c & 63) * 2;
      cp_async16(dst + krow * (G_BM + G_PAD) + mc, src + (long)(k0 + krow) * ld + mn0 + mc);
    }
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_dmma_kernel(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  using S = GemmSmem<TA, TB>;
  double* As = smem;
  double* Bs = smem + G_STAGES * S::A_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
  const int g = lane >> 2, t = lane & 3;

  int tm, tn;
  if (p.lower_only) {
    const int idx = blockIdx.x;
    tm = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((long)(tm + 1) * (tm + 2) / 2 <= idx) ++tm;
    while ((long)tm * (tm + 1) / 2 > idx) --tm;
    tn = idx - tm * (tm + 1) / 2;
  } else {
    const int tiles_m = p.M / G_BM;
    tm = blockIdx.x % tiles_m;
    tn = blockIdx.x / tiles_m;
    if (p.kmode == K_LT_M1) tm = tiles_m - 1 - tm;   // longest K ranges first
    if (p.kmode == K_LE_N1) tn = p.N / G_BN - 1 - tn;
  }
  const int m0 = tm * G_BM, n0 = tn * G_BN;
  int kb = 0, ke = p.K;
  if (p.kmode == K_GE_N0) kb = n0;
  else if (p.kmode == K_LT_M1) ke = min(p.K, m0 + G_BM);
  else if (p.kmode == K_GE_M0) kb = m0;
  else if (p.kmode == K_LE_N1) ke = min(p.K, n0 + G_BN);
  const int nk = (ke - kb) / G_BK;

  const double* A = p.A + (long)blockIdx.z * p.strideA;
  const double* B = p.B + (long)blockIdx.z * p.strideB;
  double* C = p.C + (long)blockIdx.z * p.strideC;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < nk) {
      gemm_load_operand<TA>(As + s * S::A_STAGE, A, p.lda, m0, kb + s * G_BK, tid);
      gemm_load_operand<TB>(Bs + s * S::B_STAGE, B, p.ldb, n0, kb + s * G_BK, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + G_STAGES - 1;
      if (nxt < nk) {
        const int s = nxt % G_STAGES;
        gemm_load_operand<TA>(As + s * S::A_STAGE, A, p.lda, m0, kb + nxt * G_BK, tid);
        gemm_load_operand<TB>(Bs + s * S::B_STAGE, B, p.ldb, n0, kb + nxt * G_BK, tid);
      }
      cp_async_commit();
    }
    const double* as = As + (kt % G_STAGES) * S::A_STAGE;
    const double* bs = Bs + (kt % G_STAGES) * S::B_STAGE;
#pragma unroll
    for (int kk = 0; kk < G_BK / 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        a[i] = TA ? as[(kk * 4 + t) * (G_BM + G_PAD) + wm * 64 + i * 8 + g] : as[(wm * 64 + i * 8 + g) * (G_BK + G_PAD) + kk * 4 + t];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = TB ? bs[(kk * 4 + t) * (G_BN + G_PAD) + wn * 32 + j * 8 + g] : bs[(wn * 32 + j * 8 + g) * (G_BK + G_PAD) + kk * 4 + t];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  // All operand reads are complete before any store: makes tile-exclusive in-place updates (C aliasing A or B) safe.
  cp_async_wait<0>();
  __syncthreads();

  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long row = m0 + wm * 64 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + wn * 32 + j * 8 + t * 2;
      double2* cp = reinterpret_cast<double2*>(C + row * p.ldc + col);
      double2 v;
      if (beta == 0.0) {
        v.x = alpha * acc[i][j][0];
        v.y = alpha * acc[i][j][1];
      } else {
        const double2 o = *cp;
        v.x = fma(alpha, acc[i][j][0], beta * o.x);
        v.y = fma(alpha, acc[i][j][1], beta * o.y);
      }
      *cp = v;
    }
  }
}

template <bool TA, bool TB>
inline int launch_gemm(const GemmArgs& a, int batch, cudaStream_t stream) {
  using S = GemmSmem<TA, TB>;
  static bool configured = false;   // per instantiation; benign race (idempotent attribute set)
  if (!configured) {
    RC_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
    configured = true;
  }
  if (a.M <= 0 || a.N <= 0 || batch <= 0) return 0;
  RC_REQUIRE(a.M % G_BM == 0 && a.N % G_BN == 0 && a.K % G_BK == 0, -2, "gemm_dmma: M,N must be multiples of 128 and K of 16 (got %d,%d,%d)", a.M, a.N, a.K);
  const long tm = a.M / G_BM, tn = a.N / G_BN;
  const long tiles = a.lower_only ? tm * (tm + 1) / 2 : tm * tn;
  dim3 grid((unsigned)tiles, 1, (unsigned)batch);
  gemm_dmma_kernel<TA, TB><<<grid, G_THREADS, S::BYTES, stream>>>(a);
  RC_LAUNCH_OK();
  return 0;
}

}  // namespace rc
