// FP64 tensor-core (DMMA.8x8x4) tile GEMM used by the Cholesky, triangular-inverse, LAUUM and TRSM drivers.
//
//   C[m][n] = alpha * sum_{k in range(tile)} Aop[m][k] * Bop[n][k] + beta * C[m][n]
//   Aop[m][k] = TA ? A[k*lda + m] : A[m*lda + k]        (row-major storage everywhere)
//   Bop[n][k] = TB ? B[k*ldb + n] : B[n*ldb + k]
//
// CTA tile 128x128x16, 8 warps (2x4), warp tile 64x32 = 8x4 DMMA tiles, 4-stage cp.async pipeline (160 KB smem, 1 CTA/SM),
// persistent over the tile list (grid = number of SMs).
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
  int sel_block;    // > 0 (with lower_only): compute only tiles that intersect the diagonal blocks of size sel_block ("selected" LAUUM)
  int col_tiles;    // > 0 (with lower_only, K_FULL): only the first col_tiles tile columns of the lower triangle (look-ahead panel update)
  // Second (outer) batch level, for launches that are already batched over sub-problems of ONE matrix (the pairs of a trtri level) and have to
  // cover several matrices as well: problem z = z2 * batch + z1 reads A + z1*strideA + z2*strideA2 (B, C likewise).  0 / 1 = a single level.
  int batch2;
  long strideA2, strideB2, strideC2;
  int batch1;       // set by the launcher: the inner batch count (what `batch` of launch_gemm_ws is)
};

// number of 128 x 128 tiles of one matrix in the list the arguments describe
inline long gemm_tile_count(const GemmArgs& a) {
  const long tm = a.M / 128, tn = a.N / 128;
  if (!a.lower_only) return tm * tn;
  if (a.col_tiles > 0 && a.col_tiles < tm) return (long)a.col_tiles * (a.col_tiles + 1) / 2 + (tm - a.col_tiles) * a.col_tiles;
  return tm * (tm + 1) / 2;
}

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

constexpr int G_RASTER = 12;   // super-tile edge of the L2-blocked tile order (12 x 12 = 144 tiles ~ one wave of 148 SMs)

struct GemmTile {
  int m0, n0, kb, nk;   // nk == 0: nothing to do for this tile (skipped by sel_block)
  const double* A; const double* B; double* C;
};

// tile index (over all matrices of the batch) -> coordinates, K range and base pointers
// LOWER: -1 = decide at run time from p.lower_only (the cp.async kernel), 0 / 1 = known at compile time (the warp-specialised kernel is
// instantiated per case so that each copy carries only its own index arithmetic).
__host__ __device__ __forceinline__ int gemm_imin(int a, int b) { return a < b ? a : b; }

template <int LOWER = -1>
__host__ __device__ __forceinline__ GemmTile gemm_decode_tile(const GemmArgs& p, long tile, long tiles_per_matrix) {
  const bool lower = LOWER < 0 ? (p.lower_only != 0) : (LOWER != 0);
  const int z = (int)(tile / tiles_per_matrix);
  const long idx = tile - (long)z * tiles_per_matrix;
  int tm, tn;
  if (lower && p.col_tiles > 0 && p.col_tiles < p.M / G_BM) {
    // the first col_tiles tile columns of the lower triangle, row-major: the triangle on the diagonal first, then full rows of col_tiles tiles
    // (a strip of tiles shares one row panel of A and the same col_tiles panels of B, which stay in L2)
    const int w = p.col_tiles, tri = w * (w + 1) / 2, i32 = (int)idx;
    if (i32 < tri) {
      tm = (int)((sqrt(8.0 * (double)i32 + 1.0) - 1.0) * 0.5);
      while ((tm + 1) * (tm + 2) / 2 <= i32) ++tm;
      while (tm * (tm + 1) / 2 > i32) --tm;
      tn = i32 - tm * (tm + 1) / 2;
    } else {
      const int r = i32 - tri;
      tm = w + r / w;
      tn = r - (tm - w) * w;
    }
  } else if (lower && p.kmode != K_FULL) {
    // row-major over the lower triangle: tile row tm has the longest K of what is left (K_GE_M0: LAUUM, where the raster below costs 2 % -
    // the selected tile lists leave it too few tiles per super-tile - and saves no traffic)
    tm = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((long)(tm + 1) * (tm + 2) / 2 <= idx) ++tm;
    while ((long)tm * (tm + 1) / 2 > idx) --tm;
    tn = (int)(idx - (long)tm * (tm + 1) / 2);
  } else if (lower) {
    // Lower triangle in the same L2-blocked raster (rank-k updates: first rank-512 trailing update 3.2 -> 2.4 GB of DRAM traffic, 2.1 GB algorithmic): super-rows of G_RASTER tile rows (top first: longest K for K_GE_M0), inside one
    // the full G_RASTER-wide super-tiles left to right, then the triangular one on the diagonal; rows first inside a super-tile.
    // C(s) = tiles above super-row s = 72 s^2 + 6 s for G_RASTER = 12 (every super-row above the last one is full).
    constexpr int R = G_RASTER, FULL = R * R, DIAG = R * (R + 1) / 2;
    const int tiles_m = p.M / G_BM, i32 = (int)idx;
    int si = (int)((sqrt((double)DIAG * DIAG + 2.0 * FULL * (double)i32) - DIAG) / FULL);     // root of FULL s^2/2 + (DIAG - FULL/2) s = idx ...
    auto above = [&](int s_) { return FULL * (s_ * (s_ - 1) / 2) + DIAG * s_; };
    while (above(si + 1) <= i32) ++si;
    while (above(si) > i32) --si;
    const int h = gemm_imin(R, tiles_m - si * R);
    int r = i32 - above(si);
    int tml, tnl, sj;
    if (r < si * h * R) {
      sj = r / (h * R);
      r -= sj * h * R;
      tml = r / R;
      tnl = r - tml * R;
    } else {
      sj = si;
      r -= si * h * R;
      tml = (int)((sqrt(8.0 * (double)r + 1.0) - 1.0) * 0.5);
      while ((tml + 1) * (tml + 2) / 2 <= r) ++tml;
      while (tml * (tml + 1) / 2 > r) --tml;
      tnl = r - tml * (tml + 1) / 2;
    }
    tm = si * R + tml;
    tn = sj * R + tnl;
  } else {
    // L2-blocked raster: the ~148 tiles in flight at any time form a G_RASTER x G_RASTER super-tile, so they share G_RASTER row panels
    // of A and G_RASTER column panels of B and walk k almost in step - each operand panel comes from DRAM about once per super-tile
    // instead of once per tile (column-major order re-read the 128 x K row panels for every tile column: 18 GB for the 8192^3
    // triangular product of trtri's top level against 1.3 GB algorithmic).  Super-tiles are ordered along the dimension the K range
    // depends on, longest K first, and so are the tiles inside one: the dynamic scheduler's load balance is unchanged.
    const int tiles_m = p.M / G_BM, tiles_n = p.N / G_BN;
    const bool k_by_row = p.kmode == K_LT_M1 || p.kmode == K_GE_M0;
    const int slow_n = k_by_row ? tiles_m : tiles_n, fast_n = k_by_row ? tiles_n : tiles_m;    // slow: the dimension K depends on
    const int idx32 = (int)idx;                                        // tiles of ONE matrix: < 2^31, 32-bit divisions
    const int per_slow_band = G_RASTER * fast_n;
    const int sb = idx32 / per_slow_band;                              // band along the slow dimension (only the last one is narrower)
    const int slow_w = gemm_imin(G_RASTER, slow_n - sb * G_RASTER);
    const int r1 = idx32 - sb * per_slow_band;
    const int fb = r1 / (G_RASTER * slow_w);                           // band along the fast dimension
    const int fast_w = gemm_imin(G_RASTER, fast_n - fb * G_RASTER);
    const int r2 = r1 - fb * G_RASTER * slow_w;
    int slow = sb * G_RASTER + r2 / fast_w, fast = fb * G_RASTER + r2 % fast_w;
    if (p.kmode == K_LT_M1 || p.kmode == K_LE_N1) slow = slow_n - 1 - slow;   // K grows with the index: start from the far end
    tm = k_by_row ? slow : fast;
    tn = k_by_row ? fast : slow;
  }
  GemmTile t;
  t.m0 = tm * G_BM;
  t.n0 = tn * G_BN;
  int kb = 0, ke = p.K;
  if (p.kmode == K_GE_N0) kb = t.n0;
  else if (p.kmode == K_LT_M1) ke = gemm_imin(p.K, t.m0 + G_BM);
  else if (p.kmode == K_GE_M0) kb = t.m0;
  else if (p.kmode == K_LE_N1) ke = gemm_imin(p.K, t.n0 + G_BN);
  t.kb = kb;
  t.nk = (ke - kb) / G_BK;
  if (p.sel_block > 0 && t.m0 / p.sel_block > (t.n0 + G_BN - 1) / p.sel_block) t.nk = -1;   // row blocks all above the column blocks
  if (p.batch2 > 1) {
    const int z2 = z / p.batch1, z1 = z - z2 * p.batch1;
    t.A = p.A + (long)z1 * p.strideA + (long)z2 * p.strideA2;
    t.B = p.B + (long)z1 * p.strideB + (long)z2 * p.strideB2;
    t.C = p.C + (long)z1 * p.strideC + (long)z2 * p.strideC2;
  } else {
    t.A = p.A + (long)z * p.strideA;
    t.B = p.B + (long)z * p.strideB;
    t.C = p.C + (long)z * p.strideC;
  }
  return t;
}

// Persistent kernel: one CTA per SM walks the tile list with stride gridDim.x.  The cp.async pipeline runs ACROSS tiles: the first
// stages of the next tile are requested before the epilogue of the current one, and the C tile is prefetched into L2 while the
// main loop runs, so the read-modify-write epilogue overlaps the next tile's loads instead of draining the SM.
template <bool TA, bool TB>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_dmma_kernel(GemmArgs p, long tiles_per_matrix, long total_tiles) {
  extern __shared__ __align__(16) double smem[];
  using S = GemmSmem<TA, TB>;
  double* As = smem;
  double* Bs = smem + G_STAGES * S::A_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
  const int g = lane >> 2, t = lane & 3;

  auto prologue = [&](const GemmTile& T) {
#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
      if (s < T.nk) {
        gemm_load_operand<TA>(As + s * S::A_STAGE, T.A, p.lda, T.m0, T.kb + s * G_BK, tid);
        gemm_load_operand<TB>(Bs + s * S::B_STAGE, T.B, p.ldb, T.n0, T.kb + s * G_BK, tid);
      }
      cp_async_commit();
    }
  };

  long tile = blockIdx.x;
  GemmTile T;
  auto next_tile = [&]() -> bool {        // advance `tile` to the next one that has work; false when the list is exhausted
    for (; tile < total_tiles; tile += gridDim.x) {
      T = gemm_decode_tile(p, tile, tiles_per_matrix);
      if (T.nk >= 0) return true;
    }
    return false;
  };
  if (!next_tile()) return;
  prologue(T);

  for (;;) {
    if (p.beta != 0.0) {   // pull the C tile (128 rows x 1 KB) into L2 while the main loop runs
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int line = tid + r * G_THREADS;           // 1024 lines of 128 B
        const double* addr = T.C + (long)(T.m0 + (line >> 3)) * p.ldc + T.n0 + (line & 7) * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
      }
    }
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < T.nk; ++kt) {
      cp_async_wait<G_STAGES - 2>();
      __syncthreads();
      {
        const int nxt = kt + G_STAGES - 1;
        if (nxt < T.nk) {
          const int s = nxt % G_STAGES;
          gemm_load_operand<TA>(As + s * S::A_STAGE, T.A, p.lda, T.m0, T.kb + nxt * G_BK, tid);
          gemm_load_operand<TB>(Bs + s * S::B_STAGE, T.B, p.ldb, T.n0, T.kb + nxt * G_BK, tid);
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
    // All operand reads of this tile are complete before any of its stores: tile-exclusive in-place updates (C aliasing A or B) are safe.
    cp_async_wait<0>();
    __syncthreads();

    const GemmTile done = T;
    tile += gridDim.x;
    const bool more = next_tile();
    if (more) prologue(T);                           // overlaps the epilogue below

    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long row = done.m0 + wm * 64 + i * 8 + g;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = done.n0 + wn * 32 + j * 8 + t * 2;
        double2* cp = reinterpret_cast<double2*>(done.C + row * p.ldc + col);
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
    if (!more) break;
  }
}

// flops the tile list of one launch executes (2 * 128 * 128 * K-range per tile): host mirror of gemm_decode_tile, profiling only
inline double gemm_tile_flops(const GemmArgs& p, int batch) {
  const long tm_n = p.M / G_BM, tn_n = p.N / G_BN;
  double ksum = 0.0;
  for (long tm = 0; tm < tm_n; ++tm)
    for (long tn = 0; tn < (p.lower_only ? (p.col_tiles > 0 && p.col_tiles < tm + 1 ? p.col_tiles : tm + 1) : tn_n); ++tn) {
      const long m0 = tm * G_BM, n0 = tn * G_BN;
      long kb = 0, ke = p.K;
      if (p.kmode == K_GE_N0) kb = n0;
      else if (p.kmode == K_LT_M1) ke = (p.K < m0 + G_BM) ? p.K : m0 + G_BM;
      else if (p.kmode == K_GE_M0) kb = m0;
      else if (p.kmode == K_LE_N1) ke = (p.K < n0 + G_BN) ? p.K : n0 + G_BN;
      if (p.sel_block > 0 && m0 / p.sel_block > (n0 + G_BN - 1) / p.sel_block) continue;
      if (ke > kb) ksum += (double)(ke - kb);
    }
  return 2.0 * G_BM * G_BN * ksum * batch * (p.batch2 > 1 ? p.batch2 : 1);
}

template <bool TA, bool TB>
inline int launch_gemm(const GemmArgs& a, int batch, cudaStream_t stream) {
  using S = GemmSmem<TA, TB>;
  RC_ENSURE_SMEM((gemm_dmma_kernel<TA, TB>), S::BYTES);
  const int num_sms = device_sm_count();
  if (a.M <= 0 || a.N <= 0 || batch <= 0) return 0;
  RC_REQUIRE(a.M % G_BM == 0 && a.N % G_BN == 0 && a.K % G_BK == 0, -2, "gemm_dmma: M,N must be multiples of 128 and K of 16 (got %d,%d,%d)", a.M, a.N, a.K);
  const long tiles = gemm_tile_count(a);
  const long total = tiles * batch;
  const unsigned grid = (unsigned)(total < num_sms ? total : num_sms);   // persistent: one CTA per SM (160 KB smem each)
  const bool prof = profile_enabled();
  if (prof) profile_gemm_begin(stream);
  gemm_dmma_kernel<TA, TB><<<grid, G_THREADS, S::BYTES, stream>>>(a, tiles, total);
  if (prof) profile_gemm_end(stream, gemm_tile_flops(a, batch));
  RC_LAUNCH_OK();
  return 0;
}

}  // namespace rc
