// Blocked FP64 Cholesky, triangular solves, triangular inverse and LAUUM on row-major lower-triangular storage.
// Replaces the tf.linalg.cholesky / triangular_solve / cholesky_solve call sites of the reference
// (romcomma/gpf/models.py:81-82, romcomma/gpr/models.py:439,444, gpflow base_conditional) and provides the explicit
// inverse the analytic LML gradient needs.  All O(n^3) work runs through gemm_dmma_kernel.
//
// Conventions: n is a multiple of 128 (callers pad with identity), `ld` is the row stride in doubles, `batch` independent
// matrices are `strideA` doubles apart.  `dinv` holds one 128x128 inverse per diagonal block (upper triangle zero):
// block b of batch z lives at dinv + z*strideD + b*128*128.
#include "chol.h"
#include "gemm_dmma_ws.cuh"

#include <algorithm>
#include <vector>
#include <utility>
#include <cstdlib>
#include <atomic>
#include <mutex>

namespace rc {

// Scheduler scratch of the warp-specialised GEMM (see gemm_dmma_ws.cuh): per device, allocated and zeroed on first use.
int* gemm_sched_slot(int device) {
  constexpr int MAX_DEV = 64;
  static int* base[MAX_DEV] = {nullptr};
  static std::atomic<unsigned> seq[MAX_DEV];
  static std::mutex mu;
  if (device < 0 || device >= MAX_DEV) return nullptr;
  if (!base[device]) {
    std::lock_guard<std::mutex> lock(mu);
    if (!base[device]) {
      int* p = nullptr;
      if (cudaMalloc(&p, (size_t)SCHED_SLOTS * 2 * sizeof(int)) != cudaSuccess) return nullptr;
      if (cudaMemset(p, 0, (size_t)SCHED_SLOTS * 2 * sizeof(int)) != cudaSuccess) return nullptr;   // synchronous: ordered before any launch
      base[device] = p;
    }
  }
  const unsigned s = seq[device].fetch_add(1u, std::memory_order_relaxed) % SCHED_SLOTS;
  return base[device] + 2 * s;
}

// ----------------------------------------------------------------------------------------------------------------
// Diagonal block: L = chol(A_kk) in place (lower part only is read/written), Dinv = L^-1, partial log-determinant.
// ----------------------------------------------------------------------------------------------------------------
constexpr int DB = 128;
constexpr int DB_LD = DB + 1;
constexpr int DB_THREADS = 256;
constexpr size_t DB_SMEM = (size_t)(DB * DB_LD + 5 * DB + 32) * sizeof(double);

// One CTA per diagonal block, 16 x 16 threads; thread (ty,tx) owns the elements (ty + 16a, tx + 16b), a >= b, of the lower triangle in
// REGISTERS.  Phase 1 is a right-looking rank-1 Cholesky with ONE barrier per column: every thread tracks the running diagonal of its
// rows and columns redundantly, so the owners of column j+1 can scale and publish it (double-buffered `colbuf`) in the same step that
// applies column j.  Phase 2 builds L^-1 the same way (apply E_j^-1 to the identity, one barrier per row, L read from shared memory).
// The pivot loop is unrolled over the eight 16-column blocks (template parameter JB) so that every register index is static and the
// blocks that are already finished are skipped at compile time.
struct DiagCtx {
  double* colbuf; double* rowbuf; double* rinvs; const double* S;
  int ty, tx; int* info; int info_base;
};

// publish column jn = 16*NB + (jn & 15):  colbuf[jn&1][r] = L[r][jn] (r > jn), 0 for the other rows of blocks >= NB
template <int NB>
__device__ __forceinline__ void diag_produce_col(double (&a)[8][8], const double (&dc)[8], const DiagCtx& c, int jn) {
  if (c.tx != (jn & 15)) return;
  const double d = dc[NB];
  const double rinv = rsqrt(d);
  double* cb = c.colbuf + (jn & 1) * DB;
#pragma unroll
  for (int ia = NB; ia < 8; ++ia) {
    const int r = c.ty + 16 * ia;
    double v = 0.0;
    if (r > jn) {
      v = a[ia][NB] * rinv;
      a[ia][NB] = v;
    } else if (ia == NB && r == jn) {
      a[ia][NB] = d * rinv;       // L_jj = sqrt(d)
      c.rinvs[jn] = rinv;
      if (!(d > 0.0)) atomicCAS(c.info, 0, c.info_base + jn + 1);
    }
    cb[r] = v;
  }
}

template <int JB>
__device__ __forceinline__ void diag_potrf_block(double (&a)[8][8], double (&dr)[8], double (&dc)[8], const DiagCtx& c) {
  for (int jj = 0; jj < 16; ++jj) {
    const int j = 16 * JB + jj;
    __syncthreads();
    const double* cb = c.colbuf + (j & 1) * DB;
    double lr[8], lc[8];
#pragma unroll
    for (int i = JB; i < 8; ++i) {
      lr[i] = cb[c.ty + 16 * i];
      lc[i] = cb[c.tx + 16 * i];
      dr[i] = fma(-lr[i], lr[i], dr[i]);
      dc[i] = fma(-lc[i], lc[i], dc[i]);
    }
    // the two column blocks that can hold the next pivot first, then publish it, then the rest
#pragma unroll
    for (int ib = JB; ib < 8 && ib <= JB + 1; ++ib)
#pragma unroll
      for (int ia = ib; ia < 8; ++ia) a[ia][ib] = fma(-lr[ia], lc[ib], a[ia][ib]);
    if (jj < 15) diag_produce_col<JB>(a, dc, c, j + 1);
    else if (JB < 7) diag_produce_col<(JB < 7 ? JB + 1 : 7)>(a, dc, c, j + 1);
#pragma unroll
    for (int ib = JB + 2; ib < 8; ++ib)
#pragma unroll
      for (int ia = ib; ia < 8; ++ia) a[ia][ib] = fma(-lr[ia], lc[ib], a[ia][ib]);
  }
}

// publish row jn of Z:  rowbuf[jn&1][col] = Z[jn][col] / L_jj for col <= jn, 0 for the other columns of blocks <= AN
template <int AN>
__device__ __forceinline__ void diag_produce_row(double (&zz)[8][8], const DiagCtx& c, int jn) {
  if (c.ty != (jn & 15)) return;
  const double rinv = c.rinvs[jn];
  double* rb = c.rowbuf + (jn & 1) * DB;
#pragma unroll
  for (int ib = 0; ib <= AN; ++ib) {
    const int col = c.tx + 16 * ib;
    double v = 0.0;
    if (col <= jn) {
      v = zz[AN][ib] * rinv;
      zz[AN][ib] = v;
    }
    rb[col] = v;
  }
}

template <int JB>
__device__ __forceinline__ void diag_trtri_block(double (&zz)[8][8], const DiagCtx& c) {
  for (int jj = 0; jj < 16; ++jj) {
    const int j = 16 * JB + jj;
    __syncthreads();
    const double* rb = c.rowbuf + (j & 1) * DB;
    double lr[8], zr[8];
#pragma unroll
    for (int i = JB; i < 8; ++i) {
      const int r = c.ty + 16 * i;
      lr[i] = (i > JB || c.ty > jj) ? c.S[r * DB_LD + j] : 0.0;
    }
#pragma unroll
    for (int i = 0; i <= JB; ++i) zr[i] = rb[c.tx + 16 * i];
#pragma unroll
    for (int ia = JB; ia < 8 && ia <= JB + 1; ++ia)
#pragma unroll
      for (int ib = 0; ib <= JB; ++ib) zz[ia][ib] = fma(-lr[ia], zr[ib], zz[ia][ib]);
    if (jj < 15) diag_produce_row<JB>(zz, c, j + 1);
    else if (JB < 7) diag_produce_row<(JB < 7 ? JB + 1 : 7)>(zz, c, j + 1);
#pragma unroll
    for (int ia = JB + 2; ia < 8; ++ia)
#pragma unroll
      for (int ib = 0; ib <= JB; ++ib) zz[ia][ib] = fma(-lr[ia], zr[ib], zz[ia][ib]);
  }
}

// MODE 0: factor block `blk` and invert it (one launch per block).  MODE 1: factor only - what the critical path of potrf_lower needs
// (the panel below is solved against L directly).  MODE 2: invert only, all blocks at once (grid = (nblk, batch), blk = blockIdx.x):
// the inverses feed the later solves / trtri and are computed off the critical path, in one wave.
template <int MODE>
__global__ void __launch_bounds__(DB_THREADS, 1)
diag_potrf_inv_kernel(double* __restrict__ A, long ld, long strideA, double* __restrict__ dinv, long strideD, int blk_arg,
                      double* __restrict__ logdet_parts, int nblk, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                       // [128][129]  L, for phase 2
  double* red = sm + DB * DB_LD + 5 * DB;
  const int tid = threadIdx.x;
  const int z = MODE == 2 ? blockIdx.y : blockIdx.x;
  const int blk = MODE == 2 ? blockIdx.x : blk_arg;
  DiagCtx c;
  c.colbuf = sm + DB * DB_LD;           // [2][128]
  c.rowbuf = c.colbuf + 2 * DB;         // [2][128]
  c.rinvs = c.rowbuf + 2 * DB;          // [128]  1 / L_jj
  c.S = S;
  c.ty = tid >> 4;
  c.tx = tid & 15;
  c.info = info + z;
  c.info_base = blk * DB;
  const int ty = c.ty, tx = c.tx;
  double* Ab = A + (long)z * strideA + (long)blk * DB * (ld + 1);
  double* Db = dinv + (long)z * strideD + (long)blk * DB * DB;

  pdl_launch_dependents();
  pdl_wait();
  double a[8][8], dr[8], dc[8];
#pragma unroll
  for (int ia = 0; ia < 8; ++ia) {
    const int r = ty + 16 * ia;
    dr[ia] = Ab[(long)r * ld + r];
    dc[ia] = Ab[(long)(tx + 16 * ia) * ld + tx + 16 * ia];
#pragma unroll
    for (int ib = 0; ib < 8; ++ib) {
      const int col = tx + 16 * ib;
      a[ia][ib] = (ib <= ia && col <= r) ? Ab[(long)r * ld + col] : 0.0;
    }
  }
  if (MODE != 2) {
    diag_produce_col<0>(a, dc, c, 0);
    diag_potrf_block<0>(a, dr, dc, c);
    diag_potrf_block<1>(a, dr, dc, c);
    diag_potrf_block<2>(a, dr, dc, c);
    diag_potrf_block<3>(a, dr, dc, c);
    diag_potrf_block<4>(a, dr, dc, c);
    diag_potrf_block<5>(a, dr, dc, c);
    diag_potrf_block<6>(a, dr, dc, c);
    diag_potrf_block<7>(a, dr, dc, c);
  }

  // L -> global (lower part) and shared memory; log-determinant part
  double logsum = 0.0;
#pragma unroll
  for (int ia = 0; ia < 8; ++ia) {
    const int r = ty + 16 * ia;
#pragma unroll
    for (int ib = 0; ib <= ia; ++ib) {
      const int col = tx + 16 * ib;
      if (col <= r) {
        if (MODE != 2) Ab[(long)r * ld + col] = a[ia][ib];
        if (MODE != 1) S[r * DB_LD + col] = a[ia][ib];
        if (col == r) {
          logsum += log(a[ia][ib]);
          if (MODE == 2) c.rinvs[r] = 1.0 / a[ia][ib];      // a holds the finished factor read back from A
        }
      }
    }
  }
  if (MODE != 2) {
    logsum = block_sum(logsum, red);
    if (tid == 0) logdet_parts[(long)z * nblk + blk] = logsum;
  }
  if (MODE == 1) return;
  __syncthreads();                               // publishes S and rinvs

  // Phase 2: Z = L^-1.  Start from the identity and apply E_j^-1 for j = 0..127: row j <- row j / L_jj, rows i > j -= L[i][j] * row j.
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int ib = 0; ib < 8; ++ib) a[ia][ib] = (ty + 16 * ia == tx + 16 * ib) ? 1.0 : 0.0;
  diag_produce_row<0>(a, c, 0);
  diag_trtri_block<0>(a, c);
  diag_trtri_block<1>(a, c);
  diag_trtri_block<2>(a, c);
  diag_trtri_block<3>(a, c);
  diag_trtri_block<4>(a, c);
  diag_trtri_block<5>(a, c);
  diag_trtri_block<6>(a, c);
  diag_trtri_block<7>(a, c);
#pragma unroll
  for (int ia = 0; ia < 8; ++ia) {
    const int r = ty + 16 * ia;
#pragma unroll
    for (int ib = 0; ib < 8; ++ib) {
      const int col = tx + 16 * ib;
      Db[r * DB + col] = (ib <= ia && col <= r) ? a[ia][ib] : 0.0;
    }
  }
}

static int launch_diag_factor(double* A, long ld, long strideA, double* dinv, long strideD, int blk, double* logdet_parts, int nblk, int* info,
                              int batch, cudaStream_t st) {
  RC_ENSURE_SMEM(diag_potrf_inv_kernel<1>, DB_SMEM);
  RC_CUDA_OK(launch_pdl(diag_potrf_inv_kernel<1>, dim3(batch), dim3(DB_THREADS), DB_SMEM, st, A, ld, strideA, dinv, strideD, blk, logdet_parts, nblk, info));
  RC_LAUNCH_OK();
  return 0;
}

static int launch_diag_invert_all(double* A, long ld, long strideA, double* dinv, long strideD, int nblk, int batch, cudaStream_t st) {
  RC_ENSURE_SMEM(diag_potrf_inv_kernel<2>, DB_SMEM);
  diag_potrf_inv_kernel<2><<<dim3(nblk, batch), DB_THREADS, DB_SMEM, st>>>(A, ld, strideA, dinv, strideD, 0, nullptr, nblk, nullptr);
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Panel solve  P <- P * L_kk^-T  for the 128-row tiles below a freshly factored diagonal block (the TRSM of the right-looking step),
// against L itself: the inverse of the block is not on the critical path any more.
// One CTA per 128 x 128 tile; warp w owns rows [16w, 16w+16) and keeps them as DMMA accumulators (2 x 16 tiles of 8 x 8).  The 128
// columns are processed in four blocks of 32:  (1) the block's accumulators go to a per-warp staging area, (2) two lanes per row solve the
// 16 x 32 system against the 32 x 32 diagonal sub-block by substitution in registers (column j finished, broadcast by a shuffle, eliminated
// from the later columns; all indices static), (3) the solved block X eliminates itself from the later column blocks on the tensor cores:
// acc[:, c] -= X * L[c, block]^T with A fragments from the staging area and B fragments from the staged factor.  No CTA barrier after staging.
// ----------------------------------------------------------------------------------------------------------------
constexpr int PT_LP = DB + 4;                       // pitch of the staged factor (conflict-free half-warp fragment reads)
constexpr int PT_XP = 32 + 4;                       // pitch of a warp's 16 x 32 staging block
constexpr size_t PT_SMEM = (size_t)(DB * PT_LP + 8 * 16 * PT_XP + DB) * sizeof(double);

__global__ void __launch_bounds__(256, 1)
panel_trsm_kernel(double* __restrict__ A, long ld, long strideA, int blk) {
  extern __shared__ __align__(16) double sm[];
  double* Ls = sm;                                   // [128][PT_LP]  L[k][j], lower part valid
  double* Xs_all = sm + DB * PT_LP;                  // [8][16][PT_XP]
  double* rinv = Xs_all + 8 * 16 * PT_XP;            // [128]  1 / L[j][j]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, z = blockIdx.y;
  const int g = lane >> 2, t = lane & 3;
  double* Az = A + (long)z * strideA;
  const double* Lb = Az + (long)blk * DB * (ld + 1);
  double* Pt = Az + ((long)(blk + 1 + blockIdx.x) * DB + warp * 16) * ld + (long)blk * DB;    // this warp's 16 rows
  pdl_launch_dependents();
  pdl_wait();
  // accumulators: acc[h][c][e] = P[8h + g][8c + 2t + e]
  double acc[2][16][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const double2 v = *reinterpret_cast<const double2*>(Pt + (long)(8 * h + g) * ld + 8 * c + 2 * t);
      acc[h][c][0] = v.x;
      acc[h][c][1] = v.y;
    }
  // Stage the factor with cp.async (16 B per request, coalesced, all 32 requests of a thread in flight at once): as a load -> store loop
  // the 32 iterations were 32 serial L2 round trips - half of this kernel's 40 us (ncu source view: long-scoreboard + barrier samples).
  for (int idx = tid; idx < DB * (DB / 2); idx += 256) {
    const int k = idx >> 6, j2 = (idx & 63) * 2;
    if (j2 <= k) cp_async16(Ls + k * PT_LP + j2, Lb + (long)k * ld + j2);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (tid < DB) rinv[tid] = 1.0 / Ls[tid * PT_LP + tid];
  __syncthreads();
  double* Xs = Xs_all + warp * 16 * PT_XP;
  const int rr = lane >> 1, par = lane & 1;                        // substitution: row rr of the warp, columns of parity par
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    // (1) accumulators of this column block -> staging (C layout: row 8h+g, columns 8c' + 2t + e)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        Xs[(8 * h + g) * PT_XP + 8 * c + 2 * t] = acc[h][4 * cb + c][0];
        Xs[(8 * h + g) * PT_XP + 8 * c + 2 * t + 1] = acc[h][4 * cb + c][1];
      }
    __syncwarp();
    // (2) 16 x 32 substitution against the diagonal 32 x 32 sub-block
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = Xs[rr * PT_XP + 2 * i + par];
    const double* Ld = Ls + (32 * cb) * PT_LP + 32 * cb;          // L[32cb + k][32cb + j] = Ld[k * PT_LP + j]
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int i0 = j >> 1;
      const double mine = x[i0] * rinv[32 * cb + j];
      const double xj = __shfl_sync(0xffffffffu, mine, (lane & ~1) | (j & 1));
      if (par == (j & 1)) x[i0] = xj;
      if ((j & 1) == 0 && par == 1) x[i0] = fma(-xj, Ld[(j + 1) * PT_LP + j], x[i0]);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i > i0) x[i] = fma(-xj, Ld[(2 * i + par) * PT_LP + j], x[i]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      Xs[rr * PT_XP + 2 * i + par] = x[i];                         // solved block: A operand of the update ...
      Pt[(long)rr * ld + 32 * cb + 2 * i + par] = x[i];            // ... and the result
    }
    __syncwarp();
    // (3) eliminate the solved block from the later column blocks: acc[:, c] -= X * L[c, block]^T
    if (cb < 3) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const double a0 = -Xs[g * PT_XP + 4 * ks + t], a1 = -Xs[(8 + g) * PT_XP + 4 * ks + t];
#pragma unroll
        for (int c = 4 * (cb + 1); c < 16; ++c) {
          const double bv = Ls[(8 * c + g) * PT_LP + 32 * cb + 4 * ks + t];
          dmma884(acc[0][c][0], acc[0][c][1], a0, bv);
          dmma884(acc[1][c][0], acc[1][c][1], a1, bv);
        }
      }
      __syncwarp();                                                // staging area is rewritten by step (1) of the next block
    }
  }
}

static int launch_panel_trsm(double* A, int n, long ld, long strideA, int blk, int batch, cudaStream_t st) {
  const int tiles = n / DB - blk - 1;
  if (tiles <= 0) return 0;
  RC_ENSURE_SMEM(panel_trsm_kernel, PT_SMEM);
  RC_CUDA_OK(launch_pdl(panel_trsm_kernel, dim3(tiles, batch), dim3(256), PT_SMEM, st, A, ld, strideA, blk));
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// potrf
// ----------------------------------------------------------------------------------------------------------------
size_t potrf_workspace_bytes(int n, int batch) {
  const size_t nblk = (size_t)n / DB;
  return (size_t)batch * nblk * DB * DB * sizeof(double)   // dinv
         + (size_t)batch * nblk * sizeof(double);          // logdet parts
}

// The factorisation proper.  `after_panel(done)` (optional) is called once the panel solve of block `done - 1` has been enqueued, i.e. when
// block columns [0, done) of L are final for ALL rows in stream order - the hook the overlapped inverse below forks its side work from.
// `invert_all`: finish with the batched launch of the 128 x 128 inverses (off the critical path).
// Block columns factored between two trailing updates (rank 128 * group): 16 while 96+ blocks remain (RC_POTRF_T16), 8 while 48+ remain
// (RC_POTRF_T8), 4 while 40+ remain (RC_POTRF_T4), else 2; RC_POTRF_GROUP (1..16) fixes the width.  Wide groups amortise the read-modify-write
// of the trailing matrix (K = 1024 / 2048 tiles run at 97 % of the DMMA pipe); their longer in-group chain is hidden by the look-ahead below and
// cut into wide launches by factor_columns.  Round 2 at n = 16384 / 24576 / 32768: 50.3 / 152.4 / 346.9 ms (round 1, widths 4/2: 52.9 / 158.2 / -).
static int potrf_group_env() {
  const char* e = getenv("RC_POTRF_GROUP");
  const int v = e ? atoi(e) : 0;
  return v < 1 ? 0 : (v > 16 ? 16 : v);
}
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int potrf_group(int blocks_left) {
  const int v = potrf_group_env();
  if (v) return v;
  static const int t4 = env_int("RC_POTRF_T4", 40), t8 = env_int("RC_POTRF_T8", 48), t16 = env_int("RC_POTRF_T16", 96);
  return blocks_left >= t16 ? 16 : (blocks_left >= t8 ? 8 : (blocks_left >= t4 ? 4 : 2));
}

// Width of the group that starts at block column b0.  RC_POTRF_FIRST > 0 opens a large factorisation with a narrow group (nothing can overlap the
// first group's chain); measured at n = 16384: 50.4 ms without, 50.7-51.2 ms with a first group of 1, 2 or 4 - off by default.
static int potrf_group_at(int b0, int nblk) {
  static const int first = env_int("RC_POTRF_FIRST", 0), first_min = env_int("RC_POTRF_LA_MIN_BLOCKS", 32);
  int w = potrf_group(nblk - b0);
  if (b0 == 0 && nblk >= first_min && first >= 1 && w > first && potrf_group_env() == 0) w = first;
  return std::min(w, nblk - b0);
}

// Factor block columns [c0, c0 + w) of a group, all rows below included, RECURSIVELY: left half, then ONE update of the right half's columns
// with the left half as K (a lower-trapezoidal tile list: the first w - h tile columns of the square that starts at block c0 + h), then the
// right half.  The same flops as updating every column on its own with everything to its left in the group (K = 128 j, N = 128 - seven
// tall-skinny launches per 8-wide group that ran at half the tile kernel's rate: 6.6 ms of a 16384 factorisation), in wider launches:
// per 8-wide group one K = 512 x 4 columns, two K = 256 x 2 columns and four K = 128 x 1 column.  RC_POTRF_FLAT=1 restores the flat form.
static int factor_columns(double* A, int n, long ld, long strideA, int batch, double* dinv, long strideD, double* logdet_parts, int nblk, int* info,
                          int c0, int w, cudaStream_t st) {
  int rc;
  if (w == 1) {
    if ((rc = launch_diag_factor(A, ld, strideA, dinv, strideD, c0, logdet_parts, nblk, info, batch, st))) return rc;
    return launch_panel_trsm(A, n, ld, strideA, c0, batch, st);
  }
  int h = 1;
  while (2 * h < w) h *= 2;
  if ((rc = factor_columns(A, n, ld, strideA, batch, dinv, strideD, logdet_parts, nblk, info, c0, h, st))) return rc;
  const long r0 = (long)(c0 + h) * DB;
  GemmArgs u{};
  u.A = A + r0 * ld + (long)c0 * DB; u.lda = ld; u.strideA = strideA;
  u.B = u.A; u.ldb = ld; u.strideB = strideA;
  u.C = A + r0 * ld + r0; u.ldc = ld; u.strideC = strideA;
  u.M = u.N = n - (int)r0; u.K = h * DB; u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1; u.kmode = K_FULL;
  u.col_tiles = w - h;
  if ((rc = launch_gemm_ws<false, false>(u, batch, st))) return rc;
  return factor_columns(A, n, ld, strideA, batch, dinv, strideD, logdet_parts, nblk, info, c0 + h, w - h, st);
}

static int factor_group_flat(double* A, int n, long ld, long strideA, int batch, double* dinv, long strideD, double* logdet_parts, int nblk, int* info,
                             int b0, int w, cudaStream_t st) {
  int rc;
  for (int j = 0; j < w; ++j) {
    const long r0 = (long)(b0 + j) * DB;
    if (j > 0) {   // block column b0+j, rows from block b0+j down:  A -= P[:, b0:b0+j] * P[b0+j, b0:b0+j]^T
      GemmArgs g{};
      g.A = A + r0 * ld + (long)b0 * DB; g.lda = ld; g.strideA = strideA;
      g.B = g.A; g.ldb = ld; g.strideB = strideA;
      g.C = A + r0 * ld + r0; g.ldc = ld; g.strideC = strideA;
      g.M = n - (int)r0; g.N = DB; g.K = j * DB; g.alpha = -1.0; g.beta = 1.0; g.lower_only = 0; g.kmode = K_FULL;
      if ((rc = launch_gemm_ws<false, false>(g, batch, st))) return rc;
    }
    if ((rc = launch_diag_factor(A, ld, strideA, dinv, strideD, b0 + j, logdet_parts, nblk, info, batch, st))) return rc;
    if ((rc = launch_panel_trsm(A, n, ld, strideA, b0 + j, batch, st))) return rc;
  }
  return 0;
}

static int factor_group(double* A, int n, long ld, long strideA, int batch, double* dinv, long strideD, double* logdet_parts, int nblk, int* info,
                        int b0, int w, cudaStream_t st) {
  static const bool flat = env_int("RC_POTRF_FLAT", 0) != 0;
  return flat ? factor_group_flat(A, n, ld, strideA, batch, dinv, strideD, logdet_parts, nblk, info, b0, w, st)
              : factor_columns(A, n, ld, strideA, batch, dinv, strideD, logdet_parts, nblk, info, b0, w, st);
}

template <typename Hook>
static int potrf_core(double* A, int n, long ld, long strideA, int batch, double* dinv, double* logdet_parts, int* info, cudaStream_t st,
                      Hook&& after_panel, bool invert_all) {
  RC_REQUIRE(n > 0 && n % DB == 0, -2, "potrf_lower: n=%d must be a positive multiple of 128", n);
  RC_REQUIRE(ld >= n && ld % 2 == 0, -2, "potrf_lower: ld=%ld must be even and >= n", ld);
  const int nblk = n / DB;
  const long strideD = (long)nblk * DB * DB;
  RC_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
  int rc;
  for (int b0 = 0, w = 0; b0 < nblk; b0 += w) {
    w = potrf_group_at(b0, nblk);
    if ((rc = factor_group(A, n, ld, strideA, batch, dinv, strideD, logdet_parts, nblk, info, b0, w, st))) return rc;
    if ((rc = after_panel(b0 + w))) return rc;
    const long r0 = (long)(b0 + w) * DB;
    if (r0 < n) {   // trailing update, lower tiles only, rank 128*w
      GemmArgs g{};
      g.A = A + r0 * ld + (long)b0 * DB; g.lda = ld; g.strideA = strideA;
      g.B = g.A; g.ldb = ld; g.strideB = strideA;
      g.C = A + r0 * ld + r0; g.ldc = ld; g.strideC = strideA;
      g.M = g.N = n - (int)r0; g.K = w * DB; g.alpha = -1.0; g.beta = 1.0; g.lower_only = 1; g.kmode = K_FULL;
      if ((rc = launch_gemm_ws<false, false>(g, batch, st))) return rc;
    }
  }
  // the 128 x 128 inverses the solves / trtri need: all blocks in one wave, off the critical path
  if (invert_all && (rc = launch_diag_invert_all(A, ld, strideA, dinv, strideD, nblk, batch, st))) return rc;
  return 0;
}


// ----------------------------------------------------------------------------------------------------------------
// Vector triangular solves (one right-hand side per matrix), one launch per block step, no atomics.
// ----------------------------------------------------------------------------------------------------------------
// forward:  x = L^-1 y.   Step k: x_k = Dinv_k * w_k ;  w_i -= L[i,k] x_k  for block rows i > k.   (w is a scratch copy of y)
__global__ void __launch_bounds__(256) trsv_fwd_step_kernel(const double* __restrict__ A, long ld, long strideA, const double* __restrict__ dinv,
                                                            long strideD, double* __restrict__ w, double* __restrict__ x, long strideV, int k) {
  __shared__ double wk[DB], xk[DB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, z = blockIdx.y;
  const double* Az = A + (long)z * strideA;
  const double* Dk = dinv + (long)z * strideD + (long)k * DB * DB;
  double* wz = w + (long)z * strideV;
  pdl_launch_dependents();
  pdl_wait();
  if (tid < DB) wk[tid] = wz[(long)k * DB + tid];
  __syncthreads();
  for (int r = warp; r < DB; r += 8) {     // x_k[r] = sum_c Dinv[r][c] w_k[c]
    double s = 0.0;
    for (int c = lane; c <= (r | 31); c += 32) s = fma(Dk[r * DB + c], wk[c], s);
    s = warp_sum(s);
    if (lane == 0) xk[r] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < DB) x[(long)z * strideV + (long)k * DB + tid] = xk[tid];
    return;
  }
  const long i0 = (long)(k + blockIdx.x) * DB;
  const double* Lik = Az + i0 * ld + (long)k * DB;
  for (int r = warp; r < DB; r += 8) {
    double s = 0.0;
#pragma unroll
    for (int c = lane; c < DB; c += 32) s = fma(Lik[(long)r * ld + c], xk[c], s);
    s = warp_sum(s);
    if (lane == 0) wz[i0 + r] -= s;
  }
}

// backward:  x = L^-T y.  Step k (descending): x_k = Dinv_k^T w_k ;  w_j -= L[k,j]^T x_k  for block columns j < k.
__global__ void __launch_bounds__(256) trsv_bwd_step_kernel(const double* __restrict__ A, long ld, long strideA, const double* __restrict__ dinv,
                                                            long strideD, double* __restrict__ w, double* __restrict__ x, long strideV, int k) {
  __shared__ double wk[DB], xk[DB], part[2 * DB];
  const int tid = threadIdx.x, z = blockIdx.y;
  const int c = tid & (DB - 1), half = tid >> 7;
  const double* Az = A + (long)z * strideA;
  const double* Dk = dinv + (long)z * strideD + (long)k * DB * DB;
  double* wz = w + (long)z * strideV;
  pdl_launch_dependents();
  pdl_wait();
  if (tid < DB) wk[tid] = wz[(long)k * DB + tid];
  __syncthreads();
  {
    double s = 0.0;
    for (int r = half * 64; r < half * 64 + 64; ++r) s = fma(Dk[r * DB + c], wk[r], s);   // Dinv upper part is zero
    part[half * DB + c] = s;
  }
  __syncthreads();
  if (tid < DB) xk[tid] = part[tid] + part[DB + tid];
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < DB) x[(long)z * strideV + (long)k * DB + tid] = xk[tid];
    return;
  }
  const long j0 = (long)(blockIdx.x - 1) * DB;
  const double* Lkj = Az + (long)k * DB * ld + j0;
  double s = 0.0;
  for (int r = half * 64; r < half * 64 + 64; ++r) s = fma(Lkj[(long)r * ld + c], xk[r], s);
  part[half * DB + c] = s;
  __syncthreads();
  if (tid < DB) wz[j0 + tid] -= part[tid] + part[DB + tid];
}

int trsv_lower(const double* A, int n, long ld, long strideA, int batch, const double* dinv, double* w, double* x, long strideV, int transpose,
               cudaStream_t st) {
  RC_REQUIRE(n > 0 && n % DB == 0, -2, "trsv_lower: n=%d must be a positive multiple of 128", n);
  const int nblk = n / DB;
  const long strideD = (long)nblk * DB * DB;
  if (!transpose) {
    for (int k = 0; k < nblk; ++k) {
      dim3 grid(nblk - k, batch);
      RC_CUDA_OK(launch_pdl(trsv_fwd_step_kernel, grid, dim3(256), 0, st, A, ld, strideA, dinv, strideD, w, x, strideV, k));
    }
    count_launches(nblk - 1);
  } else {
    for (int k = nblk - 1; k >= 0; --k) {
      dim3 grid(k + 1, batch);
      RC_CUDA_OK(launch_pdl(trsv_bwd_step_kernel, grid, dim3(256), 0, st, A, ld, strideA, dinv, strideD, w, x, strideV, k));
    }
    count_launches(nblk - 1);
  }
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Multi right-hand-side forward solve  B <- L^-1 B  (B is n x nrhs row-major, nrhs a multiple of 128), GEMM based.
// ----------------------------------------------------------------------------------------------------------------
// Two levels (wide right-hand sides): inside a super-block of TRSM_SB 128-blocks the classic right-looking substitution (B_k <- Dinv_k B_k, then a K = 128 update
// of the few block rows left in the super-block - latency-sized launches), then ONE update of everything below the super-block with
// K = 128 * TRSM_SB, long enough for the tile kernel to run near its peak (K = 128 updates of the whole remainder ran at ~60 % of it).
// Measured at n = 16384 (tools/time_predict.py, profiles/r01_trsm_predict.jsonl): 512 right-hand sides 10.2 ms with super-block 1 vs 11.1 with 8
// (the launches are latency-sized either way); 2048: 24.7 -> 22.4 ms; 8192: 81.2 -> 70.6 ms (27.1 -> 31.2 TFLOP/s).
static int trsm_super_block(int nrhs) {
  const char* e = getenv("RC_TRSM_SB");
  const int v = e ? atoi(e) : (nrhs <= 1024 ? 1 : 8);
  return v < 1 ? 1 : v;
}
int trsm_lower_fwd(const double* A, int n, long ld, long strideA, int batch, const double* dinv, double* B, int nrhs, long ldb, long strideB,
                   cudaStream_t st) {
  RC_REQUIRE(n % DB == 0 && nrhs % DB == 0, -2, "trsm_lower_fwd: n=%d and nrhs=%d must be multiples of 128", n, nrhs);
  const int nblk = n / DB, TRSM_SB = trsm_super_block(nrhs);
  const long strideD = (long)nblk * DB * DB;
  int rc;
  auto update = [&](int r0, int r1, int k0, int k1) -> int {   // B[r0:r1] -= L[r0:r1, k0:k1] * B[k0:k1]   (block indices)
    GemmArgs u{};
    u.A = A + (long)r0 * DB * ld + (long)k0 * DB; u.lda = ld; u.strideA = strideA;
    u.B = B + (long)k0 * DB * ldb; u.ldb = ldb; u.strideB = strideB;
    u.C = B + (long)r0 * DB * ldb; u.ldc = ldb; u.strideC = strideB;
    u.M = (r1 - r0) * DB; u.N = nrhs; u.K = (k1 - k0) * DB; u.alpha = -1.0; u.beta = 1.0; u.kmode = K_FULL;
    return launch_gemm_ws<false, true>(u, batch, st);
  };
  for (int k0 = 0; k0 < nblk; k0 += TRSM_SB) {
    const int k1 = std::min(nblk, k0 + TRSM_SB);
    for (int k = k0; k < k1; ++k) {
      GemmArgs g{};   // B_k <- Dinv_k * B_k   (tile-exclusive in place: a CTA owns its 128 columns over all 128 k-rows)
      g.A = dinv + (long)k * DB * DB; g.lda = DB; g.strideA = strideD;
      g.B = B + (long)k * DB * ldb; g.ldb = ldb; g.strideB = strideB;
      g.C = B + (long)k * DB * ldb; g.ldc = ldb; g.strideC = strideB;
      g.M = DB; g.N = nrhs; g.K = DB; g.alpha = 1.0; g.beta = 0.0; g.kmode = K_FULL;
      if ((rc = launch_gemm_ws<false, true>(g, batch, st))) return rc;
      if (k + 1 < k1 && (rc = update(k + 1, k1, k, k + 1))) return rc;
    }
    if (k1 < nblk && (rc = update(k1, nblk, k0, k1))) return rc;
  }
  return 0;
}

// Few right-hand sides against a large factor (Sobol error columns, predictions: nrhs <= 1024 at n = 16384).  The block substitution above is
// then 2 n/128 latency-sized launches of at most nrhs/128 (diagonal block) and (n/128 - k) nrhs/128 (rank-128 update) tiles: 10 ms for
// 1.4e11 flop.  Here the diagonal SUPER-blocks of TRSM_SBI rows are inverted once per factor (trsm_sbinv_prepare: a batched trtri_lower on a
// copy, 3 levels), so that a super-block costs one triangular product T = Z_sb B_sb (out of place, copied back) and ONE rank-TRSM_SBI update
// of everything below - n / TRSM_SBI steps with K = 1024 updates instead of n / 128 with K = 128.  Rows beyond the last full super-block take
// the block substitution.  (Measured and dropped: a look-ahead as in potrf_lookahead - the triangular products and the next super-block's rows
// as a chain on the high-priority stream underneath yielding updates, same bits - gained 0.27 ms at 512 columns (7.21 -> 6.95 ms) and lost
// 0.4 ms at 2048: the 32-tile chain kernels still need their SMs and the yielding updates run 5 % slower.)
// work: trsm_sbinv_workspace_doubles(n, nrhs_max) doubles = [Z: nsb x SBI x SBI][trtri tmp: nsb x SBI^2/4][T: SBI x nrhs].
constexpr int TRSM_SBI = 1024;
size_t trsm_sbinv_workspace_doubles(int n, int nrhs) {
  const size_t nsb = (size_t)(n / TRSM_SBI);
  return nsb * TRSM_SBI * TRSM_SBI + nsb * (TRSM_SBI * (size_t)TRSM_SBI / 4) + (size_t)TRSM_SBI * nrhs;
}
int trsm_sbinv_prepare(const double* A, int n, long ld, const double* dinv, double* work, cudaStream_t st) {
  const int nsb = n / TRSM_SBI;
  if (nsb == 0) return 0;
  double* Z = work;
  double* tmp = Z + (size_t)nsb * TRSM_SBI * TRSM_SBI;
  for (int z = 0; z < nsb; ++z)
    RC_CUDA_OK(cudaMemcpy2DAsync(Z + (size_t)z * TRSM_SBI * TRSM_SBI, TRSM_SBI * sizeof(double), A + (long)z * TRSM_SBI * (ld + 1), ld * sizeof(double),
                                 TRSM_SBI * sizeof(double), TRSM_SBI, cudaMemcpyDeviceToDevice, st));
  return trtri_lower(Z, TRSM_SBI, TRSM_SBI, (long)TRSM_SBI * TRSM_SBI, nsb, dinv, tmp, (long)TRSM_SBI * TRSM_SBI / 4, st, 0, TRSM_SBI / DB, 0);
}
int trsm_lower_fwd_sbinv(const double* A, int n, long ld, const double* dinv, const double* work, double* Tbuf, double* B, int nrhs, long ldb,
                         cudaStream_t st) {
  RC_REQUIRE(n % DB == 0 && nrhs % DB == 0, -2, "trsm_lower_fwd_sbinv: n=%d and nrhs=%d must be multiples of 128", n, nrhs);
  const int nsb = n / TRSM_SBI, nblk = n / DB, SBB = TRSM_SBI / DB;
  const double* Z = work;
  int rc;
  for (int z = 0; z < nsb; ++z) {
    double* Bz = B + (long)z * TRSM_SBI * ldb;
    GemmArgs g{};   // T = Z_z * B_z   (Z_z lower, stored [m][k]  ->  k < m0 + 128)
    g.A = Z + (size_t)z * TRSM_SBI * TRSM_SBI; g.lda = TRSM_SBI;
    g.B = Bz; g.ldb = ldb;
    g.C = Tbuf; g.ldc = nrhs;
    g.M = TRSM_SBI; g.N = nrhs; g.K = TRSM_SBI; g.alpha = 1.0; g.beta = 0.0; g.kmode = K_LT_M1;
    if ((rc = launch_gemm_ws<false, true>(g, 1, st))) return rc;
    RC_CUDA_OK(cudaMemcpy2DAsync(Bz, ldb * sizeof(double), Tbuf, nrhs * sizeof(double), nrhs * sizeof(double), TRSM_SBI, cudaMemcpyDeviceToDevice, st));
    const int r0 = (z + 1) * SBB;
    if (r0 < nblk) {   // B[below] -= L[below, super-block z] * B_z
      GemmArgs u{};
      u.A = A + (long)r0 * DB * ld + (long)z * TRSM_SBI; u.lda = ld;
      u.B = Bz; u.ldb = ldb;
      u.C = B + (long)r0 * DB * ldb; u.ldc = ldb;
      u.M = (nblk - r0) * DB; u.N = nrhs; u.K = TRSM_SBI; u.alpha = -1.0; u.beta = 1.0; u.kmode = K_FULL;
      if ((rc = launch_gemm_ws<false, true>(u, 1, st))) return rc;
    }
  }
  const int k_tail = nsb * SBB;
  if (k_tail < nblk)   // the ragged rest: block substitution on the trailing (n - k_tail*128) rows, already updated by everything above
    return trsm_lower_fwd(A + (long)k_tail * DB * (ld + 1), (nblk - k_tail) * DB, ld, 0, 1, dinv + (long)k_tail * DB * DB, B + (long)k_tail * DB * ldb, nrhs,
                          ldb, 0, st);
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Triangular inverse (in place) and K^-1 = L^-T L^-1 (out of place, lower triangle).
// ----------------------------------------------------------------------------------------------------------------
__global__ void copy_dinv_to_diag_kernel(double* __restrict__ A, long ld, long strideA, const double* __restrict__ dinv, long strideD, int first_block) {
  const int blk = first_block + blockIdx.x, z = blockIdx.y;
  double* Ab = A + (long)z * strideA + (long)blk * DB * (ld + 1);
  const double* Db = dinv + (long)z * strideD + (long)blk * DB * DB;
  for (int idx = threadIdx.x; idx < DB * DB; idx += blockDim.x) Ab[(long)(idx >> 7) * ld + (idx & 127)] = Db[idx];
}

// Bottom-up recursive doubling: with Z11, Z22 the inverses of two adjacent diagonal blocks of size h,
//   Z21 = -Z22 * (L21 * Z11),  all pairs of a level batched into two launches.  `tmp` needs >= n*n/4 doubles per matrix.
// rest_from > 0 (a power of two times 128, < n): the leading rest_from x rest_from block has been inverted already (by a call with n = rest_from,
// possibly while the factorisation of the trailing columns was still running - potrf_trtri_lower); do everything else: the pairs of every level
// that lie beyond it and the level(s) that couple it to the rest.  Every pair is computed exactly as in the one-call form.  strideD_blocks: blocks
// per matrix in dinv (0 = n / 128).  tiles_per_cta > 0: yielding launches (background work on a low-priority stream).
int trtri_lower(double* A, int n, long ld, long strideA, int batch, const double* dinv, double* tmp, long strideT, cudaStream_t st, int rest_from,
                int strideD_blocks, int tiles_per_cta) {
  RC_REQUIRE(n > 0 && n % DB == 0, -2, "trtri_lower: n=%d must be a positive multiple of 128", n);
  RC_REQUIRE(rest_from >= 0 && rest_from < n && rest_from % DB == 0 && (rest_from & (rest_from - 1)) == 0, -2,
             "trtri_lower: rest_from=%d must be 0 or a power of two times 128 below n=%d", rest_from, n);
  const int nblk = n / DB;
  const long strideD = (long)(strideD_blocks > 0 ? strideD_blocks : nblk) * DB * DB;
  copy_dinv_to_diag_kernel<<<dim3(nblk - rest_from / DB, batch), 256, 0, st>>>(A, ld, strideA, dinv, strideD, rest_from / DB);
  RC_LAUNCH_OK();
  int rc;
  for (long h = DB; h < n; h *= 2) {
    const int np = (int)(n / (2 * h));
    const long rem = n - (long)np * 2 * h;
    const int first_pair = (rest_from > 0 && 2 * h <= rest_from) ? (int)(rest_from / (2 * h)) : 0;     // pairs inside the finished leading block
    // (count, h2): full problems then the ragged one (second block shorter)
    for (int pass = 0; pass < 2; ++pass) {
      const int count = pass == 0 ? np - first_pair : ((rem > h) ? 1 : 0);
      if (count <= 0) continue;
      const long h2 = pass == 0 ? h : rem - h;
      const long base = pass == 0 ? (long)first_pair * 2 * h : (long)np * 2 * h;
      {   // the pairs of a level are the inner batch of one launch, the matrices its outer batch (looping over the matrices cost 120 latency-sized
          // launches = 7.5 of the 11 ms of a batched evaluation of ten n = 1920 folds)
        double* Az = A + base * (ld + 1);
        double* Tz = tmp;
        GemmArgs g{};   // T = L21 * Z11     (Z11 lower, stored [k][n]  ->  k >= n0)
        g.A = Az + h * ld; g.lda = ld; g.strideA = 2 * h * (ld + 1);
        g.B = Az; g.ldb = ld; g.strideB = 2 * h * (ld + 1);
        g.C = Tz; g.ldc = h; g.strideC = h * h;
        g.M = (int)h2; g.N = (int)h; g.K = (int)h; g.alpha = 1.0; g.beta = 0.0; g.kmode = K_GE_N0;
        g.batch2 = batch; g.strideA2 = strideA; g.strideB2 = strideA; g.strideC2 = strideT;
        if ((rc = launch_gemm_ws<false, true>(g, count, st, tiles_per_cta))) return rc;
        GemmArgs u{};   // Z21 = -Z22 * T    (Z22 lower, stored [m][k]  ->  k < m0 + 128)
        u.A = Az + h * ld + h; u.lda = ld; u.strideA = 2 * h * (ld + 1);
        u.B = Tz; u.ldb = h; u.strideB = h * h;
        u.C = Az + h * ld; u.ldc = ld; u.strideC = 2 * h * (ld + 1);
        u.M = (int)h2; u.N = (int)h; u.K = (int)h2; u.alpha = -1.0; u.beta = 0.0; u.kmode = K_LT_M1;
        u.batch2 = batch; u.strideA2 = strideA; u.strideB2 = strideT; u.strideC2 = strideA;
        if ((rc = launch_gemm_ws<false, true>(u, count, st, tiles_per_cta))) return rc;
      }
    }
  }
  return 0;
}

// Kinv (lower tiles) = Z^T Z with Z = L^-1 lower (diagonal 128-blocks carry explicit zeros above the diagonal).
// sel_block > 0: only the tiles that intersect the diagonal blocks of that size (all the gradient of a diagonal F needs).
int lauum_lower(const double* Z, int n, long ld, long strideZ, int batch, double* Kinv, long ldk, long strideK, int sel_block, cudaStream_t st) {
  GemmArgs g{};
  g.sel_block = sel_block;
  g.A = Z; g.lda = ld; g.strideA = strideZ;
  g.B = Z; g.ldb = ld; g.strideB = strideZ;
  g.C = Kinv; g.ldc = ldk; g.strideC = strideK;
  g.M = g.N = g.K = n; g.alpha = 1.0; g.beta = 0.0; g.lower_only = 1; g.kmode = K_GE_M0;
  return launch_gemm_ws<true, true>(g, batch, st);
}

// ----------------------------------------------------------------------------------------------------------------
// potrf + trtri with the inverse OVERLAPPED into the factorisation (one matrix).
//
// The critical path of the right-looking Cholesky (diagonal factor -> panel solve, ~80 us per 128-block, ~10 ms at n = 16384) leaves
// most SMs idle; trtri has no such chain.  With the block columns cut into `panels` column panels P_i = [s_i, s_i+1),
//     Z_ii        = trtri(L_ii)                                   needs columns < s_i+1 of L
//     T_i         = L[s_i+1:, P_i] * Z_ii                         needs columns < s_i+1 of L  (final for all rows once potrf passes s_i+1)
//     Z[s_i+1:, P_i] = -Z[s_i+1:, s_i+1:] * T_i                   needs the inverse of the whole trailing block: after the factorisation, i descending
// the first two lines (a third of the trtri flops at 8 panels) are enqueued on a low-priority SIDE stream as soon as the factorisation
// has passed s_i+1, and the factorisation itself runs on an internal HIGH-priority stream: the T_i products are YIELDING launches (one
// tile per CTA, launch_gemm_ws), so their CTAs fill whatever SMs the chain leaves idle and the block scheduler hands every SM that
// retires to the pending chain / trailing-update CTAs first.  Same flops as trtri_lower (n^3/3), same tile kernel; every
// tile is still computed by one CTA in a fixed order, so the result is bitwise reproducible (it differs in the last bits from
// trtri_lower's pairing).  Works under stream capture (both internal streams are forked from and joined to `st` by events).
// ----------------------------------------------------------------------------------------------------------------
namespace {
constexpr int OV_MAX_PANELS = 16;
struct OverlapCtx {
  int device = -1;
  cudaStream_t user = nullptr;
  cudaStream_t side = nullptr, hi = nullptr, back = nullptr;       // bulk / chain / background (lowest priority)
  cudaEvent_t fork[OV_MAX_PANELS + 1] = {};
  cudaEvent_t fork0 = nullptr, join = nullptr, join_hi = nullptr, early_fork = nullptr, early_done = nullptr;
};
// one side stream + event set per (device, caller stream): two evaluations in flight on two streams do not serialise each other
OverlapCtx* overlap_ctx(int device, cudaStream_t user) {
  constexpr int MAX_CTX = 32;
  static OverlapCtx ctx[MAX_CTX];
  static int used = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (ctx[i].device == device && ctx[i].user == user) return &ctx[i];
  if (used == MAX_CTX) return nullptr;
  OverlapCtx& c = ctx[used];
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;   // lo = least priority (what default streams have), hi = greatest
  // three levels where the device has them (B200: 0 .. -5): chain > bulk > background; with two levels bulk and background share the lower one
  int mid = (lo + hi) / 2;
  if (mid == hi && lo != hi) mid = lo;
  if (cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, mid) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&c.hi, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&c.back, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&c.early_fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&c.early_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&c.fork0, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&c.join_hi, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  for (int i = 0; i <= OV_MAX_PANELS; ++i)
    if (cudaEventCreateWithFlags(&c.fork[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  c.device = device;
  c.user = user;
  ++used;
  return &c;
}
}  // namespace

// ----------------------------------------------------------------------------------------------------------------
// potrf with LOOK-AHEAD (one matrix).
//
// Right-looking order leaves the GPU idle along the chain  diagonal factor (1 CTA, ~41 us) -> panel solve (~25 us) -> column update
// of the next block column: 128 x 66 us + 87 tall-skinny updates = 12 of the 52.7 ms of a 16384 factorisation (round-1 launch list).
// With groups g of block columns C_g (the same widths potrf_core uses) and P_g = L[:, C_g]:
//     F_g        factor group g (diagonal factors, panel solves, column updates inside the group)        needs every update of C_g
//     U_g        A[c_g+1:, C_g+1] -= P_g P_g[C_g+1]^T      the NEXT group's columns only (column-limited lower tile list)
//     B_g        A[c_g+2:, c_g+2:] -= P_g P_g^T            the rest of the trailing update, lower tiles
// the CHAIN  F_0 U_0 F_1 U_1 F_2 ...  runs on an internal high-priority stream and the BULK  B_0 B_1 ...  on an internal low-priority one as
// YIELDING launches (each CTA retires after RC_POTRF_YIELD tiles, default 2): B_g only needs F_g, and U_g+1 only needs B_g (both write
// C_g+2, always in this order), so while B_g runs the chain of group g+1 is scheduled onto every SM a bulk CTA gives back - the block
// scheduler serves the higher-priority stream first.  The chain's work still costs SM time, but no SM waits for it any more.
// Same tiles, same K ranges, same order of updates per tile as potrf_core: the factor is bit-for-bit the same.  Capturable (fork/join by
// events).  The last groups (fewer than LA_TAIL blocks left, where a bulk update is a fraction of a wave) run the plain sequence on the
// chain stream.  RC_POTRF_LOOKAHEAD=0 switches it off.
// ----------------------------------------------------------------------------------------------------------------
static bool lookahead_env() {
  static const int v = env_int("RC_POTRF_LOOKAHEAD", 1);
  return v != 0;
}

// early_blocks > 0 (a power of two, < n / 128) with early_tmp: as soon as the chain has passed block column early_blocks, the inverse of the
// leading early_blocks x early_blocks block of L - which is final from then on - is formed on the background stream (yielding launches, lowest
// priority): it fills the SMs that the chain-bound second half of the factorisation leaves idle.  On return A's leading block holds Z11 = L11^-1;
// the caller finishes with trtri_lower(..., rest_from = 128 * early_blocks).  Returns LA_NO_EARLY if the group boundaries never allowed it.
constexpr int LA_NO_EARLY = -7;
static int potrf_lookahead(double* A, int n, long ld, double* dinv, double* logdet_parts, int* info, cudaStream_t st, OverlapCtx* cx,
                           int early_blocks = 0, double* early_tmp = nullptr) {
  const int nblk = n / DB;
  const long strideD = (long)nblk * DB * DB;
  static const int LA_TAIL = env_int("RC_POTRF_LA_TAIL", 16), YIELD = std::max(1, env_int("RC_POTRF_YIELD", 2));
  int start[DB * 4], G = 0;                      // group boundaries, as potrf_core chooses them
  for (int b0 = 0, w = 0; b0 < nblk; b0 += w) {
    w = potrf_group_at(b0, nblk);
    start[G++] = b0;
    RC_REQUIRE(G < DB * 4 - 1, -2, "potrf_lookahead: too many block-column groups");
  }
  start[G] = nblk;
  cudaStream_t hi = cx->hi, lo = cx->side;
  RC_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int), st));
  RC_CUDA_OK(cudaEventRecord(cx->fork0, st));
  RC_CUDA_OK(cudaStreamWaitEvent(hi, cx->fork0, 0));
  int rc;
  auto factor_group_on_chain = [&](int g) -> int {
    return factor_group(A, n, ld, 0, 1, dinv, strideD, logdet_parts, nblk, info, start[g], start[g + 1] - start[g], hi);
  };
  // lower tiles of the square that starts at block `from`, rank = group g; col_tiles > 0: only its first col_tiles tile columns
  auto update = [&](int g, int from, int col_tiles, cudaStream_t stream, int yield) -> int {
    const long r0 = (long)from * DB;
    if (r0 >= n) return 0;
    GemmArgs u{};
    u.A = A + r0 * ld + (long)start[g] * DB; u.lda = ld;
    u.B = u.A; u.ldb = ld;
    u.C = A + r0 * ld + r0; u.ldc = ld;
    u.M = u.N = n - (int)r0; u.K = (start[g + 1] - start[g]) * DB; u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1; u.kmode = K_FULL;
    u.col_tiles = col_tiles;
    return launch_gemm_ws<false, false>(u, 1, stream, yield);
  };
  cudaEvent_t evF[2] = {cx->fork[0], cx->fork[1]}, evB[2] = {cx->fork[2], cx->fork[3]};
  bool bulk_pending = false;                     // a bulk update whose event the chain has not waited for yet
  bool early_launched = false;
  int last_bulk = -1;
  // RC_POTRF_TIMELINE=1 (diagnostic, host-blocking): timestamps after every F_g / U_g on the chain and every B_g on the bulk stream
  static const bool timeline = env_int("RC_POTRF_TIMELINE", 0) != 0;
  std::vector<cudaEvent_t> tl_ev;
  std::vector<std::pair<char, int>> tl_tag;
  auto mark = [&](char what, int g, cudaStream_t stream) {
    if (!timeline) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream);
    tl_ev.push_back(e);
    tl_tag.push_back({what, g});
  };
  mark('0', 0, hi);
  for (int g = 0; g < G; ++g) {
    if ((rc = factor_group_on_chain(g))) return rc;
    mark('F', g, hi);
    if (early_blocks > 0 && !early_launched && start[g + 1] >= early_blocks && start[g + 1] < nblk) {
      static const int YIELD_BACK = std::max(1, env_int("RC_EARLY_TRTRI_YIELD", 1));
      RC_CUDA_OK(cudaEventRecord(cx->early_fork, hi));
      RC_CUDA_OK(cudaStreamWaitEvent(cx->back, cx->early_fork, 0));
      if ((rc = launch_diag_invert_all(A, ld, 0, dinv, strideD, early_blocks, 1, cx->back))) return rc;
      if ((rc = trtri_lower(A, early_blocks * DB, ld, 0, 1, dinv, early_tmp, 0, cx->back, 0, nblk, YIELD_BACK))) return rc;
      RC_CUDA_OK(cudaEventRecord(cx->early_done, cx->back));
      early_launched = true;
      mark('E', g, hi);
    }
    if (g + 1 == G) break;
    const bool split = g + 2 < G && nblk - start[g + 2] >= LA_TAIL;
    if (split) {
      RC_CUDA_OK(cudaEventRecord(evF[g & 1], hi));
      RC_CUDA_OK(cudaStreamWaitEvent(lo, evF[g & 1], 0));
    }
    if (bulk_pending) {                          // B_(g-1) wrote the columns U_g / the plain update touches: keep the order
      RC_CUDA_OK(cudaStreamWaitEvent(hi, evB[last_bulk & 1], 0));
      bulk_pending = false;
    }
    if (split) {
      if ((rc = update(g, start[g + 2], 0, lo, YIELD))) return rc;
      RC_CUDA_OK(cudaEventRecord(evB[g & 1], lo));
      mark('B', g, lo);
      bulk_pending = true;
      last_bulk = g;
      if ((rc = update(g, start[g + 1], start[g + 2] - start[g + 1], hi, 0))) return rc;
    } else if ((rc = update(g, start[g + 1], 0, hi, 0))) {
      return rc;
    }
    mark('U', g, hi);
  }
  {   // the 128 x 128 inverses of the blocks the background work has not done already
    const int b_first = early_launched ? early_blocks : 0;
    if ((rc = launch_diag_invert_all(A + (long)b_first * DB * (ld + 1), ld, 0, dinv + (long)b_first * DB * DB, strideD, nblk - b_first, 1, hi))) return rc;
  }
  RC_CUDA_OK(cudaEventRecord(cx->join_hi, hi));
  RC_CUDA_OK(cudaStreamWaitEvent(st, cx->join_hi, 0));
  if (early_launched) RC_CUDA_OK(cudaStreamWaitEvent(st, cx->early_done, 0));
  if (last_bulk >= 0) {                          // every bulk update has been waited for by the chain already; join the stream itself for capture
    RC_CUDA_OK(cudaEventRecord(cx->join, lo));
    RC_CUDA_OK(cudaStreamWaitEvent(st, cx->join, 0));
  }
  if (timeline) {
    cudaStreamSynchronize(st);
    fprintf(stderr, "potrf_lookahead timeline n=%d groups=%d (ms since start):", n, G);
    for (size_t i = 1; i < tl_ev.size(); ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, tl_ev[0], tl_ev[i]);
      fprintf(stderr, " %c%d=%.3f", tl_tag[i].first, tl_tag[i].second, t);
    }
    fprintf(stderr, "\n");
    for (cudaEvent_t e : tl_ev) cudaEventDestroy(e);
  }
  return (early_blocks > 0 && !early_launched) ? LA_NO_EARLY : 0;
}

int potrf_lower(double* A, int n, long ld, long strideA, int batch, double* dinv, double* logdet_parts, int* info, cudaStream_t st, bool lookahead) {
  static const int LA_MIN = env_int("RC_POTRF_LA_MIN_BLOCKS", 32);
  if (lookahead && batch == 1 && lookahead_env() && n % DB == 0 && n / DB >= LA_MIN && ld >= n && ld % 2 == 0) {
    int dev = 0;
    RC_CUDA_OK(cudaGetDevice(&dev));
    if (OverlapCtx* cx = overlap_ctx(dev, st)) return potrf_lookahead(A, n, ld, dinv, logdet_parts, info, st, cx);
  }
  return potrf_core(A, n, ld, strideA, batch, dinv, logdet_parts, info, st, [](int) { return 0; }, true);
}

size_t potrf_trtri_tmp_doubles(int n, int panels) {
  const long nblk = n / DB;
  if (panels < 2 || nblk < 2 * panels) return (size_t)n * n / 4;
  long wblk = (nblk + panels - 1) / panels;
  wblk = (wblk + 7) / 8 * 8;
  size_t total = (size_t)(wblk * DB) * (wblk * DB) / 4;   // scratch of the panel-local trtri
  for (long s = 0; s + wblk < nblk; s += wblk) total += (size_t)(nblk - s - wblk) * DB * wblk * DB;
  return total;
}

int potrf_trtri_lower(double* A, int n, long ld, double* dinv, double* logdet_parts, int* info, double* tmp, size_t tmp_doubles, int panels,
                      cudaStream_t st, bool lookahead) {
  RC_REQUIRE(n > 0 && n % DB == 0, -2, "potrf_trtri_lower: n=%d must be a positive multiple of 128", n);
  const int nblk = n / DB;
  int rc;
  if (panels > OV_MAX_PANELS) panels = OV_MAX_PANELS;
  int dev = 0;
  RC_CUDA_OK(cudaGetDevice(&dev));
  const bool aligned = potrf_group_env() == 0 || (8 % potrf_group_env()) == 0;      // panel boundaries (multiples of 8 blocks) must fall on group boundaries
  OverlapCtx* cx = (aligned && panels >= 2 && nblk >= 2 * panels && potrf_trtri_tmp_doubles(n, panels) <= tmp_doubles) ? overlap_ctx(dev, st) : nullptr;
  if (!cx) {   // no panel overlap asked for (the default), small problem or no scratch
    RC_REQUIRE(tmp_doubles >= (size_t)n * n / 4, -2, "potrf_trtri_lower: scratch too small");
    static const bool early_on = env_int("RC_EARLY_TRTRI", 1) != 0;
    static const int la_min = env_int("RC_POTRF_LA_MIN_BLOCKS", 32);
    if (lookahead && early_on && lookahead_env() && nblk >= la_min && ld >= n && ld % 2 == 0) {
      // look-ahead factorisation with the inverse of the leading block (the largest power of two of blocks below nblk) formed in its second half
      int early = 1;
      while (2 * early < nblk) early *= 2;
      if (OverlapCtx* la = overlap_ctx(dev, st)) {
        rc = potrf_lookahead(A, n, ld, dinv, logdet_parts, info, st, la, early, tmp);
        if (rc == 0) return trtri_lower(A, n, ld, 0, 1, dinv, tmp, 0, st, early * DB);
        if (rc != LA_NO_EARLY) return rc;
        return trtri_lower(A, n, ld, 0, 1, dinv, tmp, 0, st);        // the group boundaries never reached the block: plain inverse
      }
    }
    if ((rc = potrf_lower(A, n, ld, 0, 1, dinv, logdet_parts, info, st, lookahead))) return rc;
    return trtri_lower(A, n, ld, 0, 1, dinv, tmp, 0, st);
  }
  int wblk = (nblk + panels - 1) / panels;
  wblk = (wblk + 7) / 8 * 8;                          // panel boundaries on the factorisation's group steps (any RC_POTRF_GROUP in 1, 2, 4, 8)
  const int np = (nblk + wblk - 1) / wblk;
  double* scratch = tmp;                              // panel-local trtri scratch, then the T_i one after the other
  double* Tbase = tmp + (size_t)(wblk * DB) * (wblk * DB) / 4;
  double* Tptr[OV_MAX_PANELS + 1];
  {
    double* t = Tbase;
    for (int i = 0; i < np; ++i) {
      Tptr[i] = t;
      const long s1 = (long)(i + 1) * wblk;
      if (s1 < nblk) t += (size_t)(nblk - s1) * DB * wblk * DB;
    }
  }
  cudaStream_t side = cx->side, hi = cx->hi;
  RC_CUDA_OK(cudaEventRecord(cx->fork0, st));
  RC_CUDA_OK(cudaStreamWaitEvent(hi, cx->fork0, 0));
  // side work of panel i, forked from the factorisation stream at its current position
  auto panel_side_work = [&](int i) -> int {
    const int s0 = i * wblk, s1 = std::min(nblk, (i + 1) * wblk), wn = (s1 - s0) * DB;
    RC_CUDA_OK(cudaEventRecord(cx->fork[i], hi));
    RC_CUDA_OK(cudaStreamWaitEvent(side, cx->fork[i], 0));
    double* Aii = A + (long)s0 * DB * (ld + 1);
    double* Dii = dinv + (long)s0 * DB * DB;
    int r;
    if ((r = launch_diag_invert_all(Aii, ld, 0, Dii, 0, s1 - s0, 1, side))) return r;
    if ((r = trtri_lower(Aii, wn, ld, 0, 1, Dii, scratch, 0, side))) return r;
    if (s1 < nblk) {   // T_i = L[s1:, P_i] * Z_ii, yielding
      GemmArgs g{};
      g.A = A + (long)s1 * DB * ld + (long)s0 * DB; g.lda = ld;
      g.B = Aii; g.ldb = ld;
      g.C = Tptr[i]; g.ldc = wn;
      g.M = (nblk - s1) * DB; g.N = wn; g.K = wn; g.alpha = 1.0; g.beta = 0.0; g.kmode = K_GE_N0;
      if ((r = launch_gemm_ws<false, true>(g, 1, side, 1))) return r;
    }
    return 0;
  };
  auto after_panel = [&](int done) -> int {
    if (done % wblk == 0 || done == nblk) return panel_side_work((done - 1) / wblk);
    return 0;
  };
  if ((rc = potrf_core(A, n, ld, 0, 1, dinv, logdet_parts, info, hi, after_panel, false))) return rc;
  RC_CUDA_OK(cudaEventRecord(cx->join_hi, hi));
  RC_CUDA_OK(cudaEventRecord(cx->join, side));
  RC_CUDA_OK(cudaStreamWaitEvent(st, cx->join_hi, 0));
  RC_CUDA_OK(cudaStreamWaitEvent(st, cx->join, 0));
  for (int i = np - 2; i >= 0; --i) {   // Z[s1:, P_i] = -Z[s1:, s1:] * T_i
    const int s0 = i * wblk, s1 = (i + 1) * wblk, wn = wblk * DB;
    const long r = (long)s1 * DB;
    GemmArgs u{};
    u.A = A + r * ld + r; u.lda = ld;
    u.B = Tptr[i]; u.ldb = wn;
    u.C = A + r * ld + (long)s0 * DB; u.ldc = ld;
    u.M = n - (int)r; u.N = wn; u.K = n - (int)r; u.alpha = -1.0; u.beta = 0.0; u.kmode = K_LT_M1;
    if ((rc = launch_gemm_ws<false, true>(u, 1, st))) return rc;
  }
  return 0;
}

// C = alpha * A^T A + beta * C for an n x c block A (row-major, both multiples of 128): every 128-tile of the c x c result.
int syrk_tn(const double* A, int n, int c, long lda, long strideA, int batch, double alpha, double beta, double* C, long ldc, long strideC, cudaStream_t st) {
  RC_REQUIRE(n > 0 && c > 0 && n % DB == 0 && c % DB == 0, -2, "syrk_tn: n=%d and c=%d must be positive multiples of 128", n, c);
  GemmArgs g{};
  g.A = A; g.lda = lda; g.strideA = strideA;
  g.B = A; g.ldb = lda; g.strideB = strideA;
  g.C = C; g.ldc = ldc; g.strideC = strideC;
  g.M = g.N = c; g.K = n; g.alpha = alpha; g.beta = beta; g.kmode = K_FULL;
  return launch_gemm_ws<true, true>(g, batch, st);
}

// dots[pair(l > l')][i] = sum_{k >= l*N+i} Z[k][l*N+i] * Z[k][l'*N+i] = K^-1[(l,i),(l',i)]: the diagonals of the off-diagonal
// (l,l') blocks of K^-1 = Z^T Z, which is all of those blocks that dLML/dE needs.  Row-split partial sums, fixed-order finish.
constexpr int BD_SPLIT = 16;

__global__ void __launch_bounds__(128) block_diag_dots_partial_kernel(const double* __restrict__ Z, long ld, int n, int N, int L, double* __restrict__ parts) {
  const int pair = blockIdx.y, split = blockIdx.z;
  int l = (int)((sqrt(8.0 * pair + 1.0) + 1.0) * 0.5);
  while (l * (l - 1) / 2 > pair) --l;
  while ((l + 1) * l / 2 <= pair) ++l;
  const int lp = pair - l * (l - 1) / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long c1 = (long)l * N + i, c2 = (long)lp * N + i;
  // rows k in [l*N + chunk start, ...) restricted to k >= c1; the split is over the rows of the whole matrix
  const long rows = (n + BD_SPLIT - 1) / BD_SPLIT;
  const long k0 = max((long)split * rows, c1), k1 = min((long)n, (long)(split + 1) * rows);
  double s = 0.0;
  for (long k = k0; k < k1; ++k) s = fma(Z[k * ld + c1], Z[k * ld + c2], s);
  parts[((long)pair * BD_SPLIT + split) * N + i] = s;
}

__global__ void block_diag_dots_finish_kernel(const double* __restrict__ parts, int N, double* __restrict__ dots) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = 0.0;
  for (int sp = 0; sp < BD_SPLIT; ++sp) s += parts[((long)pair * BD_SPLIT + sp) * N + i];
  dots[(long)pair * N + i] = s;
}

size_t block_diag_dots_workspace_bytes(int N, int L) { return (size_t)(L * (L - 1) / 2) * BD_SPLIT * N * sizeof(double); }

int block_diag_dots(const double* Z, long ld, int n, int N, int L, double* parts, double* dots, cudaStream_t st) {
  const int pairs = L * (L - 1) / 2;
  if (pairs == 0) return 0;
  block_diag_dots_partial_kernel<<<dim3((N + 127) / 128, pairs, BD_SPLIT), 128, 0, st>>>(Z, ld, n, N, L, parts);
  RC_LAUNCH_OK();
  block_diag_dots_finish_kernel<<<dim3((N + 127) / 128, pairs), 128, 0, st>>>(parts, N, dots);
  RC_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Triangular matrix-vector products with Z = L^-1 (lower, row-major):  out = Z v  or  out = Z^T v.
// Once Z exists (gradient path) these replace the two block-sequential substitutions: one pass over the triangle (HBM bound)
// instead of 2 x n/128 dependent launches.  Fixed reduction order: bitwise reproducible.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tri_gemv_n_kernel(const double* __restrict__ Z, long ld, long strideZ, const double* __restrict__ v,
                                                         double* __restrict__ out, long strideV, int n) {
  const int lane = threadIdx.x & 31, z = blockIdx.y;
  const long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5);   // one warp per row
  if (i >= n) return;
  const double* row = Z + (long)z * strideZ + i * ld;
  const double* vz = v + (long)z * strideV;
  double s0 = 0.0, s1 = 0.0;
  // columns 0..i in double2 steps; the element right of the diagonal (if any) is an explicit zero of the diagonal block or masked here
  for (long j = 2 * lane; j <= i; j += 64) {
    const double2 a = *reinterpret_cast<const double2*>(row + j);
    const double2 b = *reinterpret_cast<const double2*>(vz + j);
    s0 = fma(a.x, b.x, s0);
    if (j + 1 <= i) s1 = fma(a.y, b.y, s1);
  }
  const double s = warp_sum(s0 + s1);
  if (lane == 0) out[(long)z * strideV + i] = s;
}

constexpr int TG_SPLIT = 32;

__global__ void __launch_bounds__(128) tri_gemv_t_partial_kernel(const double* __restrict__ Z, long ld, long strideZ, const double* __restrict__ v,
                                                                 long strideV, int n, double* __restrict__ parts) {
  const int z = blockIdx.z, split = blockIdx.y;
  const long j = (long)blockIdx.x * 128 + threadIdx.x;
  const long rows = ((long)n + TG_SPLIT - 1) / TG_SPLIT;
  const long i1 = min((long)n, (long)(split + 1) * rows);
  const long i0 = max((long)split * rows, (long)blockIdx.x * 128);   // no row above this column block contributes
  double s = 0.0;
  if (j < n) {
    const double* Zz = Z + (long)z * strideZ;
    const double* vz = v + (long)z * strideV;
    for (long i = i0; i < i1; ++i)
      if (i >= j) s = fma(Zz[i * ld + j], vz[i], s);
  }
  if (j < n) parts[((long)z * TG_SPLIT + split) * n + j] = s;
}

__global__ void tri_gemv_t_finish_kernel(const double* __restrict__ parts, int n, double* __restrict__ out, long strideV) {
  const int z = blockIdx.y;
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int sp = 0; sp < TG_SPLIT; ++sp) s += parts[((long)z * TG_SPLIT + sp) * n + j];
  out[(long)z * strideV + j] = s;
}

size_t tri_gemv_workspace_bytes(int n, int batch) { return (size_t)batch * TG_SPLIT * n * sizeof(double); }

int tri_gemv_lower(const double* Z, int n, long ld, long strideZ, int batch, const double* v, double* out, long strideV, int transpose,
                   double* parts, cudaStream_t st) {
  RC_REQUIRE(ld % 2 == 0, -2, "tri_gemv_lower: ld=%ld must be even", ld);
  if (!transpose) {
    tri_gemv_n_kernel<<<dim3((n + 7) / 8, batch), 256, 0, st>>>(Z, ld, strideZ, v, out, strideV, n);
    RC_LAUNCH_OK();
  } else {
    tri_gemv_t_partial_kernel<<<dim3((n + 127) / 128, TG_SPLIT, batch), 128, 0, st>>>(Z, ld, strideZ, v, strideV, n, parts);
    RC_LAUNCH_OK();
    tri_gemv_t_finish_kernel<<<dim3((n + 127) / 128, batch), 128, 0, st>>>(parts, n, out, strideV);
    RC_LAUNCH_OK();
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// small utilities
// ----------------------------------------------------------------------------------------------------------------
__global__ void sum_parts_kernel(const double* __restrict__ parts, int count, double* __restrict__ out, double scale) {
  __shared__ double red[32];
  const int z = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s += parts[(long)z * count + i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[z] = s * scale;
}

int sum_parts(const double* parts, int count, int batch, double* out, double scale, cudaStream_t st) {
  sum_parts_kernel<<<batch, 256, 0, st>>>(parts, count, out, scale);
  RC_LAUNCH_OK();
  return 0;
}

__global__ void dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long n, long stride, double* __restrict__ out) {
  __shared__ double red[32];
  const int z = blockIdx.x;
  double s = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) s = fma(a[z * stride + i], b[z * stride + i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[z] = s;
}

int dot_batched(const double* a, const double* b, long n, long stride, int batch, double* out, cudaStream_t st) {
  dot_kernel<<<batch, 1024, 0, st>>>(a, b, n, stride, out);
  RC_LAUNCH_OK();
  return 0;
}

// dst (n x n, ld=n, zero upper) <- lower triangle of src (n_pad storage)
__global__ void extract_lower_kernel(const double* __restrict__ src, long lds, long strideS, double* __restrict__ dst, int n, long strideDst,
                                     int symmetrize) {
  const int z = blockIdx.z;
  const double* s = src + z * strideS;
  double* d = dst + z * strideDst;
  for (long i = blockIdx.y; i < n; i += gridDim.y)              // rows in a grid-stride loop: gridDim.y is limited to 65535
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long)gridDim.x * blockDim.x) {
      double v;
      if (j <= i) v = s[i * lds + j];
      else v = symmetrize ? s[j * lds + i] : 0.0;
      d[i * n + j] = v;
    }
}

int extract_lower(const double* src, long lds, long strideS, double* dst, int n, long strideDst, int batch, int symmetrize, cudaStream_t st) {
  dim3 grid((n + 1023) / 1024 > 0 ? (n + 1023) / 1024 : 1, n < 65535 ? n : 65535, batch);
  extract_lower_kernel<<<grid, 256, 0, st>>>(src, lds, strideS, dst, n, strideDst, symmetrize);
  RC_LAUNCH_OK();
  return 0;
}

// dst (n_pad storage) <- src (n x n dense) with identity padding
__global__ void pad_identity_kernel(const double* __restrict__ src, int n, long strideS, double* __restrict__ dst, int n_pad, long ldd, long strideD) {
  const int z = blockIdx.z;
  for (long i = blockIdx.y; i < n_pad; i += gridDim.y)
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n_pad; j += (long)gridDim.x * blockDim.x) {
      double v = (i < n && j < n) ? src[z * strideS + i * n + j] : (i == j ? 1.0 : 0.0);
      dst[z * strideD + i * ldd + j] = v;
    }
}

int pad_identity(const double* src, int n, long strideS, double* dst, int n_pad, long ldd, long strideD, int batch, cudaStream_t st) {
  dim3 grid((n_pad + 1023) / 1024, n_pad < 65535 ? n_pad : 65535, batch);
  pad_identity_kernel<<<grid, 256, 0, st>>>(src, n, strideS, dst, n_pad, ldd, strideD);
  RC_LAUNCH_OK();
  return 0;
}

// Host mirror of the device tile order (gemm_decode_tile is __host__ __device__): what rc_debug_tile_order reports.
int debug_tile_order(int M, int N, int K, int lower_only, int kmode, int sel_block, int* out) {
  RC_REQUIRE(M > 0 && N > 0 && K > 0 && M % G_BM == 0 && N % G_BN == 0 && K % G_BK == 0 && out, -2, "rc_debug_tile_order: bad shape");
  RC_REQUIRE(kmode >= K_FULL && kmode <= K_LE_N1, -2, "rc_debug_tile_order: kmode %d", kmode);
  RC_REQUIRE(!lower_only || M == N, -2, "rc_debug_tile_order: lower_only needs M == N");
  GemmArgs p{};
  p.M = M; p.N = N; p.K = K; p.lower_only = lower_only; p.kmode = kmode;
  p.sel_block = sel_block > 0 ? sel_block : 0;
  p.col_tiles = sel_block < 0 ? -sel_block : 0;      // negative sel_block: the column-limited lower list of the look-ahead panel update
  RC_REQUIRE(p.col_tiles == 0 || (lower_only && kmode == K_FULL), -2, "rc_debug_tile_order: a column limit needs lower_only and kmode 0");
  const long tiles = gemm_tile_count(p);
  for (long t = 0; t < tiles; ++t) {
    const GemmTile T = lower_only ? gemm_decode_tile<1>(p, t, tiles) : gemm_decode_tile<0>(p, t, tiles);
    out[4 * t + 0] = T.m0; out[4 * t + 1] = T.n0; out[4 * t + 2] = T.kb; out[4 * t + 3] = T.nk;
  }
  return (int)tiles;
}

}  // namespace rc
