// Warp-specialised FP64 tensor-core (DMMA.8x8x4) tile GEMM: same contract as gemm_dmma.cuh (C = alpha * Aop * Bop^T + beta * C over a
// persistent tile list, triangular structure at tile granularity), different feed.
//
//   * one PRODUCER warp streams operand tiles global -> shared with TMA bulk copies (cp.async.bulk, SASS UBLKCP) that complete on an
//     mbarrier per stage ("full"); it runs ahead of the math across tile boundaries, limited only by the ring depth;
//   * eight CONSUMER warps (2 x 4, warp tile 64 x 32 = 8 x 4 DMMA tiles, accumulators in registers - FP64 has no tcgen05/TMEM path)
//     wait on "full", issue DMMA from shared memory and release the stage through an "empty" mbarrier.  There is no CTA-wide
//     barrier anywhere in the main loop or the epilogue, so the two warps that share a tensor pipe drift apart and keep it busy
//     while the other one waits, loads fragments or writes its C sub-tile.
//   * tiles are handed out DYNAMICALLY: the producer draws the next tile index from a global counter (one atomicAdd per tile) and
//     passes it to the consumers through a 4-entry shared-memory ring.  Tile lists are ordered by decreasing K range, so with
//     triangular K ranges (trtri, LAUUM) and skipped tiles ("selected" LAUUM) every SM stays busy until the list is empty
//     (static striding left the selected LAUUM at 59 % of peak).  Each tile is still computed by one CTA in a fixed order of
//     operations: results are bitwise independent of the schedule.
//
// Operand staging, by storage order of the operand:
//   * k-contiguous operands ([m][k], "non-transposed"): ONE cp.async.bulk.tensor (TMA tiled load, SASS UTMALDG) per k-slice through a
//     CUtensorMap built per launch, box 16 (k) x 128 (rows) doubles, SWIZZLE_128B: the 16-byte chunk index of a 128-byte row is
//     XORed with (row & 7), which makes every warp fragment read (8 rows x 4 k) hit all 32 banks twice = the 2-wavefront minimum,
//     without padding.  (128 separate 128-byte bulk copies per stage are TMA-issue bound: measured 10.5 TFLOP/s.)
//   * m-contiguous operands ([k][m], "transposed"): 16 one-dimensional bulk copies of 1 KB (SASS UBLKCP) into rows padded by 4 doubles.
#pragma once
#include "gemm_dmma.cuh"
#include <cuda.h>
#include <cstring>

namespace rc {

// Three warpgroups: two of consumers (8 warps) and one for the producer (one working warp).  The launch gives every thread 168 registers
// (65536 / 384); setmaxnreg then moves registers from the producer's warpgroup to the consumers (W_REGS_PRODUCER / W_REGS_CONSUMER:
// 128 x 56 + 256 x 224 = 64512), whose 128 accumulator registers plus fragments, tile state and the batched C loads of the epilogue do not
// fit 168 without spilling.
// A pipeline stage holds W_SUB k-slices of 16 (two TMA boxes per operand: the 128-byte swizzle limits a box to 16 doubles along k) behind ONE
// full/empty barrier pair: half the barrier waits, proxy fences and arrivals per flop of a 16-wide stage.
constexpr int W_SUB = 2;
constexpr int W_STAGES = 3, W_CONSUMERS = 8, W_THREADS = 32 * (W_CONSUMERS + 4), W_RING = 4, W_EPI_ROWS = 4;
constexpr int W_REGS_PRODUCER = 56, W_REGS_CONSUMER = 224;

template <bool TA, bool TB>
struct GemmWsSmem {
  static constexpr int A_STAGE = TA ? G_BK * (G_BM + G_PAD) : G_BM * G_BK;   // doubles; k-contiguous operands are dense + swizzled
  static constexpr int B_STAGE = TB ? G_BK * (G_BN + G_PAD) : G_BN * G_BK;
  static constexpr size_t up1k(size_t b) { return (b + 1023) / 1024 * 1024; }
  static constexpr size_t A_BYTES = up1k((size_t)W_STAGES * W_SUB * A_STAGE * sizeof(double));
  static constexpr size_t B_BYTES = up1k((size_t)W_STAGES * W_SUB * B_STAGE * sizeof(double));
  static constexpr size_t BYTES = 1024 + A_BYTES + B_BYTES + (2 * W_STAGES + 2 * W_RING) * sizeof(unsigned long long) + W_RING * sizeof(int);
  static constexpr unsigned STAGE_TX = 2u * W_SUB * G_BM * G_BK * sizeof(double);   // bytes landing per stage
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tma_g2s_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}

// One operand stage by the producer warp.  Transposed storage: 16 rows of 1 KB, one bulk copy per lane.  k-contiguous storage: one
// tiled TMA load of the 16 x 128 box at (k, row, z1, z2) relative to the operand pointer the tensor map was built on.
template <bool T>
__device__ __forceinline__ void ws_load_operand(double* dst, const double* src, long ld, const CUtensorMap* map, int z1, int z2, int mn0, int k0, int lane,
                                                unsigned long long* bar) {
  if (!T) {
    if (lane == 0) tma_g2s_4d(dst, map, k0, mn0, z1, z2, bar);
  } else {
    if (lane < G_BK) bulk_g2s(dst + lane * (G_BM + G_PAD), src + (long)(k0 + lane) * ld + mn0, G_BM * sizeof(double), bar);
  }
}

template <bool TA, bool TB, bool LOWER>
__global__ void __launch_bounds__(W_THREADS, 1) gemm_dmma_ws_kernel(GemmArgs p, long tiles_per_matrix, long total_tiles, int* __restrict__ sched,
                                                                    int tiles_per_cta, const __grid_constant__ CUtensorMap mapA,
                                                                    const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(16) double smem_raw[];
  using S = GemmWsSmem<TA, TB>;
  // SWIZZLE_128B needs its tiles on 1 KB boundaries of the shared address space
  char* base = reinterpret_cast<char*>(smem_raw) + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* As = reinterpret_cast<double*>(base);
  double* Bs = reinterpret_cast<double*>(base + S::A_BYTES);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(base + S::A_BYTES + S::B_BYTES);
  unsigned long long* empty = full + W_STAGES;
  unsigned long long* tfull = empty + W_STAGES;     // tile-index ring: producer -> consumers
  unsigned long long* tempty = tfull + W_RING;
  volatile int* ring = reinterpret_cast<volatile int*>(tempty + W_RING);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, W_CONSUMERS);
    }
#pragma unroll
    for (int s = 0; s < W_RING; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, W_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();                                   // everything above overlaps the tail of the previous kernel in the stream

  if (warp >= W_CONSUMERS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(W_REGS_PRODUCER));
    if (warp != W_CONSUMERS) return;          // the other three warps of the producer's warpgroup only give their registers away
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(W_REGS_CONSUMER));
  }

  if (warp == W_CONSUMERS) {
    // ---------------------------------------------------------------- producer
    int stage = 0, slot = 0, drawn = 0;
    unsigned phase = 0, tphase = 0;
    for (;;) {
      long tile = -1;
      GemmTile T;
      for (;;) {   // draw tiles until one has work (skipped tiles of a selected list cost one decode)
        if (tiles_per_cta > 0 && drawn == tiles_per_cta) break;   // yielding launch: this CTA has had its share, retire
        int t = 0;
        if (lane == 0) t = atomicAdd(sched, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total_tiles) break;
        T = gemm_decode_tile<LOWER ? 1 : 0>(p, t, tiles_per_matrix);
        if (T.nk >= 0) {
          tile = t;
          ++drawn;
          break;
        }
      }
      mbar_wait(tempty + slot, tphase ^ 1u);
      if (lane == 0) {
        ring[slot] = (int)tile;
        mbar_arrive(tfull + slot);      // release: the consumers' acquire on tfull sees the index
      }
      if (++slot == W_RING) {
        slot = 0;
        tphase ^= 1u;
      }
      if (tile < 0) break;
      const int z = (int)(tile / tiles_per_matrix);
      const int z2 = p.batch2 > 1 ? z / p.batch1 : 0, z1 = z - z2 * p.batch1;
      // an operand shared by the problems of a level (stride 0) has a single slice in its tensor map
      const int zA1 = p.strideA ? z1 : 0, zA2 = p.strideA2 ? z2 : 0, zB1 = p.strideB ? z1 : 0, zB2 = p.strideB2 ? z2 : 0;
      for (int kt = 0; kt < T.nk; kt += W_SUB) {   // K ranges are multiples of 128: nk is a multiple of W_SUB
        mbar_wait(empty + stage, phase ^ 1u);
        if (lane == 0) mbar_arrive_expect_tx(full + stage, S::STAGE_TX);
        __syncwarp();
#pragma unroll
        for (int sub = 0; sub < W_SUB; ++sub) {
          ws_load_operand<TA>(As + (stage * W_SUB + sub) * S::A_STAGE, T.A, p.lda, &mapA, zA1, zA2, T.m0, T.kb + (kt + sub) * G_BK, lane, full + stage);
          ws_load_operand<TB>(Bs + (stage * W_SUB + sub) * S::B_STAGE, T.B, p.ldb, &mapB, zB1, zB2, T.n0, T.kb + (kt + sub) * G_BK, lane, full + stage);
        }
        if (++stage == W_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    // The last CTA to run dry re-arms the counter pair for the next launch that uses this slot.
    if (lane == 0) {
      const int done = atomicAdd(sched + 1, 1);
      if (done == (int)gridDim.x - 1) {
        sched[0] = 0;
        sched[1] = 0;
        __threadfence();
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  const int swz = ((t >> 1) ^ g), todd = t & 1;    // swizzled k-contiguous tiles: double index = row*16 + (((2*kk) ^ swz) << 1) + todd
  int stage = 0, slot = 0;
  unsigned phase = 0, tphase = 0;
  for (;;) {
    mbar_wait(tfull + slot, tphase);
    const int tile = ring[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty + slot);
    if (++slot == W_RING) {
      slot = 0;
      tphase ^= 1u;
    }
    if (tile < 0) break;
    const GemmTile T = gemm_decode_tile<LOWER ? 1 : 0>(p, tile, tiles_per_matrix);
    if (p.beta != 0.0) {   // pull this warp's 64 x 32 part of C (64 rows x 256 B) towards L2 while the main loop runs
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int line = lane + r * 32;                 // 128 lines of 128 B
        const double* addr = T.C + (long)(T.m0 + wm * 64 + (line >> 1)) * p.ldc + T.n0 + wn * 32 + (line & 1) * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
      }
    }
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < T.nk; kt += W_SUB) {
      mbar_wait(full + stage, phase);
      __syncwarp();
#pragma unroll
      for (int sub = 0; sub < W_SUB; ++sub) {
      const double* as = As + (stage * W_SUB + sub) * S::A_STAGE;
      const double* bs = Bs + (stage * W_SUB + sub) * S::B_STAGE;
#pragma unroll
      for (int kk = 0; kk < G_BK / 4; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          a[i] = TA ? as[(kk * 4 + t) * (G_BM + G_PAD) + wm * 64 + i * 8 + g] : as[(wm * 64 + i * 8 + g) * G_BK + (((2 * kk) ^ swz) << 1) + todd];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          b[j] = TB ? bs[(kk * 4 + t) * (G_BN + G_PAD) + wn * 32 + j * 8 + g] : bs[(wn * 32 + j * 8 + g) * G_BK + (((2 * kk) ^ swz) << 1) + todd];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      }
      // Release the stage to the producer.  The fragment loads above are generic-proxy reads, the refill is an async-proxy (TMA) write:
      // without a cross-proxy fence ptxas may schedule the arrive right after the last LDS is *issued* and the bulk copy of the next
      // lap can overtake its second (bank-conflict) wavefront - observed as rare stale 8x8 fragments before this fence was added.
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
      if (++stage == W_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
    // Every operand byte of this tile has landed in shared memory (this warp waited on the tile's last "full" barrier), so the
    // tile-exclusive in-place updates the drivers rely on (C aliasing A or B of the SAME tile) stay safe.
    const double alpha = p.alpha, beta = p.beta;
    if (beta == 0.0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long row = T.m0 + wm * 64 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = T.n0 + wn * 32 + j * 8 + t * 2;
          *reinterpret_cast<double2*>(T.C + row * p.ldc + col) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        }
      }
    } else {
      // Read-modify-write of C with streaming (evict-first) loads and stores - every C element is touched once per launch, the operand
      // panels are what should stay in L2 -, W_EPI_ROWS fragment rows at a time: all their loads are issued before the first store.  Written as
      // load / fma / store per element the compiler may not move a load above an earlier store (same array, run-time ldc), which made
      // the epilogue eight dependent L2 round trips per warp - most of what a K = 256 trailing-update tile lost against a long-K tile.
#pragma unroll
      for (int i0 = 0; i0 < 8; i0 += W_EPI_ROWS) {
        double2 o[W_EPI_ROWS][4];
#pragma unroll
        for (int i = 0; i < W_EPI_ROWS; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[i][j] = __ldcs(reinterpret_cast<const double2*>(T.C + (long)(T.m0 + wm * 64 + (i0 + i) * 8 + g) * p.ldc + T.n0 + wn * 32 + j * 8 + t * 2));
#pragma unroll
        for (int i = 0; i < W_EPI_ROWS; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double2 v;
            v.x = fma(alpha, acc[i0 + i][j][0], beta * o[i][j].x);
            v.y = fma(alpha, acc[i0 + i][j][1], beta * o[i][j].y);
            __stcs(reinterpret_cast<double2*>(T.C + (long)(T.m0 + wm * 64 + (i0 + i) * 8 + g) * p.ldc + T.n0 + wn * 32 + j * 8 + t * 2), v);
          }
      }
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                           CUtensorMapFloatOOBfill);
inline TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map over a k-contiguous operand: dims (k, rows, inner batch, outer batch), box 16 x 128 x 1 x 1 doubles, 128-byte swizzle.  A batch
// level whose stride is 0 (an operand shared by the problems) or whose count is 1 gets a single slice; the kernel then addresses slice 0.
inline int make_operand_map(CUtensorMap* map, const double* ptr, long ld, long stride, int rows, int K, int batch, long stride2, int batch2) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  RC_REQUIRE(enc != nullptr, -3, "gemm_dmma_ws: cuTensorMapEncodeTiled is not available from this driver");
  const bool level1 = batch > 1 && stride != 0, level2 = batch2 > 1 && stride2 != 0;
  const long dummy = (long)rows * ld;
  const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(level1 ? batch : 1), (cuuint64_t)(level2 ? batch2 : 1)};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * sizeof(double), (cuuint64_t)(level1 ? stride : dummy) * sizeof(double),
                                 (cuuint64_t)(level2 ? stride2 : (level1 ? stride * batch : dummy)) * sizeof(double)};
  const cuuint32_t box[4] = {(cuuint32_t)G_BK, (cuuint32_t)G_BM, 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RC_REQUIRE(r == CUDA_SUCCESS, -3, "gemm_dmma_ws: cuTensorMapEncodeTiled failed (%d) for ptr=%p ld=%ld stride=%ld/%ld rows=%d K=%d batch=%d/%d", (int)r,
             (const void*)ptr, ld, stride, stride2, rows, K, batch, batch2);
  return 0;
}

// Per-device scratch of {next tile, CTAs done} counter pairs for the dynamic tile scheduler: the one piece of device memory this
// library allocates itself (32 KB, once per device, zero-initialised; every launch takes the next pair and leaves it zeroed).
constexpr int SCHED_SLOTS = 4096;
int* gemm_sched_slot(int device);   // defined in chol.cu

// tiles_per_cta = 0: persistent grid (one CTA per SM draws tiles until the list is empty).  tiles_per_cta = k > 0: a YIELDING launch of
// ceil(tiles / k) CTAs that retire after k tiles each - for background work on a low-priority stream: the SMs return to the block
// scheduler every few tiles, so pending CTAs of a higher-priority stream never wait longer than that.
template <bool TA, bool TB, bool LOWER>
inline int launch_gemm_ws_impl(const GemmArgs& a_in, int batch, cudaStream_t stream, int tiles_per_cta) {
  using S = GemmWsSmem<TA, TB>;
  GemmArgs a = a_in;
  a.batch1 = batch;
  const int batch2 = a.batch2 > 1 ? a.batch2 : 1;
  int dev = 0;
  RC_CUDA_OK(cudaGetDevice(&dev));
  RC_ENSURE_SMEM((gemm_dmma_ws_kernel<TA, TB, LOWER>), S::BYTES);
  const int num_sms = device_sm_count();
  if (a.M <= 0 || a.N <= 0 || batch <= 0) return 0;
  RC_REQUIRE(a.M % G_BM == 0 && a.N % G_BN == 0 && a.K % (G_BK * W_SUB) == 0, -2, "gemm_dmma_ws: M,N must be multiples of 128 and K of 32 (got %d,%d,%d)", a.M,
             a.N, a.K);
  RC_REQUIRE(a.lda % 2 == 0 && a.ldb % 2 == 0 && a.ldc % 2 == 0, -2, "gemm_dmma_ws: leading dimensions must be even (16-byte rows)");
  RC_REQUIRE(a.col_tiles == 0 || (a.lower_only && a.kmode == K_FULL && a.sel_block == 0), -2, "gemm_dmma_ws: col_tiles needs a plain lower-triangular list");
  const long tiles = gemm_tile_count(a);
  const long total = tiles * batch * batch2;
  const unsigned grid = tiles_per_cta > 0 ? (unsigned)((total + tiles_per_cta - 1) / tiles_per_cta) : (unsigned)(total < num_sms ? total : num_sms);
  alignas(64) CUtensorMap mapA, mapB;
  memset(&mapA, 0, sizeof(mapA));
  memset(&mapB, 0, sizeof(mapB));
  int rc;
  if (!TA && (rc = make_operand_map(&mapA, a.A, a.lda, a.strideA, a.M, a.K, batch, a.strideA2, batch2))) return rc;
  if (!TB && (rc = make_operand_map(&mapB, a.B, a.ldb, a.strideB, a.N, a.K, batch, a.strideB2, batch2))) return rc;
  const bool prof = profile_enabled();
  if (prof) profile_gemm_begin(stream);
  int* sched = gemm_sched_slot(dev);
  RC_REQUIRE(sched != nullptr, -3, "gemm_dmma_ws: could not allocate the tile-scheduler scratch");
  RC_CUDA_OK(launch_pdl(gemm_dmma_ws_kernel<TA, TB, LOWER>, dim3(grid), dim3(W_THREADS), S::BYTES, stream, a, tiles, total, sched, tiles_per_cta, mapA, mapB));
  if (prof) profile_gemm_end(stream, gemm_tile_flops(a, batch));
  RC_LAUNCH_OK();
  return 0;
}

template <bool TA, bool TB>
inline int launch_gemm_ws(const GemmArgs& a, int batch, cudaStream_t stream, int tiles_per_cta = 0) {
  return a.lower_only ? launch_gemm_ws_impl<TA, TB, true>(a, batch, stream, tiles_per_cta) : launch_gemm_ws_impl<TA, TB, false>(a, batch, stream, tiles_per_cta);
}

}  // namespace rc
