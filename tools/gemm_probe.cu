// Probe: how fast can the 128x128x16 DMMA mainloop run (a) from static shared memory, (b) with the cp.async feed from L2?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../rom-comma_b200/csrc -o gemm_probe gemm_probe.cu
#include "gemm_dmma.cuh"
#include <cstdio>
#include <cstdlib>
namespace rc { void set_error(const char*, ...) {} void count_launches(long) {} bool profile_enabled() { return false; } void profile_gemm_begin(cudaStream_t) {} void profile_gemm_end(cudaStream_t, double) {} }
using namespace rc;

template <int MODE, int BARRIER>   // MODE 0: no loads; 1: cp.async loads
__global__ void __launch_bounds__(256, 1) probe(const double* A, const double* B, long ld, int nk, double* out) {
  extern __shared__ __align__(16) double smem[];
  using S = GemmSmem<false, false>;
  double* As = smem; double* Bs = smem + G_STAGES * S::A_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < G_STAGES * (S::A_STAGE + S::B_STAGE); i += 256) smem[i] = 1e-3 * (i % 7);
  __syncthreads();
  double acc[8][4][2];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int m0 = (blockIdx.x % 64) * 128, n0 = (blockIdx.x / 64) * 128;
  if (MODE == 1) {
    for (int s = 0; s < G_STAGES - 1; ++s) {
      gemm_load_operand<false>(As + s * S::A_STAGE, A, ld, m0, s * G_BK, tid);
      gemm_load_operand<false>(Bs + s * S::B_STAGE, B, ld, n0, s * G_BK, tid);
      cp_async_commit();
    }
  }
  for (int kt = 0; kt < nk; ++kt) {
    if (MODE == 1) cp_async_wait<G_STAGES - 2>();
    if (BARRIER) __syncthreads();
    if (MODE == 1) {
      const int nxt = kt + G_STAGES - 1; const int s = nxt % G_STAGES;
      gemm_load_operand<false>(As + s * S::A_STAGE, A, ld, m0, (nxt * G_BK) % 256, tid);
      gemm_load_operand<false>(Bs + s * S::B_STAGE, B, ld, n0, (nxt * G_BK) % 256, tid);
      cp_async_commit();
    }
    const double* as = As + (kt % G_STAGES) * S::A_STAGE; const double* bs = Bs + (kt % G_STAGES) * S::B_STAGE;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = as[(wm * 64 + i * 8 + g) * (G_BK + G_PAD) + kk * 4 + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[(wn * 32 + j * 8 + g) * (G_BK + G_PAD) + kk * 4 + t];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int MODE, int BARRIER> void run(const char* name, const double* A, const double* B, long ld, double* out, int grid) {
  using S = GemmSmem<false, false>;
  cudaFuncSetAttribute(probe<MODE, BARRIER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES);
  const int nk = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE, BARRIER><<<grid, 256, S::BYTES>>>(A, B, ld, nk, out); cudaDeviceSynchronize();
  cudaEventRecord(e0); probe<MODE, BARRIER><<<grid, 256, S::BYTES>>>(A, B, ld, nk, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s grid %4d: %.2f TFLOP/s (%s)\n", name, grid, 2.0 * 128 * 128 * 16 * nk * grid / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const long ld = 16384; double *A, *out; cudaMalloc(&A, ld * 8192 * 8); cudaMalloc(&out, 64); cudaMemset(A, 0, ld * 8192 * 8);
  for (int grid : {148, 74}) {
    run<0, 0>("static smem, no barrier", A, A, ld, out, grid);
    run<0, 1>("static smem, barrier per k-tile", A, A, ld, out, grid);
    run<1, 1>("cp.async feed (panel in L2), barrier", A, A, ld, out, grid);
  }
  return 0;
}
