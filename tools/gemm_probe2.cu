// Probe: cp.async GEMM (gemm_dmma.cuh) vs warp-specialised TMA-bulk GEMM (gemm_dmma_ws.cuh) on the shapes the drivers launch.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../rom-comma_b200/csrc -o gemm_probe2 gemm_probe2.cu
#include "gemm_dmma_ws.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace rc {
void set_error(const char* fmt, ...) { fprintf(stderr, "rc error: %s\n", fmt); }
void count_launches(long) {}
bool profile_enabled() { return false; }
void profile_gemm_begin(cudaStream_t) {}
void profile_gemm_end(cudaStream_t, double) {}
int* gemm_sched_slot(int) {
  static int* base = nullptr; static unsigned seq = 0;
  if (!base) { cudaMalloc(&base, SCHED_SLOTS * 2 * sizeof(int)); cudaMemset(base, 0, SCHED_SLOTS * 2 * sizeof(int)); }
  return base + 2 * (seq++ % SCHED_SLOTS);
}
}  // namespace rc
using namespace rc;

__global__ void fill(double* p, long n, unsigned seed) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    p[i] = ((double)(h & 0xffff) / 65536.0 - 0.5);
  }
}
__global__ void maxdiff(const double* a, const double* b, long n, double* out) {
  double m = 0.0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) m = fmax(m, fabs(a[i] - b[i]));
  atomicMax((unsigned long long*)out, __double_as_longlong(m));   // non-negative doubles order like integers
}

// direct evaluation of sampled entries: which kernel is right?
template <bool TA, bool TB>
__global__ void sample_check(GemmArgs g, const double* C1, const double* C2, double* err) {   // err[0], err[1]: max |C - ref| of impl 0 / 1
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned h = s * 2654435761u + 12345u;
  h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
  int m = h % g.M; h = h * 1664525u + 1013904223u; int n = (h >> 3) % g.N;
  if (g.lower_only && n > m) { int t = m; m = n; n = t; }
  const int m0 = m / 128 * 128, n0 = n / 128 * 128;
  int kb = 0, ke = g.K;
  if (g.kmode == K_GE_N0) kb = n0; else if (g.kmode == K_LT_M1) ke = min(g.K, m0 + 128); else if (g.kmode == K_GE_M0) kb = m0; else if (g.kmode == K_LE_N1) ke = min(g.K, n0 + 128);
  double acc = 0.0;
  for (int k = kb; k < ke; ++k) acc += (TA ? g.A[(long)k * g.lda + m] : g.A[(long)m * g.lda + k]) * (TB ? g.B[(long)k * g.ldb + n] : g.B[(long)n * g.ldb + k]);
  const double ref = g.alpha * acc;   // beta == 0 cases only
  atomicMax((unsigned long long*)&err[0], __double_as_longlong(fabs(C1[(long)m * g.ldc + n] - ref)));
  atomicMax((unsigned long long*)&err[1], __double_as_longlong(fabs(C2[(long)m * g.ldc + n] - ref)));
}

template <bool TA, bool TB>
void bench(const char* name, GemmArgs g, int batch, double* C1, double* C2, long csize, double* dout) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double flops = gemm_tile_flops(g, batch);
  float ms[2];
  for (int impl = 0; impl < 2; ++impl) {
    g.C = impl ? C2 : C1;
    fill<<<1024, 256>>>(g.C, csize, 7);
    for (int rep = 0; rep < 2; ++rep) {      // second run timed
      if (rep == 1) fill<<<1024, 256>>>(g.C, csize, 7);
      cudaEventRecord(e0);
      int rc = impl ? launch_gemm_ws<TA, TB>(g, batch, 0) : launch_gemm<TA, TB>(g, batch, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      if (rc) printf("launch rc=%d\n", rc);
    }
    cudaEventElapsedTime(&ms[impl], e0, e1);
  }
  cudaMemset(dout, 0, 8);
  maxdiff<<<1024, 256>>>(C1, C2, csize, dout);
  double d;
  cudaMemcpy(&d, dout, 8, cudaMemcpyDeviceToHost);
  if (d > 0) {   // locate the differing entries (host side, slow but rare)
    std::vector<double> h1(csize), h2(csize);
    cudaMemcpy(h1.data(), C1, csize * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), C2, csize * 8, cudaMemcpyDeviceToHost);
    long cnt = 0, first = -1, last = -1;
    long tilehist[8] = {0};
    for (long i = 0; i < csize; ++i) if (h1[i] != h2[i]) { if (first < 0) first = i; last = i; ++cnt; }
    printf("   %ld differing entries; first (%ld,%ld) last (%ld,%ld)\n", cnt, first / g.ldc, first % g.ldc, last / g.ldc, last % g.ldc);
    long shown = 0;
    for (long tm = 0; tm < g.M / 128 && shown < 1; ++tm)
      for (long tn = 0; tn < g.N / 128 && shown < 1; ++tn) {
        long c = 0;
        for (int r = 0; r < 128; ++r) for (int q = 0; q < 128; ++q) c += h1[(tm * 128 + r) * g.ldc + tn * 128 + q] != h2[(tm * 128 + r) * g.ldc + tn * 128 + q];
        if (!c) continue;
        ++shown;
        if (shown == 1) {
          for (int bi = 0; bi < 16; ++bi) for (int bj = 0; bj < 16; ++bj) {
            int cc = 0;
            for (int r = 0; r < 8; ++r) for (int q = 0; q < 8; ++q) cc += h1[(tm * 128 + bi * 8 + r) * g.ldc + tn * 128 + bj * 8 + q] != h2[(tm * 128 + bi * 8 + r) * g.ldc + tn * 128 + bj * 8 + q];
            if (cc && bj < 2) {
              printf("     block (%d,%d) diff (ws - cp.async):\n", bi, bj);
              for (int r = 0; r < 8; ++r) { printf("       "); for (int q = 0; q < 8; ++q) printf("%9.5f", h2[(tm * 128 + bi * 8 + r) * g.ldc + tn * 128 + bj * 8 + q] - h1[(tm * 128 + bi * 8 + r) * g.ldc + tn * 128 + bj * 8 + q]); printf("\n"); }
            }
          }
        }
        printf("     tile (%ld,%ld): %ld differing; 8x8 blocks (rows i=0..15, cols j=0..15), count per block:\n", tm, tn, c);
        for (int bi = 0; bi < 16; ++bi) {
          printf("       ");
          for (int bj = 0; bj < 16; ++bj) {
            int cc = 0;
            for (int r = 0; r < 8; ++r) for (int q = 0; q < 8; ++q) cc += h1[(tm * 128 + bi * 8 + r) * g.ldc + tn * 128 + bj * 8 + q] != h2[(tm * 128 + bi * 8 + r) * g.ldc + tn * 128 + bj * 8 + q];
            printf("%3d", cc);
          }
          printf("\n");
        }
      }
    (void)tilehist;
  }
  double e[2] = {0, 0};
  if (g.beta == 0.0) {
    double* derr; cudaMalloc(&derr, 16); cudaMemset(derr, 0, 16);
    sample_check<TA, TB><<<16, 256>>>(g, C1, C2, derr);
    cudaMemcpy(e, derr, 16, cudaMemcpyDeviceToHost); cudaFree(derr);
  }
  printf("%-44s cp.async %8.3f ms %6.2f TF | ws %8.3f ms %6.2f TF | maxdiff %.3e  err vs direct: %.2e / %.2e (%s)\n", name, ms[0], flops / ms[0] * 1e-9, ms[1],
         flops / ms[1] * 1e-9, d, e[0], e[1], cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const bool quick = argc > 1;
  const long n = 16384;
  double *A, *B, *C1, *C2, *dout;
  cudaMalloc(&A, n * n * 8); cudaMalloc(&B, n * n * 8); cudaMalloc(&C1, n * n * 8); cudaMalloc(&C2, n * n * 8); cudaMalloc(&dout, 8);
  fill<<<1024, 256>>>(A, n * n, 1);
  fill<<<1024, 256>>>(B, n * n, 2);
  cudaDeviceSynchronize();
  auto mk = [&](int M, int N, int K, int lower, int kmode, double beta) {
    GemmArgs g{};
    g.A = A; g.lda = n; g.strideA = 0; g.B = B; g.ldb = n; g.strideB = 0; g.C = C1; g.ldc = n; g.strideC = 0;
    g.M = M; g.N = N; g.K = K; g.alpha = -1.0; g.beta = beta; g.lower_only = lower; g.kmode = kmode; g.sel_block = 0;
    return g;
  };
  if (!quick) bench<false, false>("syrk rank-256 n=16128 (NT, lower, beta=1)", mk(16128, 16128, 256, 1, K_FULL, 1.0), 1, C1, C2, n * n, dout);
  if (!quick) bench<false, false>("syrk rank-256 n=8192", mk(8192, 8192, 256, 1, K_FULL, 1.0), 1, C1, C2, n * n, dout);
  if (!quick) bench<false, false>("syrk rank-256 n=2048", mk(2048, 2048, 256, 1, K_FULL, 1.0), 1, C1, C2, n * n, dout);
  bench<false, false>("gemm NT 8192^3 beta=0", mk(8192, 8192, 8192, 0, K_FULL, 0.0), 1, C1, C2, n * n, dout);
  if (!quick) bench<false, true>("gemm NN 8192^3 beta=0", mk(8192, 8192, 8192, 0, K_FULL, 0.0), 1, C1, C2, n * n, dout);
  if (!quick) bench<true, true>("lauum-like TN 16384 lower k>=m0", mk(16384, 16384, 16384, 1, K_GE_M0, 0.0), 1, C1, C2, n * n, dout);
  bench<false, true>("trtri-like NN 8192 k<m1", mk(8192, 8192, 8192, 0, K_LT_M1, 0.0), 1, C1, C2, n * n, dout);
  if (!quick) bench<false, true>("panel trsm 16256x128 K=128 (NN)", mk(16256, 128, 128, 0, K_FULL, 0.0), 1, C1, C2, n * n, dout);
  if (!quick) bench<false, true>("trsm update 16256x512 K=128 beta=1", mk(16256, 512, 128, 0, K_FULL, 1.0), 1, C1, C2, n * n, dout);
  return 0;
}
