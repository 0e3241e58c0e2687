// Live roofline denominators: register-resident DMMA.8x8x4 and FP64 exp loops (no memory traffic; the exp is exp_pairwise, the
// instruction stream the pairwise kernels themselves run - 14 FP64 instructions, against ~30 for libm's exp).  MEASURED_PEAKS.json has no
// FP64 entry, so bench.py measures these on the box it runs on and says so next to every fraction.
#include "../../include/romcomma_b200.h"
#include "common.cuh"
#include <algorithm>

namespace rc {

__global__ void dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void exp_peak_kernel(double* out, int iters) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = -1e-3 * (threadIdx.x + i + 1);
  double s = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s += exp_pairwise(x[i]);
      x[i] *= 1.0000001;
    }
  }
  if (s == 123.456) out[0] = s;
}

// The table form (exp_tab): the same loop, the table in shared memory as in the kernels that use it.
__global__ void exp_tab_peak_kernel(double* out, int iters) {
  __shared__ double etab[32];
  exp_table_fill(etab);
  __syncthreads();
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = -1e-3 * (threadIdx.x + i + 1);
  double s = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s += exp_tab(x[i], etab);
      x[i] *= 1.0000001;
    }
  }
  if (s == 123.456) out[0] = s;
}

// Test hook: y[i] = exp(x[i]) through either device form.
__global__ void exp_eval_kernel(const double* __restrict__ x, double* __restrict__ y, long n, int form) {
  __shared__ double etab[32];
  exp_table_fill(etab);
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = form == 0 ? exp_pairwise(x[i]) : exp_tab(x[i], etab);
}

template <typename Launch>
static int time_best(Launch launch, int reps, float* best_ms) {
  cudaEvent_t e0, e1;
  RC_CUDA_OK(cudaEventCreate(&e0));
  RC_CUDA_OK(cudaEventCreate(&e1));
  launch();
  RC_CUDA_OK(cudaDeviceSynchronize());
  *best_ms = 1e30f;
  for (int r = 0; r < reps; ++r) {
    RC_CUDA_OK(cudaEventRecord(e0));
    launch();
    RC_CUDA_OK(cudaEventRecord(e1));
    RC_CUDA_OK(cudaEventSynchronize(e1));
    float ms;
    RC_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < *best_ms) *best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

}  // namespace rc

using namespace rc;

extern "C" {

int rc_measure_dmma_tflops(double* scratch, double* tflops) {
  RC_REQUIRE(scratch && tflops, -2, "rc_measure_dmma_tflops: null pointer");
  int dev, sms;
  RC_CUDA_OK(cudaGetDevice(&dev));
  RC_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 20000, threads = 512;
  float ms;
  int rc = time_best([&] { dmma_peak_kernel<<<sms, threads>>>(scratch, iters); }, 5, &ms);
  if (rc) return rc;
  *tflops = (double)sms * (threads / 32) * iters * 8.0 * (8 * 8 * 4 * 2) / ms * 1e-9;
  return 0;
}

int rc_measure_exp_gexps(double* scratch, double* gexps) {
  RC_REQUIRE(scratch && gexps, -2, "rc_measure_exp_gexps: null pointer");
  int dev, sms;
  RC_CUDA_OK(cudaGetDevice(&dev));
  RC_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 2000, threads = 1024;
  float ms;
  int rc = time_best([&] { exp_peak_kernel<<<sms * 2, threads>>>(scratch, iters); }, 5, &ms);
  if (rc) return rc;
  *gexps = (double)sms * 2 * threads * iters * 4.0 / ms * 1e-6;
  return 0;
}

int rc_measure_exp_tab_gexps(double* scratch, double* gexps) {
  RC_REQUIRE(scratch && gexps, -2, "rc_measure_exp_tab_gexps: null pointer");
  int dev, sms;
  RC_CUDA_OK(cudaGetDevice(&dev));
  RC_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 2000, threads = 1024;
  float ms;
  int rc = time_best([&] { exp_tab_peak_kernel<<<sms * 2, threads>>>(scratch, iters); }, 5, &ms);
  if (rc) return rc;
  *gexps = (double)sms * 2 * threads * iters * 4.0 / ms * 1e-6;
  return 0;
}

int rc_debug_exp(const double* x, double* y, long n, int form, rc_stream_t stream) {
  RC_REQUIRE(x && y && n > 0 && (form == 0 || form == 1), -2, "rc_debug_exp: bad argument");
  exp_eval_kernel<<<(unsigned)std::min<long>((n + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(x, y, n, form);
  RC_LAUNCH_OK();
  return 0;
}

}  // extern "C"
