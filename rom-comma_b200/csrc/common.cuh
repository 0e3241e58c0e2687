// Shared device/host helpers for the rom-comma B200 (sm_100a) FP64 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace rc {

constexpr int TILE = 128;   // every factorisation matrix is padded to a multiple of this (identity padding)

// Thread-local last error text for the C ABI (rc_last_error).
void set_error(const char* fmt, ...);
// Kernel-launch counter behind rc_launch_count(): bench.py's "gpu_launches" claim is counted here, not estimated.
void count_launches(long n);

// Optional per-launch timing of the DMMA GEMM (rc_profile_begin / rc_profile_end): CUDA events on the launching stream around every
// gemm_dmma_kernel launch plus the flops its tile list executes.  Off by default; a diagnostic, not thread-safe.
bool profile_enabled();
void profile_gemm_begin(cudaStream_t st);
void profile_gemm_end(cudaStream_t st, double flops);

#define RC_CUDA_OK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      rc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -1000 - (int)_e;                                                                     \
    }                                                                                             \
  } while (0)

#define RC_LAUNCH_OK()                \
  do {                                 \
    rc::count_launches(1);             \
    RC_CUDA_OK(cudaGetLastError());    \
  } while (0)

#define RC_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      rc::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// One native FP64 tensor-core op (SASS: DMMA.8x8x4): D(8x8) += A(8x4, row) * B(4x8, col).
// Lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction (fixed tree order); result valid in thread 0. `red` needs >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// exp(x) for the pairwise Gaussian-integrand kernels: branch-free, 14 FP64 + 6 integer instructions (libm's exp needs ~30 with its
// special-case handling; the Sobol sweep kernels are bound by exactly this instruction stream).  k = round(x / ln 2) by the 1.5 * 2^52
// trick, r = x - k ln 2 in two steps (Cody-Waite), exp(r) by a degree-10 near-minimax polynomial on |r| <= 0.3466,
// 2^k added into the exponent field.  Max relative error 1e-15.  Range: x below -708 is replaced by -708 with an INTEGER
// compare of its high word and two selects (an FP64 fmin/fmax pair costs two DSETP on the FP64 pipe and four selects per call): the
// result there is ~3e-308 instead of the true value below that - irrelevant to a sum of O(1) terms - however negative x gets (a line
// search can make a lengthscale tiny).  x > 709 is the CALLER's business: it cannot occur in the sweep kernels (single-input exponents are
// bounded by gamma (1-p) x^2 with |x| <= 7.04) nor in the gram / gradient kernels (arguments <= 0); the general-subset kernels, which sum up to
// M such exponents before the exp, pass fmin(e, 708).
// (Constants as literals on purpose: from a __constant__ table the register-resident loop gains 12 %, but the sweep kernels, which are
// out of registers, lose 7 % to the extra uniform-register loads.)
__device__ __forceinline__ double exp_pairwise(double x) {
  // high word in (hi(-708), hi(-inf)]: negative, |x| > 708, not a NaN (sign-magnitude order of the high word; NaNs propagate)
  if ((unsigned)__double2hiint(x) - 0xC0862001u <= 0xFFF00000u - 0xC0862001u) x = -708.0;
  const double magic = 6755399441055744.0;
  const double t = fma(x, 1.4426950408889634074, magic);
  const double k = t - magic;
  double r = fma(k, -6.93147180369123816490e-01, x);
  r = fma(k, -1.90821492927058770002e-10, r);
  double p = 2.74649892754840101e-07;             // degree-10 Chebyshev interpolant of exp on |r| <= 0.34662 (coefficients computed in
  p = fma(p, r, 2.76428065197613745e-06);         // 80-bit arithmetic, tools/exp_poly.py): max relative error 8e-16 - ten times tighter than
  p = fma(p, r, 2.48019467832707630e-05);         // the degree-11 Taylor polynomial it replaces, with one FMA less
  p = fma(p, r, 1.98411638732997871e-04);
  p = fma(p, r, 1.38888885130029937e-03);
  p = fma(p, r, 8.33333339078338141e-03);
  p = fma(p, r, 4.16666666681718770e-02);
  p = fma(p, r, 1.66666666665395619e-01);
  p = fma(p, r, 4.99999999999979683e-01);
  p = fma(p, r, 1.00000000000000799e+00);
  p = fma(p, r, 1.0);
  const int ki = __double2loint(t);
  return __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
}

// Table form of the same exp for the kernels that are bound by FP64 issue slots: 9 FP64 instructions instead of 14.
//   x = (32 e + j) ln2/32 + r,  |r| <= ln2/64:   exp(x) = 2^e * T[j] * (1 + r (1 + r g(r))),   T[j] = 2^(j/32) from a 256-byte table in
// SHARED memory (exp_table_fill; one 8-byte LDS per call - the table's 32 entries sit on 16 bank pairs, so a warp's lookup costs 2-4
// wavefronts), g the degree-3 Chebyshev interpolant of (e^r - 1 - r)/r^2 (tools/exp_poly.py table: max relative error of the polynomial
// 2.8e-16; with the rounding of T[j] and of the last FMA 5e-16 overall, tighter than the degree-10 form above).  The reduction is ONE FMA
// with ln2/32 rounded to double (relative error 3.3e-17): r is off by 3.3e-17 |x|, i.e. the result carries a relative error of that size -
// below one ulp for |x| < 3, and on terms of size e^x an absolute error <= 3.3e-17 |x| e^x <= 1.3e-17 otherwise.  Same range
// convention as exp_pairwise (x < -708 -> -708; x > 709 is the caller's business).  Measured alternatives: 16 entries (conflict-free by
// construction) with a degree-4 g, i.e. one more FMA: sweep 1.50 -> 1.56 ms, lattice blocks 37.3 -> 38.9 ms; the 32 entries stored 16 times so
// that every lane owns a bank pair: 1.50 -> 1.60 ms (one more address instruction).  FP64 instruction count decides, not the conflicts.
static __constant__ double rc_exp2_table[32] = {
    1.00000000000000000e+00, 1.02189714865411663e+00, 1.04427378242741375e+00, 1.06714040067682370e+00,
    1.09050773266525769e+00, 1.11438674259589243e+00, 1.13878863475669156e+00, 1.16372485877757748e+00,
    1.18920711500272103e+00, 1.21524735998046896e+00, 1.24185781207348400e+00, 1.26905095719173322e+00,
    1.29683955465100964e+00, 1.32523664315974132e+00, 1.35425554693689265e+00, 1.38390988196383202e+00,
    1.41421356237309515e+00, 1.44518080697704665e+00, 1.47682614593949935e+00, 1.50916442759342284e+00,
    1.54221082540794074e+00, 1.57598084510788650e+00, 1.61049033194925428e+00, 1.64575547815396495e+00,
    1.68179283050742900e+00, 1.71861929812247793e+00, 1.75625216037329945e+00, 1.79470907500310717e+00,
    1.83400808640934243e+00, 1.87416763411029996e+00, 1.91520656139714740e+00, 1.95714412417540018e+00,
};
// Every thread with threadIdx.x < 32 copies one entry; the caller's next __syncthreads() publishes the table.
__device__ __forceinline__ void exp_table_fill(double* tab) {
  if (threadIdx.x < 32) tab[threadIdx.x] = rc_exp2_table[threadIdx.x];
}
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  // x < -708 -> (-708.00002, -708]: ONE unsigned min on the high word (negative doubles order by magnitude there; positive values and
  // positive NaNs lie below the bound and pass; -inf and negative NaNs become -708, i.e. ~0)
  x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC0862000u), __double2loint(x));
  const double magic = 6755399441055744.0;
  const double t = fma(x, 4.61662413084468283841e+01, magic);      // 32 / ln 2
  const double k = t - magic;
  const double r = fma(k, -2.16608493924982901946e-02, x);         // ln 2 / 32
  const int ki = __double2loint(t);
  const double T = tab[ki & 31];
  double s = 8.33335685528634532e-03;
  s = fma(s, r, 4.16668302325693199e-02);
  s = fma(s, r, 1.66666666666312663e-01);
  s = fma(s, r, 4.99999999997592204e-01);
  s = fma(s, r, 1.0);
  const double p = fma(T * r, s, T);
  return __hiloint2double(__double2hiint(p) + (ki & ~31) * 32768, __double2loint(p));   // 2^(ki >> 5) into the exponent field
}

// Programmatic dependent launch: a kernel launched through launch_pdl may start (CTA scheduling, barrier / register set-up, descriptor
// prefetch) while its predecessor in the stream drains; it must execute pdl_wait() before its first access to global memory, which also
// makes the predecessor's writes visible.  pdl_launch_dependents() at the top of a kernel lets ITS successor do the same.  Kernels
// launched the ordinary way after one of these still wait for full completion.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember per (call site = kernel instantiation, device)
// that it has been raised to `bytes`, so that a process which works on several devices configures each of them (round-1 advice: a
// process-wide flag left every device but the first at the 48 KB default).  Idempotent, so the benign race between two threads is harmless.
constexpr int RC_MAX_DEVICES = 64;
#define RC_ENSURE_SMEM(kernel, bytes)                                                                              \
  do {                                                                                                             \
    static int _rc_have[rc::RC_MAX_DEVICES] = {};                                                                  \
    int _rc_dev = 0;                                                                                               \
    RC_CUDA_OK(cudaGetDevice(&_rc_dev));                                                                           \
    const bool _rc_tracked = _rc_dev >= 0 && _rc_dev < rc::RC_MAX_DEVICES;                                         \
    if (!_rc_tracked || _rc_have[_rc_dev] < (int)(bytes)) {                                                        \
      RC_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));         \
      if (_rc_tracked) _rc_have[_rc_dev] = (int)(bytes);                                                           \
    }                                                                                                              \
  } while (0)

// Number of SMs of the CURRENT device (cached per device): the size of every persistent grid.
inline int device_sm_count() {
  static int sms[RC_MAX_DEVICES] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev < 0 || dev >= RC_MAX_DEVICES) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }
  if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

}  // namespace rc
