// Fold data on the device: the probit normalisation of a fold (romcomma/data/storage.py:469-485, statistics :532-558) and the test metrics of
// GPR.test (romcomma/gpr/models.py:235-272).  O(N (M + L)) element-wise work around the hot path - bound by HBM and launch latency, nothing
// else - so that a fold's samples and predictions need not return to the host between the repository and the GP.
#include "../../include/romcomma_b200.h"
#include "common.cuh"
#include <algorithm>

namespace rc {

// One CTA per column of data (N, C) row-major: mean, then the centred sum of squares (two passes: the one-pass form cancels), std with
// ddof = 1 as pandas' DataFrame.std (storage.py:547).  stats rows (5, C): mean, std, rng = 2 sqrt(3) std, min = mean - sqrt(3) std, max.
__global__ void column_stats_kernel(const double* __restrict__ data, int N, int C, double* __restrict__ stats) {
  __shared__ double red[32];
  __shared__ double mean_s;
  const int c = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += data[(long)i * C + c];
  s = block_sum(s, red);
  if (threadIdx.x == 0) mean_s = s / (double)N;
  __syncthreads();
  const double mean = mean_s;
  double q = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double d = data[(long)i * C + c] - mean;
    q = fma(d, d, q);
  }
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    const double sd = sqrt(q / (double)(N - 1)), semi = sd * sqrt(3.0);
    stats[c] = mean;
    stats[C + c] = sd;
    stats[2 * C + c] = 2.0 * semi;
    stats[3 * C + c] = mean - semi;
    stats[4 * C + c] = mean + semi;
  }
}

// direction +1 (apply_to, storage.py:469-485): inputs (columns < M) -> probit( clip((x - min) / rng, margin, 1 - margin) ), outputs -> (y - mean) / std;
// direction -1 (undo_from, :487-503): inputs -> min + rng * Phi(x), outputs -> mean + std * y.
__global__ void normalize_kernel(const double* __restrict__ data, long N, int M, int C, const double* __restrict__ stats, double margin, int direction,
                                 double* __restrict__ out) {
  const long total = N * C;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const double v = data[e];
    double r;
    if (c < M) {
      const double lo = stats[3 * C + c], span = stats[2 * C + c];
      if (direction > 0) {
        const double u = fmin(fmax((v - lo) / span, margin), 1.0 - margin);
        r = normcdfinv(u);
      } else {
        r = lo + span * normcdf(v);
      }
    } else {
      const double mean = stats[c], sd = stats[C + c];
      r = direction > 0 ? (v - mean) / sd : mean + sd * v;
    }
    out[e] = r;
  }
}

// GPR.test (gpr/models.py:235-272): per test sample i and output l   abs error, z score, outlier flag (z^2 > 4), the any / all flags per sample,
// and the summary row  RMSE[l], mean SD[l], outlier fraction[l], any-fraction, all-fraction.  One CTA; n* is a fold's test set.
// reals (n, 2L): [ |truth - mean| (L) | z (L) ];  flags (n, L + 2) as 0/1 doubles;  summary (3L + 2).
__global__ void test_metrics_kernel(const double* __restrict__ truth, const double* __restrict__ mean, const double* __restrict__ sd, int n, int L,
                                    double* __restrict__ reals, double* __restrict__ flags, double* __restrict__ summary) {
  __shared__ double red[32];
  for (int l = 0; l <= L + 1; ++l) {          // l < L: output l;  l = L: any;  l = L + 1: all
    double se = 0.0, ssd = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (l < L) {
        const double err = truth[(long)i * L + l] - mean[(long)i * L + l], s = sd[(long)i * L + l], z = err / s;
        const double out = z * z > 4.0 ? 1.0 : 0.0;
        reals[(long)i * 2 * L + l] = fabs(err);
        reals[(long)i * 2 * L + L + l] = z;
        flags[(long)i * (L + 2) + l] = out;
        se = fma(err, err, se);
        ssd += s;
        cnt += out;
      } else {
        bool any = false, all = true;
        for (int k = 0; k < L; ++k) {
          const double z = (truth[(long)i * L + k] - mean[(long)i * L + k]) / sd[(long)i * L + k];
          const bool o = z * z > 4.0;
          any = any || o;
          all = all && o;
        }
        const double out = (l == L ? any : all) ? 1.0 : 0.0;
        flags[(long)i * (L + 2) + l] = out;
        cnt += out;
      }
    }
    se = block_sum(se, red);
    ssd = block_sum(ssd, red);
    cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) {
      if (l < L) {
        summary[l] = sqrt(se / (double)n);
        summary[L + l] = ssd / (double)n;
      }
      summary[2 * L + l] = cnt / (double)n;
    }
  }
}

}  // namespace rc

using namespace rc;

extern "C" {

int rc_column_stats(const double* data, int N, int C, double* stats, rc_stream_t stream) {
  RC_REQUIRE(data && stats && N >= 2 && C >= 1, -2, "rc_column_stats: null pointer, fewer than two rows or no column");
  column_stats_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(data, N, C, stats);
  RC_LAUNCH_OK();
  return 0;
}

int rc_normalize(const double* data, long N, int M, int C, const double* stats, double margin, int direction, double* out, rc_stream_t stream) {
  RC_REQUIRE(data && stats && out && N >= 1 && C >= 1 && M >= 0 && M <= C, -2, "rc_normalize: null pointer or bad shape");
  RC_REQUIRE(direction == 1 || direction == -1, -2, "rc_normalize: direction must be +1 (apply) or -1 (undo)");
  RC_REQUIRE(margin >= 0.0 && margin < 0.5, -2, "rc_normalize: margin %g out of [0, 0.5)", margin);
  const long total = N * C;
  const unsigned blocks = (unsigned)std::min<long>((total + 255) / 256, 8L * device_sm_count());
  normalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(data, N, M, C, stats, margin, direction, out);
  RC_LAUNCH_OK();
  return 0;
}

int rc_test_metrics(const double* truth, const double* mean, const double* sd, int n, int L, double* reals, double* flags, double* summary,
                    rc_stream_t stream) {
  RC_REQUIRE(truth && mean && sd && reals && flags && summary && n >= 1 && L >= 1, -2, "rc_test_metrics: null pointer or empty test set");
  test_metrics_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(truth, mean, sd, n, L, reals, flags, summary);
  RC_LAUNCH_OK();
  return 0;
}

}  // extern "C"
