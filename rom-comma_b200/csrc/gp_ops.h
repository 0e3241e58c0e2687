// Internal (C++) interface of the gram / gradient / prediction kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace rc {

struct GramArgs {
  const double* X;  int N;      // (N, M) row-major inputs of the rows
  const double* X2; int N2;     // (N2, M) inputs of the columns (== X for the training gram)
  int M, L;
  const double* ls; long stride_ls;    // (L, M) lengthscales, batch stride
  const double* F;  const double* E;   // (L, L) each or nullptr (F: unit variance, E: no noise); batch stride stride_FE
  long stride_FE;
  double* out; long ld_out; long stride_out;
  int rows_pad, cols_pad;              // multiples of 64; indices >= L*N (L*N2) are padding
  int lower_only;                      // only 64-tiles on or below the diagonal
  int pad_identity;                    // padding gets the identity (square factorisation input) instead of zeros
  int strip;                           // set by gram(): 64-column tiles per CTA (1 for small problems, 4 once the grid exceeds a few waves)
  long stride_X;                       // problems with their OWN inputs (folds): problem z reads X + z*stride_X (X2 likewise); 0 = shared
  const int* Nz;                       // device array of per-problem sample counts (<= N) or nullptr = N for all; rows >= L*Nz[z] are padding
};
int gram(const GramArgs& a, int batch, cudaStream_t st);

int apply_variance_noise(const double* Ku, long ldu, const double* F, const double* E, int L, int N, int n_pad, double* out, long ldo,
                         int lower_only, cudaStream_t st);

struct GradArgs {
  const double* X; int N, M, L;
  const double* ls; long stride_ls;
  const double* F;  long stride_FE;
  const double* Kinv; long ldk; long stride_K;   // lower triangle of K^-1, padded storage
  const double* alpha; long stride_alpha;        // K^-1 y, padded with zeros
  double* parts;                                 // grad_workspace_bytes
  int with_ls;
  int diag_blocks_only;                          // Kinv holds only the tiles that intersect the diagonal (l,l) blocks: restrict the sums to l_i == l_j
  int nvals, slots;                              // filled in by grad_reduce
  long stride_X;                                 // per-problem inputs (folds): X + z*stride_X; 0 = shared
  const int* Nz;                                 // device array of per-problem sample counts or nullptr
};
int grad_nvals(int L, int M);   // layout: SF (L*L), SE (L*L), dls row part (L*M), dls column part (L*M)
size_t grad_workspace_bytes(int n_pad, int L, int M, int batch);
int grad_reduce(GradArgs a, int n_pad, int batch, double* out, cudaStream_t st);

size_t predict_workspace_bytes(int c_pad, int batch);
int predict_reduce(const double* A, long lda, long strideA, const double* a, long stride_a, int n, int c_pad, int batch, int L, int nstar,
                   const double* kdiag, const double* noise, double* parts, double* mean, double* var, cudaStream_t st);

// The gradient GP of a variant model (gpr/models.py:386-415): Jacobian of k(X, x) into the zero-padded right-hand sides B (batch, n_pad, c_pad),
// c = j*M + m, plus the mean; and the final assembly of the (o, o, batch, M, M) covariance from C = -W^T W.
int predict_gradient_jacobian(const double* X, int N, int M, const double* xs, int o, const double* ls, const double* variance, const double* KiY,
                              int batch, double* B, long ldb, long strideB, double* mean, cudaStream_t st);
int predict_gradient_finish(const double* C, long ldc, long strideC, const double* xs, int o, int M, const double* ls, const double* variance, int batch,
                            double* var, cudaStream_t st);

}  // namespace rc
