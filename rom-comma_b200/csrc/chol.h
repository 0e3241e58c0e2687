// Internal (C++) interface of the dense FP64 factorisation layer; the C ABI in capi.cu is built on these.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace rc {

size_t potrf_workspace_bytes(int n, int batch);

// In-place lower Cholesky of `batch` n x n matrices (only the lower triangle is read; the strict upper triangle of
// diagonal 128-tiles is clobbered). dinv: batch*(n/128) inverse diagonal blocks; logdet_parts: batch*(n/128) partial
// sums of log L_ii; info[z] = 0 or the 1-based index of the first non-positive pivot.
// lookahead: a single large matrix may be factorised with the serial diagonal/panel chain on an internal high-priority stream underneath
// the trailing updates (chol.cu: potrf_lookahead; same bits); false keeps every kernel on `st` (per-kernel timing).
int potrf_lower(double* A, int n, long ld, long strideA, int batch, double* dinv, double* logdet_parts, int* info, cudaStream_t st,
                bool lookahead = true);

// x = L^-1 w (transpose=0) or L^-T w (transpose=1); w is destroyed. Vectors of batch z start at z*strideV.
int trsv_lower(const double* A, int n, long ld, long strideA, int batch, const double* dinv, double* w, double* x, long strideV, int transpose,
               cudaStream_t st);

// B <- L^-1 B for an n x nrhs block of right-hand sides (nrhs multiple of 128).
int trsm_lower_fwd(const double* A, int n, long ld, long strideA, int batch, const double* dinv, double* B, int nrhs, long ldb, long strideB,
                   cudaStream_t st);

// The same for few right-hand sides against one large factor: the diagonal super-blocks of 1024 rows are inverted once (prepare; work =
// trsm_sbinv_workspace_doubles(n, nrhs_max) doubles, the last 1024 * nrhs_max of them the product buffer Tbuf), then every super-block is one
// triangular product and one rank-1024 update.
size_t trsm_sbinv_workspace_doubles(int n, int nrhs);
int trsm_sbinv_prepare(const double* A, int n, long ld, const double* dinv, double* work, cudaStream_t st);
int trsm_lower_fwd_sbinv(const double* A, int n, long ld, const double* dinv, const double* work, double* Tbuf, double* B, int nrhs, long ldb,
                         cudaStream_t st);

// A (holding L) <- L^-1 in place; tmp needs n*n/4 doubles per matrix.
// rest_from > 0: the leading rest_from x rest_from block is inverted already (a call with n = rest_from and strideD_blocks = blocks of the whole
// matrix); strideD_blocks: blocks per matrix in dinv (0 = n / 128); tiles_per_cta > 0: yielding launches.
int trtri_lower(double* A, int n, long ld, long strideA, int batch, const double* dinv, double* tmp, long strideT, cudaStream_t st, int rest_from = 0,
                int strideD_blocks = 0, int tiles_per_cta = 0);

// potrf_lower + trtri_lower of ONE matrix with the independent part of the inverse running on a side stream inside the factorisation's
// idle phases (chol.cu); on return (in stream order) A holds Z = L^-1, dinv the block inverses, logdet_parts / info as potrf_lower.
// tmp: potrf_trtri_tmp_doubles(n, panels) doubles; panels < 2 (or a small n) selects the plain sequence.
size_t potrf_trtri_tmp_doubles(int n, int panels);
int potrf_trtri_lower(double* A, int n, long ld, double* dinv, double* logdet_parts, int* info, double* tmp, size_t tmp_doubles, int panels,
                      cudaStream_t st, bool lookahead = true);

// Kinv(lower 128-tiles) = Z^T Z; sel_block > 0 restricts it to the tiles that intersect the diagonal blocks of that size.
int lauum_lower(const double* Z, int n, long ld, long strideZ, int batch, double* Kinv, long ldk, long strideK, int sel_block, cudaStream_t st);

// C = alpha * A^T A + beta * C (A: n x c row-major, n and c multiples of 128; all tiles of the c x c result).
int syrk_tn(const double* A, int n, int c, long lda, long strideA, int batch, double alpha, double beta, double* C, long ldc, long strideC, cudaStream_t st);

// dots[pair(l > l')][i] = K^-1[(l,i),(l',i)] from Z = L^-1 (pair index l*(l-1)/2 + l').
size_t block_diag_dots_workspace_bytes(int N, int L);
int block_diag_dots(const double* Z, long ld, int n, int N, int L, double* parts, double* dots, cudaStream_t st);

// out = Z v (transpose = 0) or Z^T v (transpose = 1) for lower-triangular Z (only i >= j is read); parts: tri_gemv_workspace_bytes.
size_t tri_gemv_workspace_bytes(int n, int batch);
int tri_gemv_lower(const double* Z, int n, long ld, long strideZ, int batch, const double* v, double* out, long strideV, int transpose,
                   double* parts, cudaStream_t st);

// Host-side: the order in which a tile list is handed out (m0, n0, k begin, k steps of 16 per tile); returns the tile count.
int debug_tile_order(int M, int N, int K, int lower_only, int kmode, int sel_block, int* out);

int sum_parts(const double* parts, int count, int batch, double* out, double scale, cudaStream_t st);
int dot_batched(const double* a, const double* b, long n, long stride, int batch, double* out, cudaStream_t st);
int extract_lower(const double* src, long lds, long strideS, double* dst, int n, long strideDst, int batch, int symmetrize, cudaStream_t st);
int pad_identity(const double* src, int n, long strideS, double* dst, int n_pad, long ldd, long strideD, int batch, cudaStream_t st);

}  // namespace rc
