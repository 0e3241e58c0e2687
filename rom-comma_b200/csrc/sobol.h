// Internal (C++) interface of the Sobol kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace rc {

constexpr int SOBOL_MAX_SLICES = 64;   // subsets per launch (longer lists are chunked)

// Phi [P][M], g0 [P][N], g0KY [P][N] (centred), P = L (diagonal F) or L*L (full F); F is (L) or (L,L); KinvY is (L,N).
int sobol_prepare(const double* X, int N, int M, const double* Lam, const double* F, const double* KinvY, int L, int is_F_diagonal, double* Phi,
                  double* g0, double* g0KY, cudaStream_t st);

size_t sobol_workspace_bytes(int N, int P, int ns);

// Position of a structured subset (single input, prefix [0:k], suffix [k:M], full, empty) in the 3M outputs of the sweep-form kernels
// { F[m] m<M | P[k] k=1..M | S[k] k=1..M-1 | E }, or -1 for a general subset.
int sobol_sweep_index(unsigned long long mask, int M);

// V[s][l][j] for s < nslices; masks are host-side bit sets over the M inputs (bit m set <=> input m is in the subset).
// part/nparts: only the 64-row tiles ti with ti % nparts == part are evaluated (V is then a partial sum; 0/1 = everything).
int sobol_contract(const double* X, int N, int M, const double* Phi, const double* c, int L, int Lp, const unsigned long long* masks, int nslices,
                   double* parts, double* V, int part, int nparts, cudaStream_t st);

// ClosedSobolWithError (diagonal F): V[s][l][i] and W[s][l][i] = (mu_phi_mu - mu_psi_mu) + transpose for every subset; with WMm != NULL
// (is_T_partial = False) also the MIXED covariances WMm[s][l][i] between the full model and the marginal s.
// Phi (L,M), g0 (L,N), g0KY (L,N) come from sobol_prepare; Achol/dinv from potrf_lower with chol_batch = 1 (covariant GP, n_pad >= L*N)
// or L (variant GP, one N x N factor per output).
size_t sobol_error_workspace_bytes(int N, int M, int L, int nslices, int n_pad, int chol_batch, int mixed);
int sobol_error(const double* X, int N, int M, const double* Lam, const double* F, const double* Phi, const double* g0, const double* g0KY, int L,
                const double* Achol, int n_pad, long ld, long strideA, int chol_batch, const double* dinv, const unsigned long long* masks,
                int nslices, void* work, double* V, double* W, double* WMm, cudaStream_t st);

}  // namespace rc
