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

// V[s][l][j] for s < nslices; masks are host-side bit sets over the M inputs (bit m set <=> input m is in the subset).
int sobol_contract(const double* X, int N, int M, const double* Phi, const double* c, int L, int Lp, const unsigned long long* masks, int nslices,
                   double* parts, double* V, cudaStream_t st);

}  // namespace rc
