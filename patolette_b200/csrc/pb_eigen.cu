// pb_eigen.cu - row N2: the 3 x 3 symmetric eigen solve on the device.  One thread per matrix runs pb_dsyev3.h, the
// same operation-for-operation restatement of LAPACK's dsyev('V', 'L', 3) the host side uses (lib/src/math/eigen.c:
// 83-140, math/pca.c:122-149): --fmad=false keeps every product and sum a separate rounding, the fused ones of the
// BLAS kernels are explicit __fma_rn, sqrt and division are IEEE.  A batch of covariances becomes a batch of principal
// axes without leaving the chip.
#include "pb_common.cuh"
#include "pb_dsyev3.h"
#include "pb_kernels.h"
#include "pb_prof.h"

namespace {

__global__ void __launch_bounds__(64) k_eigen3(const double *__restrict__ a9, int n, double *__restrict__ w3, double *__restrict__ z9,
                                               int *__restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[9], w[3];
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = a9[(size_t)9 * i + k];
    const int rc = pb_eig::dsyev3(a, w);
#pragma unroll
    for (int k = 0; k < 9; k++) z9[(size_t)9 * i + k] = a[k];
#pragma unroll
    for (int k = 0; k < 3; k++) w3[(size_t)3 * i + k] = w[k];
    if (info) info[i] = rc;
}

} // namespace

// a9: n matrices, 3 x 3 column-major, lower triangle significant (device).  w3: n x 3 ascending eigenvalues,
// z9: n x 9 eigenvectors in columns (the principal axis is z9[6..8], pca.c:136-138), info: LAPACK's info or nullptr.
void pb_launch_eigen3(const double *a9, int n, double *w3, double *z9, int *info, cudaStream_t st) {
    if (n <= 0) return;
    PbProfScope p("k_eigen3", st);
    k_eigen3<<<(n + 63) / 64, 64, 0, st>>>(a9, n, w3, z9, info);
    PB_CUDA_OK(cudaGetLastError());
}
