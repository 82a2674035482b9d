// K4: fused sampler for the dense-contraction energies (full-covariance Gaussian, ProductOfT).
#pragma once
#include "common.cuh"
namespace mjhmc {
bool dense_supported(int dtype, int kind, int ndims, int nbasis);
cudaError_t launch_dense(int dtype, int kind, const LaunchParams& p, cudaStream_t stream);
// fp32 states on tcgen05 / TMEM / TMA (dense_tc.cu)
bool dense_tc_supported(int kind, int ndims, int nbasis);
cudaError_t launch_dense_tc(int kind, const LaunchParams& p, cudaStream_t stream);
long long dense_tc_workspace_bytes(int kind, int ndims);
cudaError_t dense_tc_prepare(int kind, const float* Mx, const float* nu, const float* b, int ndims, void* workspace,
                             cudaStream_t stream);
}
