// K4: fused sampler for the dense-contraction energies (full-covariance Gaussian, ProductOfT).
#pragma once
#include "common.cuh"
namespace mjhmc {
bool dense_supported(int dtype, int kind, int ndims, int nbasis);
cudaError_t launch_dense(int dtype, int kind, const LaunchParams& p, cudaStream_t stream);
// fp32 states on tcgen05 / TMEM / TMA (dense_tc.cu)
bool dense_tf32_supported(int kind, int ndims);
cudaError_t launch_dense_tf32(const LaunchParams& p, cudaStream_t stream);
long long dense_tf32_workspace_bytes(int ndims);
cudaError_t dense_tf32_prepare(const float* S, int ndims, float* workspace, cudaStream_t stream);
}
