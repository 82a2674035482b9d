// K4: fused sampler for the dense-contraction energies (full-covariance Gaussian, ProductOfT).
#pragma once
#include "common.cuh"
namespace mjhmc {
bool dense_supported(int dtype, int kind, int ndims, int nbasis);
cudaError_t launch_dense(int dtype, int kind, const LaunchParams& p, cudaStream_t stream);
}
