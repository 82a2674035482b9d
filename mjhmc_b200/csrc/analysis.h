#pragma once
#include <cuda_runtime.h>
namespace mjhmc {
long long resample_scratch_bytes(long long m);
cudaError_t launch_resample(int dtype, int d, const double* dwell, long long m, const double* r, long long m_out,
                            const void* samples, long long ld_in, void* out, long long ld_out, long long* idx_out,
                            void* scratch, cudaStream_t s);
cudaError_t launch_autocorr(int dtype, int d, const void* samples, long long stride_k, long long stride_it,
                            long long n, int Tn, int n_lags, int circular, double* ac, cudaStream_t s);
cudaError_t launch_moments(int dtype, const void* x, long long count, double* out, cudaStream_t s);
}
