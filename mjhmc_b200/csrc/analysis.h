#pragma once
#include <cuda_runtime.h>
namespace mjhmc {
long long resample_scratch_bytes(long long m);
cudaError_t launch_resample(int dtype, int d, const double* dwell, long long m, const double* r, long long m_out,
                            const void* samples, long long ld_in, void* out, long long ld_out, long long* idx_out,
                            void* scratch, cudaStream_t s);
cudaError_t launch_autocorr(int dtype, int d, const void* samples, long long stride_k, long long stride_it,
                            long long n, int Tn, int n_lags, int circular, double* ac, cudaStream_t s);
// K7b: circular autocorrelation through batched forward FFTs (autocorr_fft.cu); T a power of two in [16, 4096]
bool autocorr_fft_supported(int Tn, int circular);
long long autocorr_fft_scratch_bytes(int Tn);
cudaError_t launch_autocorr_fft(int dtype, int d, const void* samples, long long stride_k, long long stride_it,
                                long long n, int Tn, int n_lags, double* ac, double* scratch, cudaStream_t s);
cudaError_t launch_ladder_visits(const unsigned char* choice, long long n_iter, long long n, int K, int* state,
                                 long long* visits, cudaStream_t s);
cudaError_t launch_moments(int dtype, const void* x, long long count, double* out, cudaStream_t s);
}
