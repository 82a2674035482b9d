// Streaming separable-energy sampler kernels (stream_separable.cuh): host entry points used by api.cu.
#pragma once
#include "common.cuh"

namespace mjhmc {
bool stream_supported(int dtype, int kind, int ndims);
cudaError_t launch_stream_kernel(int dtype, int kind, const LaunchParams& p, cudaStream_t stream);
void stream_set_tma(int enabled);
int stream_probe_blocks(long long smem);
void stream_last_launch(long long* out7);   // {use_tma, stages, grid, CTAs/SM, G, DT, smem bytes} of this thread's last launch     // 0 forces the cooperative-load path (tests)
}  // namespace mjhmc
