// Streaming separable-energy sampler kernels (stream_separable.cuh): host entry points used by api.cu.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace mjhmc {
bool stream_supported(int dtype, int kind, int ndims);
cudaError_t launch_stream_kernel(int dtype, int kind, const LaunchParams& p, cudaStream_t stream);
void stream_set_tma(int enabled);
bool make_tensor_map_f32(CUtensorMap* m, const void* base, int rank, const long long* dims, const long long* strides,
                         const int* box);
int stream_probe_blocks(long long smem);
void stream_last_launch(long long* out7);   // {use_tma, stages, grid, CTAs/SM, G, DT, smem bytes} of this thread's last launch     // 0 forces the cooperative-load path (tests)
}  // namespace mjhmc
