// DFMA latency / throughput on one SM (developer probe).  nvcc -arch=sm_100a -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, long long* cyc, int n, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int threads) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 64);
    const int n = 4096;
    chain<ILP><<<1, threads>>>(out, cyc, n, 0.999, 1e-3);
    chain<ILP><<<1, threads>>>(out, cyc, n, 0.999, 1e-3);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %4d (warps/SMSP %.1f) ILP %d: %.2f cycles per DFMA per warp-chain step, %.2f warp-DFMA/cycle/SM\n", threads,
           threads / 128.0, ILP, (double)h / n, (double)n * ILP * (threads / 32) / h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {32, 128, 256, 384, 512, 768, 1024}) { run<1>(th); run<2>(th); run<4>(th); run<8>(th); }
    return 0;
}
