// Probe: which way of handing a 2-D fp32 tensor map (box 32 x 100) to cp.async.bulk.tensor works on this GPU.
// usage: tma_probe <variant>   0: direct __grid_constant__ param, L2 promo 256B   1: struct member   2: struct member
// through a lambda with a run-time select   3: direct, promo NONE   4: direct, promo 128B
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>

struct Maps { CUtensorMap a, b; int ok; };

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void box_g2s(void* dst, const CUtensorMap* map, int c0, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(su32(dst)), "l"((uint64_t)map), "r"(c0), "r"(0), "r"(su32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(su32(bar)) : "memory");
}
template <int V>
__global__ void k(const __grid_constant__ CUtensorMap m, const __grid_constant__ Maps ms, float* out, int sel, int c0) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](bool first) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(32 * 100 * 4) : "memory");
            if (V == 0 || V >= 3) box_g2s(sm, &m, c0, &bar);
            if (V == 1) box_g2s(sm, &ms.a, c0, &bar);
            if (V == 2) { const CUtensorMap* mp = first ? &ms.a : &ms.b; box_g2s(sm, mp, c0, &bar); }
        }
    };
    issue(sel != 0);
    wait(&bar);
    for (int i = threadIdx.x; i < 3200; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    const long long n = 4096, d = 100;
    const int c0 = argc > 2 ? atoi(argv[2]) : 64;
    float* X; float* out;
    cudaMalloc(&X, n * d * 4); cudaMalloc(&out, 3200 * 4);
    float* h = (float*)malloc(n * d * 4);
    for (long long i = 0; i < n * d; ++i) h[i] = (float)(i % 1000);
    cudaMemcpy(X, h, n * d * 4, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled)f;
    CUtensorMap m; Maps ms; memset(&ms, 0, sizeof ms);
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)d}; cuuint64_t gs[1] = {(cuuint64_t)n * 4};
    cuuint32_t box[2] = {32, 100}, es[2] = {1, 1};
    CUtensorMapL2promotion promo = v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 4 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode %d\n", v, (int)r);
    ms.a = m; ms.b = m; ms.ok = 1;
    const size_t smem = 32 * 100 * 4 + 1024;
    cudaError_t e = cudaSuccess;
    if (v == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<1, 128, smem>>>(m, ms, out, 1, c0); }
    else if (v == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<2><<<1, 128, smem>>>(m, ms, out, 1, c0); }
    else { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<1, 128, smem>>>(m, ms, out, 1, c0); }
    e = cudaDeviceSynchronize();
    float o[4] = {0, 0, 0, 0};
    cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
    printf("variant %d: %s  out[0..2] = %g %g %g (first column %d)\n", v, cudaGetErrorString(e), o[0], o[1], o[2], c0);
    return 0;
}
