// Developer probe: one tcgen05.mma kind::tf32 (M=128, N=16, K=8) with index-filled operands, to
// pin down the shared-memory descriptor semantics on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    uint64_t d = 0; d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)lt << 61; return d;
}
struct Args { int mode; uint32_t a_lbo, a_sbo, a_lt, b_lbo, b_sbo, b_lt, idesc; };
__global__ void probe(Args a, float* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    float* A = (float*)base;              // 16 KB
    float* B = (float*)(base + 16384);    // 16 KB
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4096; i += 128) {
        A[i] = (a.mode == 1) ? 1.0f : (float)i;
        B[i] = (a.mode == 2) ? 1.0f : (float)i;
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        uint64_t ad = make_desc(smem_u32(A), a.a_lbo, a.a_sbo, a.a_lt), bd = make_desc(smem_u32(B), a.b_lbo, a.b_sbo, a.b_lt);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tm), "l"(ad), "l"(bd), "r"(a.idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\ntcgen05.wait::ld.sync.aligned;"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                   "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16)) : "memory");
    for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(u[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}
int main(int argc, char** argv) {
    Args a;
    a.mode = atoi(argv[1]);
    a.a_lbo = atoi(argv[2]); a.a_sbo = atoi(argv[3]); a.a_lt = atoi(argv[4]);
    a.b_lbo = atoi(argv[5]); a.b_sbo = atoi(argv[6]); a.b_lt = atoi(argv[7]);
    int a_major = atoi(argv[8]), b_major = atoi(argv[9]);
    a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    float* out; cudaMalloc(&out, 128 * 16 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    probe<<<1, 128, 40000>>>(a, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d aL %u aS %u alt %u bL %u bS %u blt %u amaj %d bmaj %d -> %s\n", a.mode, a.a_lbo, a.a_sbo, a.a_lt, a.b_lbo, a.b_sbo, a.b_lt, a_major, b_major, cudaGetErrorString(e));
    float h[128 * 16]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    int rows[] = {0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 127};
    for (int r : rows) { printf("m=%3d:", r); for (int j = 0; j < 16; ++j) printf(" %8.0f", h[r * 16 + j]); printf("\n"); }
    return 0;
}
