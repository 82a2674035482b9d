// Developer probe 5 (GPU box): tcgen05.mma kind::f16 with the A operand in TENSOR MEMORY (bf16 pairs packed in 32-bit
// columns, lane = row), B K-major no-swizzle in shared memory as in dense_tc.cu.  D[m][n] = sum_k A[m][k] * B[n][k].
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
constexpr int M = 128, N = 112, K = 112, KS = 7, KC = K / 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0; d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d;
}
__global__ void probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* out, long long* cyc, int reps) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    __nv_bfloat16* Bs = (__nv_bfloat16*)base;                       // N/8 row groups x KC cores x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        Bs[((n >> 3) * (KC * 128) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) / 2] = __float2bfloat16(Bg[i]);
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    const uint32_t a_col = 256;
    // my row of A -> TMEM: 8 columns (16 bf16) per K step
    for (int kg = 0; kg < KS; ++kg) {
        uint32_t u[8];
        for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(Ag[tid * K + kg * 16 + 2 * j], Ag[tid * K + kg * 16 + 2 * j + 1]);
            u[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(tm + ((uint32_t)(warp * 32) << 16) + a_col + kg * 8), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        for (int rep = 0; rep < reps; ++rep) {
            const long long t0 = clock64();
            for (int kg = 0; kg < KS; ++kg) {
                const uint64_t bd = make_desc(smem_u32(Bs) + kg * 256, 128, KC * 128);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(tm), "r"(tm + a_col + kg * 8), "l"(bd), "r"(idesc), "r"(kg > 0 ? 1u : 0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
            cyc[rep] = clock64() - t0;
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 112; c += 16) {
        uint32_t u[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\ntcgen05.wait::ld.sync.aligned;"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                       "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + c) : "memory");
        for (int j = 0; j < 16; ++j) out[tid * 112 + c + j] = __uint_as_float(u[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
int main() {
    std::vector<float> A(M * K), B(N * K), D(M * 112);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : B) v = (float)(rand() % 5 - 2);
    float *dA, *dB, *dO; long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, D.size() * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (N / 8) * KC * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dO, dC, 3);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dO, D.size() * 4, cudaMemcpyDeviceToHost);
    long long h[3]; cudaMemcpy(h, dC, sizeof h, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            const double err = fabs(ref - D[m * 112 + n]);
            if (err > worst) worst = err;
            if (err > 0.5) bad++;
        }
    if (bad) for (int m = 0; m < 2; ++m) {
        printf("  got  m=%d:", m); for (int n = 0; n < 16; ++n) printf(" %4.0f", D[m * 112 + n]); printf("\n");
        printf("  want m=%d:", m); for (int n = 0; n < 16; ++n) { double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k]; printf(" %4.0f", ref); } printf("\n");
    }
    printf("A in TMEM (bf16): %s, max abs err %g, mismatches %d; 7 MMAs: %lld cycles (%.1f / MMA)\n", cudaGetErrorString(e), worst, bad, h[2], h[2] / 7.0);
    return 0;
}
