// Developer probe 3 (GPU box): tcgen05.mma kind::f16 with bf16 operands -- can ONE no-swizzle buffer of W be read
// K-major (N = dim, K = expert) by one MMA and MN-major (N = expert, K = dim) by another?
// W[dim][expert] tiled as 8-dim x 8-expert core matrices (16-byte rows of 8 bf16 experts):
//   element (dim, expert) at (dim/8)*(KC*128) + (expert/8)*128 + (dim%8)*16 + (expert%8)*2   bytes, KC = NE/8
// K-major view : LBO = 128 (next 8 experts), SBO = KC*128 (next 8 dims), start + kstep*256 (K = 16 experts per MMA)
// MN-major view: SBO = 128 (next 8 experts along N), LBO = KC*128 (next 8 dims along K), start + kstep*2*KC*128
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
constexpr int M = 128, ND = 112, NE = 112, KS = 7, KC = NE / 8, AK = 112;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0; d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d;
}
__global__ void probe(int test, int variant, const float* __restrict__ Ag, const float* __restrict__ Wg, float* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    __nv_bfloat16* A = (__nv_bfloat16*)base;                       // 14 core columns (8 k each) x 2048 B
    __nv_bfloat16* W = (__nv_bfloat16*)(base + 14 * 2048);         // 14 dim groups x KC x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * AK; i += 128) {      // A[m][k] K-major: (k/8)*2048 + (m/8)*128 + (m%8)*16 + (k%8)*2
        const int m = i / AK, k = i % AK;
        A[((k >> 3) * 2048 + (m >> 3) * 128 + (m & 7) * 16 + (k & 7) * 2) / 2] = __float2bfloat16(Ag[i]);
    }
    for (int i = tid; i < ND * NE; i += 128) {
        const int dm = i / NE, e = i % NE;
        W[((dm >> 3) * (KC * 128) + (e >> 3) * 128 + (dm & 7) * 16 + (e & 7) * 2) / 2] = __float2bfloat16(Wg[i]);
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        const uint32_t N = 112;
        // D = f32 (1<<4), A = B = bf16 (1 << 7, 1 << 10)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((test == 1 ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        for (int kg = 0; kg < KS; ++kg) {
            const uint64_t ad = make_desc(smem_u32(A) + kg * 2 * 2048, 2048, 128);
            uint64_t bd;
            if (test == 0) bd = make_desc(smem_u32(W) + kg * 256, 128, KC * 128);                          // K = experts 16kg..
            else if (variant == 0) bd = make_desc(smem_u32(W) + kg * 2 * (KC * 128), KC * 128, 128);      // K = dims 16kg..
            else bd = make_desc(smem_u32(W) + kg * 2 * (KC * 128), 128, KC * 128);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(kg > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<float> A(M * AK), W(ND * NE), D(M * 112);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : W) v = (float)(rand() % 5 - 2);
    float *dA, *dW, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 14 * 2048 + 14 * KC * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int test = 0; test < 2; ++test) {
        cudaMemset(dO, 0, D.size() * 4);
        probe<<<1, 128, smem>>>(test, variant, dA, dW, dO);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dO, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0; int bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < 112; ++n) {
                double ref = 0;
                for (int k = 0; k < AK; ++k)
                    ref += test == 0 ? (double)A[m * AK + k] * W[n * NE + k] : (double)A[m * AK + k] * W[k * NE + n];
                const double err = fabs(ref - D[m * 112 + n]);
                if (err > worst) worst = err;
                if (err > 0.5) bad++;
            }
        if (bad) {
            for (int m = 0; m < 2; ++m) {
                printf("  got  m=%d:", m); for (int n = 0; n < 24; ++n) printf(" %4.0f", D[m * 112 + n]); printf("\n");
                printf("  want m=%d:", m);
                for (int n = 0; n < 24; ++n) { double ref = 0; for (int k = 0; k < AK; ++k) ref += test == 0 ? (double)A[m * AK + k] * W[n * NE + k] : (double)A[m * AK + k] * W[k * NE + n]; printf(" %4.0f", ref); }
                printf("\n");
            }
        }
        printf("bf16 variant %d test %d (%s): %s, max abs err %g, mismatches %d\n", variant, test, test == 0 ? "B K-major" : "B MN-major (same buffer)",
               cudaGetErrorString(e), worst, bad);
    }
    return 0;
}
