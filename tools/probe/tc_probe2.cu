// Developer probe 2 (GPU box): tcgen05.mma kind::tf32 with the operand layouts the ProductOfT kernel needs.
//   test 0: D[m][n] = sum_k A[m][k] * Bk[n][k]   A K-major, B K-major   (the layout dense_tc.cu already uses)
//   test 1: D[m][j] = sum_k A[m][k] * W[k][j]    A K-major, B = the SAME buffer as test 0 read MN-major
// The W buffer holds W[dim][expert] tiled as 8-dim x 4-expert core matrices (16-byte rows of 4 experts):
//   element (dim, expert) at (dim/8)*(KC*128) + (expert/4)*128 + (dim%8)*16 + (expert%4)*4   bytes
// K-major view  (N = dim,    K = expert): LBO = 128, SBO = KC*128, start + kstep*256
// MN-major view (N = expert, K = dim)   : SBO = 128, LBO = KC*128, start + kstep*(KC*128), idesc b_major = 1
// Operands are small integers (exact in tf32), so the result must be exact.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
constexpr int M = 128, ND = 112, NE = 112, KS = 13, KC = NE / 4;      // dims 112 (N of test 0), experts 112
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
    float* A = (float*)base;                       // 26 core columns x 2048 B
    float* W = (float*)(base + 26 * 2048);         // 14 dim groups x KC x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * 104; i += 128) {     // A[m][k] K-major: (k/4)*2048 + (m/8)*128 + (m%8)*16 + (k%4)*4
        const int m = i / 104, k = i % 104;
        A[((k >> 2) * 2048 + (m >> 3) * 128 + (m & 7) * 16 + (k & 3) * 4) / 4] = Ag[i];
    }
    for (int i = tid; i < ND * NE; i += 128) {     // W[dim][expert]
        const int dm = i / NE, e = i % NE;
        W[((dm >> 3) * (KC * 128) + (e >> 2) * 128 + (dm & 7) * 16 + (e & 3) * 4) / 4] = Wg[i];
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
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((test == 1 ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        for (int kg = 0; kg < KS; ++kg) {
            const uint64_t ad = make_desc(smem_u32(A) + kg * 2 * 2048, 2048, 128);
            uint64_t bd;
            if (test == 0) bd = make_desc(smem_u32(W) + kg * 256, 128, KC * 128);          // K = experts 8kg..8kg+7
            else if (variant == 0) bd = make_desc(smem_u32(W) + kg * (KC * 128), KC * 128, 128);   // LBO = K-group stride, SBO = MN stride
            else if (variant == 1) bd = make_desc(smem_u32(W) + kg * (KC * 128), 128, KC * 128);   // swapped
            else if (variant == 2) bd = make_desc(smem_u32(W) + kg * (KC * 128), 128, 128);
            else bd = make_desc(smem_u32(W) + kg * (KC * 128), 16, 128);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
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
    std::vector<float> A(M * 104), W(ND * NE), D(M * 112);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : W) v = (float)(rand() % 5 - 2);
    float *dA, *dW, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 26 * 2048 + 14 * KC * 128 + 2048;
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
                for (int k = 0; k < 104; ++k)
                    ref += test == 0 ? (double)A[m * 104 + k] * W[n * NE + k]       // D[m][dim n]    = sum_expert A[m][e] W[n][e]
                                     : (double)A[m * 104 + k] * W[k * NE + n];      // D[m][expert n] = sum_dim    A[m][d] W[d][n]
                const double err = fabs(ref - D[m * 112 + n]);
                if (err > worst) worst = err;
                if (err > 0.5 && bad++ < 5) printf("  test %d mismatch m=%d n=%d got %g want %g\n", test, m, n, D[m * 112 + n], ref);
            }
        if (bad) {
            for (int m = 0; m < 2; ++m) {
                printf("  got  m=%d:", m); for (int n = 0; n < 24; ++n) printf(" %4.0f", D[m * 112 + n]); printf("\n");
                printf("  want m=%d:", m);
                for (int n = 0; n < 24; ++n) { double ref = 0; for (int k = 0; k < 104; ++k) ref += test == 0 ? (double)A[m * 104 + k] * W[n * NE + k] : (double)A[m * 104 + k] * W[k * NE + n]; printf(" %4.0f", ref); }
                printf("\n");
            }
        }
        printf("variant %d test %d (%s): %s, max abs err %g, mismatches %d\n", variant, test, test == 0 ? "B K-major" : "B MN-major (same buffer)",
               cudaGetErrorString(e), worst, bad);
    }
    return 0;
}
