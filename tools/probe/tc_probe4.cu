// Developer probe 4 (GPU box): throughput of back-to-back tcgen05.mma kind::f16 (bf16, M = 128, K = 16, no-swizzle
// K-major operands laid out like dense_tc.cu) issued by one thread, timed with clock64 around issue .. commit wait.
//   argv: N  n_mma  n_acc (accumulators cycled)  mode (bit0: the 6-term plane pattern, bit1: B read MN-major)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt = 0) {
    uint64_t d = 0; d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)lt << 61; return d;
}
__global__ void probe(int N, int n_mma, int n_acc, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 40000; i += blockDim.x) ((uint32_t*)base)[i] = 0x3f803f80u;   // bf16 1.0
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
    const uint32_t ncores = 14;
    __shared__ volatile int s_stop;
    if (tid == 0) s_stop = 0;
    __syncthreads();
    if ((mode & 8) && warp == 0) {
        // all lanes converged, one elected lane issues: descriptors stay in uniform registers (dense_tc.cu)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(base), b0 = smem_u32(base) + 3 * 14 * 2048;
        const uint64_t a_hi = make_desc(0, 2048, 128), b_hi = make_desc(0, 128, ncores * 128);
        uint32_t a_lo[3], b_lo[3];
        for (int pl = 0; pl < 3; ++pl) { a_lo[pl] = ((a0 + pl * 14 * 2048) >> 4) & 0x3FFF; b_lo[pl] = ((b0 + pl * ncores * ncores * 128) >> 4) & 0x3FFF; }
        const int ksteps = n_mma / 6;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int c = 0; c < (ksteps + 1) / 2; ++c) {
                uint32_t el = 0;
                asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(el));
                if (el) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int kg = 2 * c + kk;
                        if (kg < ksteps) {
                            const uint32_t ao = (uint32_t)kg * 256u, bo = (uint32_t)kg * 16u;
                            const uint64_t a0d = a_hi | (uint64_t)(a_lo[0] + ao), a1d = a_hi | (uint64_t)(a_lo[1] + ao), a2d = a_hi | (uint64_t)(a_lo[2] + ao);
                            const uint64_t b0d = b_hi | (uint64_t)(b_lo[0] + bo), b1d = b_hi | (uint64_t)(b_lo[1] + bo), b2d = b_hi | (uint64_t)(b_lo[2] + bo);
#define MMA2(A, B, ACC) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(A), "l"(B), "r"(idesc), "r"(ACC) : "memory")
                            MMA2(a0d, b0d, kg > 0 ? 1u : 0u); MMA2(a0d, b1d, 1u); MMA2(a1d, b0d, 1u); MMA2(a1d, b1d, 1u); MMA2(a0d, b2d, 1u); MMA2(a2d, b0d, 1u);
                        }
                    }
                }
                __syncwarp();
            }
            if (tid == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
            const long long t2 = clock64();
            if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
        if (tid == 0) s_stop = 1;
    } else if ((mode & 8) && warp > 0) {
        // background traffic of "epilogue" warps while the MMAs run
        uint8_t* scratch = base + 150000 + tid * 16;
        uint32_t u[8];
        while (!s_stop) {
            if (mode & (16 | 64)) *reinterpret_cast<uint4*>(scratch) = make_uint4(tid, 1, 2, 3);
            if (mode & 16) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (mode & 32) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\ntcgen05.wait::ld.sync.aligned;"
                                        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                                        : "r"(tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u) : "memory");
        }
    } else
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(base), b0 = smem_u32(base) + 3 * 14 * 2048;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            if (mode & 4) {
                // the issue pattern of dense_tc.cu: descriptors = constant high part | (plane base + K-step offset)
                const uint64_t a_hi = make_desc(0, 2048, 128), b_hi = make_desc(0, 128, ncores * 128);
                uint32_t a_lo[3], b_lo[3];
                for (int pl = 0; pl < 3; ++pl) { a_lo[pl] = ((a0 + pl * 14 * 2048) >> 4) & 0x3FFF; b_lo[pl] = ((b0 + pl * ncores * ncores * 128) >> 4) & 0x3FFF; }
                const int ksteps = n_mma / 6;
                for (int c = 0; c < (ksteps + 1) / 2; ++c) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int kg = 2 * c + kk;
                        if (kg < ksteps) {
                            const uint32_t ao = (uint32_t)kg * 256u, bo = (uint32_t)kg * 16u;
                            const uint64_t a0d = a_hi | (uint64_t)(a_lo[0] + ao), a1d = a_hi | (uint64_t)(a_lo[1] + ao), a2d = a_hi | (uint64_t)(a_lo[2] + ao);
                            const uint64_t b0d = b_hi | (uint64_t)(b_lo[0] + bo), b1d = b_hi | (uint64_t)(b_lo[1] + bo), b2d = b_hi | (uint64_t)(b_lo[2] + bo);
#define MMA(A, B, ACC) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(A), "l"(B), "r"(idesc), "r"(ACC) : "memory")
                            MMA(a0d, b0d, kg > 0 ? 1u : 0u); MMA(a0d, b1d, 1u); MMA(a1d, b0d, 1u); MMA(a1d, b1d, 1u); MMA(a0d, b2d, 1u); MMA(a2d, b0d, 1u);
                        }
                    }
                }
            } else
            for (int i = 0; i < n_mma; ++i) {
                const int kg = (i / 6) % 7, term = i % 6;
                const int pa = (mode & 1) ? (term == 2 || term == 3 ? 1 : (term == 5 ? 2 : 0)) : 0;
                const int pb = (mode & 1) ? (term == 1 || term == 3 ? 1 : (term == 4 ? 2 : 0)) : 0;
                const uint64_t ad = make_desc(a0 + pa * 14 * 2048 + kg * 2 * 2048, 2048, 128);
                const uint64_t bd = (mode & 2) ? make_desc(b0 + pb * ncores * ncores * 128 + kg * 2 * ncores * 128, ncores * 128, 128)
                                               : make_desc(b0 + pb * ncores * ncores * 128 + kg * 256, 128, ncores * 128);
                const uint32_t id = (mode & 2) ? (idesc | (1u << 16)) : idesc;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm + (uint32_t)((i % n_acc) * 128)), "l"(ad), "l"(bd), "r"(id), "r"(i >= n_acc ? 1u : 0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
            const long long t2 = clock64();
            out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
int main(int argc, char** argv) {
    const int N = atoi(argv[1]), n_mma = atoi(argv[2]), n_acc = atoi(argv[3]), mode = atoi(argv[4]);
    long long* out; cudaMalloc(&out, 64);
    const int smem = 180000;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, (mode & 128) ? 512 : 128, smem>>>(N, n_mma, n_acc, mode, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[6]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("N=%d n_mma=%d n_acc=%d mode=%d: %s  issue %lld cyc, done %lld cyc -> %.1f cyc/MMA (nominal %d)\n", N, n_mma, n_acc, mode,
           cudaGetErrorString(e), h[4], h[5], (double)h[5] / n_mma, 128 * N / 256);
    return 0;
}
