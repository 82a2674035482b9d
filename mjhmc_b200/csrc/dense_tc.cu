// K4 (fp32 states): fused sampler for the full-covariance Gaussian on the 5th-generation tensor
// cores -- tcgen05.mma kind::tf32 with the accumulator in TMEM, the matrix staged by a TMA bulk
// copy, 3xTF32 operand splitting for fp32-grade gradients.
//
//   dEdX = S x,  E = x.(S x)/2,  S = (J + J^T)/2          misc/distributions.py:268-273
//
// Mapping.  A CTA works on a tile of 128 particles with 4 threads per particle (512 threads): thread
// (m, q) owns a quarter of the dims of particle m -- that slice of the momentum lives in its
// registers, the position in the A-operand tile in shared memory.  (One thread per particle, the
// first version, left one warp per scheduler and 90 % issue stalls: profiles/r1_dense_tc_v1.txt.)
// The gradient of the whole tile is one accumulator
//       D[128 particles x N dims] = Xtile[128 x K] . S^T[K x N]
//   A = Xtile : K-major, no swizzle: 8-particle x 4-dim core matrices (16 bytes per particle row), the
//       16 particle groups of one 4-dim core column contiguous (SBO = 128 B, LBO = 2 KB), so the 32
//       lanes of a warp store 32 consecutive 16-byte rows -- conflict-free float4 stores
//   B = S     : K-major, no swizzle (8 x 16-byte core matrices), pre-tiled in HBM by
//       tf32_prep_kernel and brought in with ONE cp.async.bulk (TMA) per CTA
//   D         : TMEM lane = particle, column = dim, so tcgen05.ld.32x32b hands every thread the
//       gradient of its own particle; kick, drift and x.g are thread-local (no shuffles)
// 3xTF32: x = x_hi + x_lo, S = S_hi + S_lo with *_hi carrying the top 10 mantissa bits;
//   D = x_hi S_hi + x_lo S_hi + x_hi S_lo  (three MMAs per 8-wide K step) is accurate to ~2^-21.
// One elected thread issues the MMAs; completion reaches the CTA through tcgen05.commit -> mbarrier.
#include "dense.h"

namespace mjhmc {

constexpr int kTcTile = 128;                   // particles per tile (M of the MMA)
constexpr int kTcSplit = 4;                    // threads per particle: each owns a quarter of the dims
constexpr int kTcThreads = kTcTile * kTcSplit; // 16 warps: warp w reads TMEM lanes 32 (w % 4) ..
constexpr int kTcMaxDim = 104;                 // padded dims (N and K of the MMA)
constexpr int kTcMaxCores = kTcMaxDim / 4;     // 4-dim core columns of the A tile
constexpr int kTcCoresPerThread = (kTcMaxCores + kTcSplit - 1) / kTcSplit;   // 7
constexpr int kTcDimsPerThread = kTcCoresPerThread * 4;                      // 28
constexpr int kTcTmemCols = 128;
constexpr int kTcCoreColBytes = 2048;          // one 4-dim core column of the A tile: 16 particle groups x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
    uint32_t u[8];
    // load and wait in ONE asm statement: the registers are only defined after tcgen05.wait::ld, and a
    // separate wait statement would not stop the compiler from consuming them earlier
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __uint_as_float(u[j]);
}

// 32 consecutive accumulator columns of this thread's TMEM lane (load + wait in one statement)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    uint32_t u[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(u[j]);
}

// shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)layout_type << 61;              // 0 = no swizzle, 2 = SWIZZLE_128B
    return d;
}

// top 19 bits (sign, exponent, 10 mantissa bits): exactly what the tf32 datapath reads
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// byte offset of the 4-dim row of particle m in core column kc (dims 4kc..4kc+3) of an A tile
__device__ __forceinline__ uint32_t a_row_offset(int m, int kc) {
    return (uint32_t)kc * kTcCoreColBytes + (uint32_t)(m >> 3) * 128u + (uint32_t)(m & 7) * 16u;
}

// Pre-tile S (fp32, d x d row-major) into the K-major core-matrix layout, split into hi / lo.
// out: [hi block | lo block], each ngroups * kcores * 32 floats; element (n, k) at
//      (n/8)*(kcores*32) + (k/4)*32 + (n%8)*4 + (k%4)
__global__ void tf32_prep_kernel(const float* __restrict__ S, int d, int ngroups, int kcores, float* __restrict__ out) {
    const int total = ngroups * 8 * kcores * 4;
    const int block = ngroups * kcores * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx / (kcores * 4), k = idx - n * (kcores * 4);
        const float val = (n < d && k < d) ? S[n * d + k] : 0.0f;
        const float hi = tf32_hi(val);
        const int off = (n >> 3) * (kcores * 32) + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
        out[off] = hi;
        out[block + off] = val - hi;
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
dense_tf32_kernel(const __grid_constant__ LaunchParams p, const float* __restrict__ Btiled) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t s_tmem;
    __shared__ unsigned long long s_tile;
    __shared__ int s_coin;
    __shared__ float s_red[4][kTcSplit][kTcTile];     // partial e_start, e_end, ev_start, ev_end per (part, particle)
    __shared__ unsigned int s_code[kTcTile];          // decision of each particle, broadcast to its 4 threads

    const int d = p.d;
    const int ksteps = (d + 7) >> 3;               // MMA K = 8
    const int N = ((d + 15) >> 4) << 4;            // MMA N: a multiple of 16 for M = 128
    const int kcores = ksteps * 2;
    const int ngroups = N >> 3;
    const uint32_t a_bytes = (uint32_t)kcores * kTcCoreColBytes;
    const uint32_t b_bytes = (uint32_t)ngroups * kcores * 128u;
    uint8_t* Ahi = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
    uint8_t* Alo = Ahi + a_bytes;
    uint8_t* Bhi = Alo + a_bytes;
    uint8_t* Blo = Bhi + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m = tid & (kTcTile - 1);             // particle of the tile
    const int q = tid / kTcTile;                   // which slice of the dims
    const int per = (kcores + kTcSplit - 1) / kTcSplit;
    const int kc0 = q * per;                       // my core columns [kc0, kc1)
    const int kc1 = min(kcores, kc0 + per);
    const bool lead = q == 0;                      // the thread that decides for the particle

    // ---- one-time setup: barriers, TMEM, the matrix via TMA
    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    if (tid == 0) {
        mbar_expect_tx(&bar_tma, 2u * b_bytes);
        tma_bulk_g2s(Bhi, Btiled, 2u * b_bytes, &bar_tma);
    }
    mbar_wait(&bar_tma, 0);

    // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D = f32, A = B = tf32,
    // both K-major, N, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi_addr = smem_u32(Ahi), a_lo_addr = smem_u32(Alo);
    const uint32_t b_hi_addr = smem_u32(Bhi), b_lo_addr = smem_u32(Blo);
    const uint32_t b_sbo = (uint32_t)kcores * 128u;
    uint32_t mma_phase = 0;

    const float eps = (float)p.eps, nhe = (float)(-p.eps / 2.0);
    const int L = p.L, sampler = p.sampler;
    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0, n_E = 0, n_exec = 0;
    unsigned long long* work_head = p.counters + (size_t)MJHMC_COUNTER_STRIPES * MJHMC_N_COUNTERS + 1;
    // my TMEM window: lane quarter of my warp, columns of my dims (a x32 load may run past them: unused)
    const uint32_t my_tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kc0 * 4);

    float v[kTcDimsPerThread];

    // D = Xtile . S^T for the positions currently in the A tiles (all threads call this)
    auto tile_gradient = [&]() {
        fence_async_smem();                        // our generic-proxy stores to A -> visible to the tensor core
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int kg = 0; kg < ksteps; ++kg) {
                const uint64_t ah = make_desc(a_hi_addr + kg * 2 * kTcCoreColBytes, kTcCoreColBytes, 128, 0);
                const uint64_t al = make_desc(a_lo_addr + kg * 2 * kTcCoreColBytes, kTcCoreColBytes, 128, 0);
                const uint64_t bh = make_desc(b_hi_addr + kg * 256, 128, b_sbo, 0);
                const uint64_t bl = make_desc(b_lo_addr + kg * 256, 128, b_sbo, 0);
                umma_tf32(tmem_base, ah, bh, idesc, kg > 0 ? 1u : 0u);
                umma_tf32(tmem_base, al, bh, idesc, 1u);
                umma_tf32(tmem_base, ah, bl, idesc, 1u);
            }
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, mma_phase);
        mma_phase ^= 1u;
        tc_fence_after();
    };

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(work_head, 1ull);
        __syncthreads();
        const unsigned long long tile = s_tile;
        __syncthreads();
        if ((long long)(tile * kTcTile) >= p.n) break;
        const long long i = (long long)tile * kTcTile + m;
        const bool live = i < p.n;

        unsigned int cflags = 0;
        float Hc = 0.0f;
        double dwell = 0.0;
        bool failed = false;
        if (lead && live && sampler == MJHMC_SAMPLER_MARKOV_JUMP) { cflags = p.ca_in[i]; Hc = ((const float*)p.Hc_in)[i]; }

        for (int it = 0; it < p.n_iter; ++it) {
            const unsigned long long attempt = p.attempt0 + (unsigned long long)it;
            const float* Xc = (const float*)(it == 0 ? p.Xin : p.Xout);
            const float* Vc = (const float*)(it == 0 ? p.Vin : p.Vout);
            float* Xo = (float*)p.Xout;
            float* Vo = (float*)p.Vout;
            const bool active = lead && live && !failed;
            if (sampler == MJHMC_SAMPLER_DISCRETE) {
                if (tid == 0) s_coin = draw_coin(p, attempt) < p.p_r;
            }

            float Hflf = Hc, H = 0.0f, Hl = 0.0f;
            bool need = false;
            if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
                need = active && !(cflags & 2u);
                if (active && !(cflags & 1u)) n_E += 1;
            }
            const int first_pass = __syncthreads_or(need ? 1 : 0) ? 0 : 1;
            const bool coin_fired = s_coin != 0;                       // read behind the barrier above

            for (int pass = first_pass; pass < 2; ++pass) {
                // ---- load my slice of (x, +-v); x goes to the A tiles split into hi / lo
                const float sign = pass == 0 ? -1.0f : 1.0f;
                float ev = 0.0f;
#pragma unroll
                for (int j = 0; j < kTcDimsPerThread; ++j) v[j] = 0.0f;
#pragma unroll
                for (int h = 0; h < kTcCoresPerThread; ++h) {
                    const int kc = kc0 + h;
                    if (kc < kc1) {
                        float xs[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = kc * 4 + j;
                            xs[j] = 0.0f;
                            v[h * 4 + j] = 0.0f;
                            if (live && k < d) { xs[j] = Xc[(long long)k * p.ld + i]; v[h * 4 + j] = sign * Vc[(long long)k * p.ld + i]; }
                            ev += v[h * 4 + j] * v[h * 4 + j];
                        }
                        const float4 hi = make_float4(tf32_hi(xs[0]), tf32_hi(xs[1]), tf32_hi(xs[2]), tf32_hi(xs[3]));
                        const uint32_t off = a_row_offset(m, kc);
                        *reinterpret_cast<float4*>(Ahi + off) = hi;
                        *reinterpret_cast<float4*>(Alo + off) = make_float4(xs[0] - hi.x, xs[1] - hi.y, xs[2] - hi.z, xs[3] - hi.w);
                    }
                }
                float e_start = 0.0f, e_end = 0.0f;

                for (int st = 0; st <= L; ++st) {
                    tile_gradient();
                    // ---- sweep over my slice of my particle's gradient (one TMEM round trip)
                    float g[32];
                    tmem_ld32(my_tmem, g);
#pragma unroll
                    for (int h = 0; h < kTcCoresPerThread; ++h) {
                        const int kc = kc0 + h;
                        if (kc < kc1) {
                            const uint32_t off = a_row_offset(m, kc);
                            float4* ph = reinterpret_cast<float4*>(Ahi + off);
                            float4* pl = reinterpret_cast<float4*>(Alo + off);
                            const float4 xh = *ph, xl = *pl;
                            float xs[4] = {xh.x + xl.x, xh.y + xl.y, xh.z + xl.z, xh.w + xl.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float gk = g[h * 4 + j];
                                float vv = v[h * 4 + j];
                                e_start += st == 0 ? xs[j] * gk : 0.0f;
                                e_end += st == L ? xs[j] * gk : 0.0f;
                                vv += st > 0 ? nhe * gk : 0.0f;        // second half kick of step st
                                vv += st < L ? nhe * gk : 0.0f;        // first half kick of step st+1
                                xs[j] += st < L ? eps * vv : 0.0f;     // drift of step st+1
                                v[h * 4 + j] = vv;
                            }
                            if (st < L) {
                                const float4 hi = make_float4(tf32_hi(xs[0]), tf32_hi(xs[1]), tf32_hi(xs[2]), tf32_hi(xs[3]));
                                *ph = hi;
                                *pl = make_float4(xs[0] - hi.x, xs[1] - hi.y, xs[2] - hi.z, xs[3] - hi.w);
                            }
                        }
                    }
                }
                if (L == 0) e_end = e_start;
                float ev_end = 0.0f;
#pragma unroll
                for (int j = 0; j < kTcDimsPerThread; ++j) ev_end += v[j] * v[j];   // slots beyond my dims hold 0
                // ---- per-particle totals: the four slices meet in shared memory
                s_red[0][q][m] = e_start; s_red[1][q][m] = e_end; s_red[2][q][m] = ev; s_red[3][q][m] = ev_end;
                __syncthreads();
                if (lead) {
                    float t[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        t[r] = 0.0f;
#pragma unroll
                        for (int qq = 0; qq < kTcSplit; ++qq) t[r] += s_red[r][qq][m];
                    }
                    const float h_start = 0.5f * t[0] + 0.5f * t[2];   // EX + EV, hmc_state.py:80-84
                    const float h_end = 0.5f * t[1] + 0.5f * t[3];
                    if (pass == 0) { if (need) { Hflf = h_end; n_exec += 1; } }
                    else { H = h_start; Hl = h_end; }
                }
            }
            if (active) { n_E += 1; n_exec += 1; }

            // ---- decision by the lead thread of each particle (same device code as the register-resident kernel)
            unsigned int take = 0, flip = 0, refresh = 0, choice = 0;
            if (active) {
                if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
                    const Decision dc = decide_mj(p, i, attempt, (double)(H - Hl), (double)(H - Hflf));
                    if (dc.fail) { report_failure(p, it); failed = true; }
                    else {
                        choice = dc.choice; dwell = dc.dwell;
                        if (choice == 0) { take = 1; Hc = H; cflags = 3u; n_l += 1; }
                        else if (choice == 1) { flip = 1; Hc = Hl; cflags = 2u; n_f += 1; }
                        else { refresh = 1; cflags = 0u; n_r += 1; }
                    }
                } else if (sampler == MJHMC_SAMPLER_CONTINUOUS_TIME) {
                    const Decision dc = decide_ct(p, i, attempt, (double)(H - Hl));
                    if (dc.fail) { report_failure(p, it); failed = true; }
                    else {
                        choice = dc.choice; dwell = dc.dwell;
                        if (choice == 1) { take = 2; n_fl += 1; }
                        else if (choice == 0) { flip = 1; n_f += 1; }
                        else { refresh = 1; n_r += 1; }
                    }
                } else {
                    const Decision dc = decide_discrete(p, i, attempt, (double)(H - Hl), coin_fired);
                    choice = dc.choice;
                    const bool acc = choice & 1u, fl = choice & 2u;
                    if (acc) take = 2;
                    flip = fl; refresh = (choice & 4u) ? 1u : 0u;
                    n_l += (acc && fl); n_f += (fl && !acc); n_fl += (acc && !fl); n_r += refresh;
                }
            }
            if (lead) s_code[m] = take | (flip << 2) | (refresh << 3) | ((active && !failed) ? 16u : 0u);
            __syncthreads();
            const unsigned int code = s_code[m];
            const unsigned int tk = code & 3u;
            const bool fp = code & 4u, rf = code & 8u, ok = code & 16u;

            // ---- apply: my slice of my particle's new state goes to the output arrays
            if (live) {
#pragma unroll
                for (int h = 0; h < kTcCoresPerThread; ++h) {
                    const int kc = kc0 + h;
                    if (kc < kc1) {
                        const uint32_t off = a_row_offset(m, kc);
                        const float4 xh = *reinterpret_cast<const float4*>(Ahi + off), xl = *reinterpret_cast<const float4*>(Alo + off);
                        const float xt[4] = {xh.x + xl.x, xh.y + xl.y, xh.z + xl.z, xh.w + xl.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = kc * 4 + j;
                            if (k < d) {
                                const long long o = (long long)k * p.ld + i;
                                float xn, vn;
                                if (ok && tk) { xn = xt[j]; vn = tk == 1 ? v[h * 4 + j] : -v[h * 4 + j]; }
                                else { xn = Xc[o]; vn = Vc[o]; }
                                if (ok && fp) vn = -vn;
                                if (ok && rf) {
                                    double z0, z1;
                                    normal_pair(p, i, attempt, k >> 1, d, z0, z1);
                                    vn = vn * (float)p.r_keep + (float)((k & 1) ? z1 : z0) * (float)p.r_mix;   // hmc_state.py:126
                                }
                                Xo[o] = xn;
                                Vo[o] = vn;
                                if (ok && p.samples) ((float*)p.samples)[(long long)k * p.s_stride_k + (long long)it * p.s_stride_it + i] = xn;
                            }
                        }
                    }
                }
                if (lead && ok) {
                    if (p.dwell) p.dwell[(long long)it * p.n + i] = dwell;
                    if (p.choice) p.choice[(long long)it * p.n + i] = (uint8_t)choice;
                }
            }
            __syncthreads();                       // s_code / s_red / the A tiles are reused by the next iteration
        }
        if (live) {
            if (lead) {
                if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) { p.ca_out[i] = (uint8_t)cflags; ((float*)p.Hc_out)[i] = Hc; }
                if (p.dwell_last && sampler != MJHMC_SAMPLER_DISCRETE) p.dwell_last[i] = dwell;
            }
            if (p.n_iter == 0) {
                for (int k = kc0 * 4; k < min(d, kc1 * 4); ++k) {
                    ((float*)p.Xout)[(long long)k * p.ld + i] = ((const float*)p.Xin)[(long long)k * p.ld + i];
                    ((float*)p.Vout)[(long long)k * p.ld + i] = ((const float*)p.Vin)[(long long)k * p.ld + i];
                }
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTcTmemCols);
    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, n_E, n_exec};
    flush_counters(p.counters, loc, (unsigned long long)L);
}

bool dense_tf32_supported(int kind, int ndims) {
    return kind == MJHMC_DIST_DENSE_GAUSSIAN && ndims >= 1 && ndims <= kTcMaxDim;
}

static void tf32_shape(int d, int& ksteps, int& N, int& kcores, int& ngroups) {
    ksteps = (d + 7) >> 3; N = ((d + 15) >> 4) << 4; kcores = ksteps * 2; ngroups = N >> 3;
}

long long dense_tf32_workspace_bytes(int ndims) {
    int ksteps, N, kcores, ngroups;
    tf32_shape(ndims, ksteps, N, kcores, ngroups);
    return 2ll * ngroups * kcores * 128;
}

// Fills the pre-tiled hi / lo copy of S (once per distribution; the sampler launches only read it).
cudaError_t dense_tf32_prepare(const float* S, int ndims, float* workspace, cudaStream_t stream) {
    int ksteps, N, kcores, ngroups;
    tf32_shape(ndims, ksteps, N, kcores, ngroups);
    tf32_prep_kernel<<<32, 256, 0, stream>>>(S, ndims, ngroups, kcores, workspace);
    return cudaGetLastError();
}

cudaError_t launch_dense_tf32(const LaunchParams& p, cudaStream_t stream) {
    int ksteps, N, kcores, ngroups;
    tf32_shape(p.d, ksteps, N, kcores, ngroups);
    if (!p.a1) return cudaErrorInvalidValue;                   // the pre-tiled matrix (mjhmc_dense_tf32_prepare)
    const size_t a_bytes = (size_t)kcores * kTcCoreColBytes;
    const size_t b_bytes = (size_t)ngroups * kcores * 128;
    const size_t smem = 2 * a_bytes + 2 * b_bytes + 1024;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e = cudaFuncSetAttribute(dense_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long tiles = (p.n + kTcTile - 1) / kTcTile;
    if (tiles > sms) tiles = sms;
    dense_tf32_kernel<<<(unsigned)tiles, kTcThreads, smem, stream>>>(p, (const float*)p.a1);
    return cudaGetLastError();
}

}  // namespace mjhmc
