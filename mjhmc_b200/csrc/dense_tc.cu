// K4 (fp32 states): fused sampler for the dense-contraction energies on the 5th-generation tensor cores --
// tcgen05.mma with the accumulators in TMEM, the matrix staged by one TMA bulk copy per CTA.
//
//   full-covariance Gaussian  dEdX = S x, E = x.(S x)/2, S = (J + J^T)/2                 misc/distributions.py:268-273
//   ProductOfT                Y = W^T x + b, E = sum_j (nu_j+1)/2 log(1 + (y_j/nu_j)^2),   misc/distributions.py:420-433
//                             G_j = (nu_j+1) y_j / (nu_j^2 + y_j^2), dEdX = W G            (autodiff of :431)
//
// Rows are JOBS.  A tile is 128 independent trajectories: the L trajectory of every particle of a chunk and, packed
// behind them, the F.L.F trajectories of only those particles whose cached FLF energy is not valid
// (hmc_state.py:109-119).  Round 1 ran the FLF pass for a whole tile whenever one of its 128 particles needed it --
// twice the work for ~1.1x the trajectories.  Persistent CTAs own contiguous particle ranges and cut them into
// chunks of P particles with P + #FLF(P) <= 128.
//
// Arithmetic: fp32 operands are split into three bf16 planes (x = x0 + x1 + x2, 8 + 8 + 8 mantissa bits; the
// split is exact) and a product is the six MMAs x0w0 + x0w1 + x1w0 + x1w1 + x0w2 + x2w0 of kind::f16 (bf16 in,
// fp32 accumulate): the dropped terms are below 2^-24.  bf16 -- not tf32 as in round 1 -- because ProductOfT needs
// W in BOTH orientations (Y = X W contracts over dims, dEdX = G W^T over experts): one no-swizzle buffer of
// 8 x 16-byte core matrices can be described K-major (N = dim, K = expert) and MN-major (N = expert, K = dim) at
// once, and the tensor core honours the MN-major view for 16-bit operands but returns zeros for tf32
// (tools/probe/tc_probe2.cu / tc_probe3.cu on the GPU box).  Two tf32 copies of a 100 x 100 W (hi + lo each) do not
// fit beside the particle tile in 227 KB; three bf16 planes of ONE copy do (75 KB), at the same MMA cycle count.
//
// Pipeline inside a tile.  16 epilogue warps (4 threads per row: thread (m, q) owns the 8-wide core columns
// of row m listed by tc_core() -- positions and momenta of those dims stay in its registers), one MMA warp and three
// helper warps.  A product is cut into K chunks of 16 / 32 columns.  The epilogue threads write chunk c of the next A
// operand (positions after the drift, or G for ProductOfT) and arrive on bar_chunk[c]; the MMA warp waits for chunk c
// only and issues its MMAs while the epilogue threads work on chunk c+1.  The LAST K chunk is issued per group of
// accumulator columns (the column ranges of the epilogue's chunks) with one commit each (bar_done[c]): the epilogue of
// the product starts on the first columns while the tensor core still works on the others.  The accumulator is
// double-buffered in TMEM (the Gaussian alternates D0 / D1, ProductOfT keeps Y in D0 and dEdX in D1), so the MMAs of
// product n+1 overwrite nothing the epilogue of product n still reads.  Round 1's kernel was MMA -> epilogue -> MMA
// strictly serial (tensor pipe 29.7 % active).
//
// Around the trajectory.  The (dims x 128 particles) boxes of X and V of a tile live in a shared-memory stash moved
// by TMA tensor boxes (cp.async.bulk.tensor, 32 particle columns each): loaded one tile ahead, read by the job rows,
// overwritten with the new state and stored -- X, V and the sample record -- by the helper warps, which also draw the
// tile's Philox uniforms and refresh the momenta of the R movers of the tile before, beside the trajectory.
#include <cstdio>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda.h>
#include "dense.h"
#include "stream.h"

namespace mjhmc {

constexpr int kTcRows = 128;                       // job rows per tile (M of the MMA)
constexpr int kTcSplit = 4;                        // epilogue threads per row
constexpr int kTcEpiThreads = kTcRows * kTcSplit;  // 16 warps; warp w reads TMEM lanes 32 (w % 4) ..
constexpr int kTcHelpers = 96;                     // three helper warps: draws of the tile, momentum refresh of the tile before
constexpr int kTcThreads = kTcEpiThreads + 32 + kTcHelpers;   // + the MMA-issuing warp + the helpers (20 warps = 5 per
                                                   // SM sub-partition: the same 96-register cap as 17 warps)
constexpr int kTcDeferMax = 24;                    // more R movers than this in a tile: every thread refreshes them at once
constexpr int kTcMaxP = 112;                       // padded dims / experts (K and N of the MMAs), a multiple of 16
constexpr int kTcCPT = 4;                          // 8-wide core columns per thread (and K chunks per product)
constexpr int kTcTmemCols = 512;                   // two accumulators of 128 columns + three A planes of 64
constexpr uint32_t kTcACol = 256;                  // first TMEM column of A plane 0
constexpr uint32_t kTcPlaneCols = 64;              // TMEM columns per A plane (112 bf16 = 56 columns, padded)
constexpr int kTcPiece = 32;                       // particle columns per TMA box of the state stash (128-byte rows)
constexpr int kTcTabs = 5;                         // ProductOfT per-expert tables: nu+1, nu^2, b, (nu+1)/2, 1/nu^2

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// (dims x 32 particles) boxes of a (dims, n) array: global -> shared on an mbarrier, shared -> global in the thread's
// current bulk group; coordinate 0 = first particle, coordinate 1 = first dim (the sample record has the iteration between)
__device__ __forceinline__ void tma_box_g2s(void* dst_smem, const CUtensorMap* map, int c0, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"((uint64_t)map), "r"(c0), "r"(0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_box_s2g(const CUtensorMap* map, int c0, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"((uint64_t)map), "r"(c0), "r"(0), "r"(smem_u32(src_smem)) : "memory");
}
__device__ __forceinline__ void tma_box_s2g_3d(const CUtensorMap* map, int c0, int c1, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(0), "r"(smem_u32(src_smem)) : "memory");
}
// tensor maps of one launch (host: launch_tc_T); ok = the four state maps exist, smp_ok = the sample-record map too
struct TcMaps { CUtensorMap xin, vin, xout, vout, smp; int ok, smp_ok; };
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void helper_bar() { asm volatile("bar.sync 1, %0;" ::"n"(96) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (job rows x 16 bf16, two per 32-bit column, lane = row) is read from
// tensor memory, only B comes from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 4 consecutive 32-bit columns (8 bf16) of this thread's TMEM lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&u)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one lane of a converged warp (elect.sync): keeps the control flow warp-uniform, so the descriptor arithmetic of the
// MMA warp stays in uniform registers instead of being moved there (R2UR) in front of every UTCHMMA
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "elect.sync _|P1, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 8 consecutive accumulator columns of this thread's TMEM lane.  Load and wait in ONE asm statement: the registers
// are only defined after tcgen05.wait::ld, and a separate wait statement would not stop the compiler from consuming
// them earlier.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __uint_as_float(u[j]);
}

__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// shared-memory matrix descriptor, no swizzle (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    return d;
}

// Core column (8 dims) that thread slice q owns in K chunk c.  The first chunk is ONE K step (cores 0, 1: slices 0 and
// 1 only), the others two (cores 4c-2 .. 4c+1): the MMA warp can start a product after half the usual wait, and
// every core 0 .. 13 has exactly one owner.  kTcMaxP / 8 = "none" (fails every bound check).
__device__ __forceinline__ int tc_core(int c, int q) { return c ? 4 * c - 2 + q : (q < 2 ? q : kTcMaxP / 8); }

// x = x0 + x1 + x2 with bf16 parts (round to nearest; the remainders are exact): 8 values -> 4 packed columns per plane
// of this thread's TMEM lane.  taddr = lane | first column of the core in plane 0; the planes are kTcPlaneCols apart.
__device__ __forceinline__ void split3_store(const float (&x)[8], uint32_t taddr) {
    uint32_t p0[4], p1[4], p2[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const float a = x[2 * jj], b = x[2 * jj + 1];
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b);                 // .x = a: low half = the even K element
        const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0);
        const float ra = a - __uint_as_float(u0 << 16), rb = b - __uint_as_float(u0 & 0xFFFF0000u);
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(ra, rb);
        const uint32_t u1 = *reinterpret_cast<const uint32_t*>(&h1);
        const float sa = ra - __uint_as_float(u1 << 16), sb = rb - __uint_as_float(u1 & 0xFFFF0000u);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(sa, sb);
        p0[jj] = u0; p1[jj] = u1; p2[jj] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    tmem_st4(taddr, p0);
    tmem_st4(taddr + kTcPlaneCols, p1);
    tmem_st4(taddr + 2u * kTcPlaneCols, p2);
}

// Pre-tile the matrix (fp32, rows x cols row-major) into three bf16 planes of 8-row x 16-byte core matrices:
// element (r, c) at (r/8)*(nc*128) + (c/8)*128 + (r%8)*16 + (c%8)*2 bytes of its plane, nc = P/8.
// Gaussian: M = S.  ProductOfT: M = W[dim][expert]; the per-expert tables follow the planes.
__global__ void tc_prep_kernel(const float* __restrict__ Mx, int rows, int cols, int P, const float* __restrict__ nu,
                               const float* __restrict__ b, uint8_t* __restrict__ out) {
    const int nc = P >> 3;
    const uint32_t plane = (uint32_t)nc * nc * 128u;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P * P; idx += gridDim.x * blockDim.x) {
        const int r = idx / P, c = idx - r * P;
        const float val = (r < rows && c < cols) ? Mx[r * cols + c] : 0.0f;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(val);
        const float r1 = val - __bfloat162float(h0);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
        const uint32_t off = ((uint32_t)(r >> 3) * (nc * 128u) + (uint32_t)(c >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(c & 7) * 2u) >> 1;
        o[off] = h0;
        o[(plane >> 1) + off] = h1;
        o[plane + off] = h2;                       // 2 * plane bytes = plane bf16 elements
    }
    if (nu) {
        float* tab = reinterpret_cast<float*>(out + 3u * plane);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P; j += gridDim.x * blockDim.x) {
            const float n = j < cols ? nu[j] : 1.0f;
            tab[0 * P + j] = j < cols ? n + 1.0f : 0.0f;
            tab[1 * P + j] = n * n;
            tab[2 * P + j] = j < cols ? b[j] : 0.0f;
            tab[3 * P + j] = j < cols ? (n + 1.0f) * 0.5f : 0.0f;
            tab[4 * P + j] = 1.0f / (n * n);
        }
    }
}

// 17 warps: the fifth warp of one SM sub-partition caps the allocation at 16384 / (5 * 32) = 102 -> 96 registers
// PC: the padded size P when it is known at compile time (112 = the 100-d benchmarks: every core / chunk bound check
// and table offset below folds to a constant), 0 = taken from the launch
template <bool POT, int PC>
__global__ void __launch_bounds__(kTcThreads, 1)
dense_tc_kernel(const __grid_constant__ LaunchParams p, const __grid_constant__ TcMaps maps) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t bar_tma, bar_st, bar_done[kTcCPT], bar_chunk[kTcCPT];
    __shared__ uint32_t s_tmem;
    __shared__ int s_coin;
    __shared__ int s_wsum[4];
    __shared__ float s_red[2][kTcSplit][kTcRows];     // partial H at the start / the end of a trajectory per (slice, row)
    __shared__ int s_flf_row[kTcRows];                // particle of the chunk -> row of its FLF job, -1 = none
    __shared__ int s_row_part[kTcRows];               // FLF row -> particle of the chunk
    __shared__ unsigned int s_code[kTcRows];          // decision of each particle, broadcast to its 4 threads
    __shared__ int s_nr[2], s_rlist[2][kTcRows];      // particles of the tile whose momentum is refreshed (R moves); two tiles
    __shared__ double s_u[3][kTcRows];                // the uniforms of the tile's particles (drawn by the helper warps)
    __shared__ RaceDraws s_rd[kTcRows];               // and the uniform-only half of the holding-time race screen

    const int d = p.d;
    const int P = PC ? PC : ((d + 15) >> 4) << 4;  // padded dims (= experts): N of the MMAs and K in steps of 16
    const int ncores = P >> 3;
    const int ksteps = P >> 4;
    const int nchunks = (ksteps + 2) >> 1;         // K chunks of a product: steps [0,1), [1,3), [3,5), [5,7)
    const uint32_t b_plane = (uint32_t)ncores * ncores * 128u;
    const uint32_t ws_bytes = 3u * b_plane + (POT ? (uint32_t)(kTcTabs * P * 4) : 0u);
    uint8_t* B0 = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
    const float* tab = reinterpret_cast<const float*>(B0 + 3u * b_plane);
    // State stash: the (dims x 128 particles) boxes of X and V the tile starts from, as four TMA boxes of 32 particle
    // columns each (element (k, m) at stash_at(m) + 32 k).  Filled by the helper warps when the tile before has let go of
    // the stash -- beside this tile's plan -- read by the job rows (an FLF job shares the column of its particle's L
    // job); the new state of the tile is assembled in it and leaves as TMA stores.
    float* const stX = reinterpret_cast<float*>(B0 + ((ws_bytes + 127u) & ~127u));
    float* const stV = stX + P * kTcRows;
    auto stash_at = [&](int mcol) { return (mcol >> 5) * (P * kTcPiece) + (mcol & (kTcPiece - 1)); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool epi = tid < kTcEpiThreads;
    const bool mma_warp = warp == kTcEpiThreads / 32;
    const bool literal_race = (p.rng_flags & MJHMC_RNG_FLAG_LITERAL_RACE) != 0;
    const int m = tid & (kTcRows - 1);             // job row of the tile
    const int q = (tid >> 7) & 3;                  // which core columns: q, 4+q, 8+q, 12+q

    // ---- one-time setup: barriers, TMEM, the matrix planes via TMA
    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_st, 1);
#pragma unroll
        for (int c = 0; c < kTcCPT; ++c) { mbar_init(&bar_done[c], 1); mbar_init(&bar_chunk[c], kTcEpiThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    if (tid == 0) {
        mbar_expect_tx(&bar_tma, ws_bytes);
        tma_bulk_g2s(B0, p.ws, ws_bytes, &bar_tma);
    }
    mbar_wait(&bar_tma, 0);

    // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D = f32, A = B = bf16, A K-major,
    // N = P, M = 128; bit 16 selects the MN-major view of B
    const uint32_t idesc_k = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_mn = idesc_k | (1u << 16);
    const uint32_t b_addr = smem_u32(B0);
    // descriptor pieces for the MMA warp: the 14-bit start address (16-byte units) of every plane is the low word,
    // LBO / SBO / version the high part; a K step advances the start address by a constant
    const uint64_t b_desc_hi_k = make_desc(0, 128, (uint32_t)ncores * 128u);            // K-major view: N = row, K = column
    const uint64_t b_desc_hi_mn = make_desc(0, (uint32_t)ncores * 128u, 128);           // MN-major view: N = column, K = row
    const uint32_t b_kstep_mn = (uint32_t)(2 * ncores * 128) >> 4;
    uint32_t b_lo[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) b_lo[pl] = ((b_addr + pl * b_plane) >> 4) & 0x3FFFu;

    float eps = (float)p.eps, nhe = (float)(-p.eps / 2.0), neg_eps = nhe + nhe;
    // pinned: ptxas otherwise re-derives these from the double in the constant bank inside the sweep (DMUL + F2F per use)
    asm volatile("" : "+f"(eps), "+f"(nhe), "+f"(neg_eps));
    const int L = p.L, sampler = p.sampler;
    const bool mj = sampler == MJHMC_SAMPLER_MARKOV_JUMP;
    const int nprod = POT ? 2 * L + 2 : L + 1;     // products of one tile trajectory
#ifdef TC_NO_NSPLIT
    constexpr bool kNSplit = false;
#else
    constexpr bool kNSplit = PC == kTcMaxP;        // 4 K chunks, the last one two full K steps
#endif
    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0, n_E = 0, n_exec = 0;
    uint32_t pc = 0;                               // running product counter: barrier parities, accumulator choice
    // my TMEM window: lane quarter of my warp (warp % 4), accumulator column of core column kc = 8 kc
    const uint32_t my_lane = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t a_lane = tmem_base + my_lane + kTcACol;       // my row of A plane 0 (core kc = columns 4 kc ..)

    // contiguous particle range of this CTA
    const long long r0 = (p.n * (long long)blockIdx.x / gridDim.x) & ~3ll;
    const long long r1 = blockIdx.x + 1 == gridDim.x ? p.n : (p.n * (long long)(blockIdx.x + 1) / gridDim.x) & ~3ll;

    float x[kTcCPT][8], v[kTcCPT][8];
#ifdef TCX_TRACE
    __shared__ long long s_ev[12][16];
    int tile_no = 0;
#define TCX_EV(prod, slot) if (blockIdx.x == 0 && tile_no == 3 && (prod) >= 4 && (prod) < 16 && (tid == 0 || tid == 512)) s_ev[(prod) - 4][slot] = clock64();
#else
#define TCX_EV(prod, slot)
#endif
#ifdef TCX_TIMING
    long long tph[7] = {0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    long long tw[6] = {0, 0, 0, 0, 0, 0};          // epilogue: [0] wait bar_done, [1] work; MMA warp: [0..3] wait chunk c, [4] issue
#define TCX_T0 const long long tq0 = clock64();
#define TCX_ACC(k) tw[k] += clock64() - tq0;
#define TCX_MARK(k) { const long long tnow = clock64(); tph[k] += tnow - tlast; tlast = tnow; }
#else
#define TCX_MARK(k)
#define TCX_T0
#define TCX_ACC(k)
#endif

    // Momentum refresh (hmc_state.py:121-129) of the R movers of one tile: (particle, Box-Muller pair) jobs dealt to
    // the threads first, first + stride, ...
    auto refresh_jobs = [&](const int* list, int nr, long long base, unsigned long long att, int first, int stride) {
        const float rk = (float)p.r_keep, rm = (float)p.r_mix;
        float* Vo = (float*)p.Vout;
        const int npairs = (d + 1) >> 1, njobs = nr * npairs;
        for (int job = first; job < njobs; job += stride) {
            const int r = job / npairs, pr = job - r * npairs;
            const long long ip = base + list[r];
            double z0, z1;
            normal_pair(p, ip, att, pr, d, z0, z1);
            const long long o0 = (long long)(2 * pr) * p.ld + ip;
            Vo[o0] = (rk != 0.0f ? Vo[o0] * rk : 0.0f) + (float)z0 * rm;  // hmc_state.py:126
            if (2 * pr + 1 < d) {
                const long long o1 = o0 + p.ld;
                Vo[o1] = (rk != 0.0f ? Vo[o1] * rk : 0.0f) + (float)z1 * rm;
            }
        }
    };
    // the boxes of particles c0 .. c0 + 127 (columns past the end of the cloud arrive as zeros), issued by eight helper
    // threads; false: the tile fills the stash with plain loads
    const int ht = tid - (kTcEpiThreads + 32);     // helper thread index (< 0: not a helper)
    auto stash_issue = [&](bool first_iter, long long c0) -> bool {
        if (maps.ok && ht == 0) {                  // one thread, the tensor map named in the instruction (no address select)
            mbar_expect_tx(&bar_st, (uint32_t)(2 * kTcRows * d * 4));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
            for (int piece = 0; piece < kTcRows / kTcPiece; ++piece) {
                float* dx = stX + piece * (P * kTcPiece);
                float* dv = stV + piece * (P * kTcPiece);
                const int cc = (int)c0 + piece * kTcPiece;
                if (first_iter) { tma_box_g2s(dx, &maps.xin, cc, &bar_st); tma_box_g2s(dv, &maps.vin, cc, &bar_st); }
                else { tma_box_g2s(dx, &maps.xout, cc, &bar_st); tma_box_g2s(dv, &maps.vout, cc, &bar_st); }
            }
        }
        return maps.ok != 0;
    };
    uint32_t st_par = 0;                           // parity of the stash fill the next bulk-filled tile waits for
    bool stash_bulk = false;                       // the coming tile's stash was requested with bulk copies
    int tb = 0;                                    // tile parity: which s_rlist / s_nr this tile fills
    int pend_nr = 0, pend_buf = 0;                 // R movers of the tile before, refreshed by the helper warps during this one
    long long pend_cur = 0;

    for (int it = 0; it < p.n_iter; ++it) {
        const unsigned long long attempt = p.attempt0 + (unsigned long long)it;
        const float* Xc = (const float*)(it == 0 ? p.Xin : p.Xout);
        const float* Vc = (const float*)(it == 0 ? p.Vin : p.Vout);
        const uint8_t* cac = it == 0 ? p.ca_in : p.ca_out;
        const float* Hcc = (const float*)(it == 0 ? p.Hc_in : p.Hc_out);
        float* Xo = (float*)p.Xout;
        float* Vo = (float*)p.Vout;
        if (sampler == MJHMC_SAMPLER_DISCRETE && tid == 0) s_coin = draw_coin(p, attempt) < p.p_r;
        if (r0 < r1) stash_bulk = stash_issue(it == 0, r0);

        for (long long cur = r0; cur < r1;) {
            // ---- plan the tile: particles cur .. cur+np-1 with np + #FLF jobs <= 128
            unsigned int cflags = 0;
            int need = 0, valid = 0, incl = 0;
            if (tid < kTcRows) {
                const long long i = cur + tid;
                valid = i < r1;
                if (valid && mj) { cflags = cac[i]; need = !(cflags & 2u); }
                incl = valid ? 1 + need : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) s_wsum[warp] = incl;
            }
            if (tid == 0) s_nr[tb] = 0;
            __syncthreads();
            if (tid < kTcRows) {
                for (int w = 0; w < warp; ++w) incl += s_wsum[w];
                s_flf_row[tid] = -1;
            }
            int np = __syncthreads_count(tid < kTcRows && valid && incl <= kTcRows);
            // tiles start at multiples of 4 particles: the first element of a TMA box has to be 16-byte aligned
            if (cur + np < r1) np &= ~3;
            const bool mine = tid < np;                                     // lead thread of an L job
            if (mine && need) {
                const int row = np + (incl - (tid + 1) - 1);                // FLF rows follow the L rows, in particle order
                s_flf_row[tid] = row;
                s_row_part[row] = tid;
            }
            __syncthreads();
            const int nflf = __syncthreads_count(mine && need);
            const int nrows = np + nflf;

            TCX_MARK(0)
            // ---- my job: row m -> (particle, sign)
            const bool is_l = epi && m < np, is_flf = epi && m >= np && m < nrows;
            const int part = is_l ? m : (is_flf ? s_row_part[m] : 0);
            const float sign = is_flf ? -1.0f : 1.0f;
            const bool live = is_l || is_flf;

            // ---- the stash holds the boxes of particles cur .. cur + 127
            if (stash_bulk) {
                if (epi) mbar_wait(&bar_st, st_par);
                st_par ^= 1u;
            } else {
                // no tensor maps (rows that are not 16-byte aligned): plain loads
                if (ht >= 0) bulk_wait_read();      // the bulk stores of the tile before may still be reading the stash
                __syncthreads();
                for (int idx = tid; idx < 2 * d * kTcRows; idx += kTcThreads) {
                    const int a = idx / (d * kTcRows), rem = idx - a * (d * kTcRows);
                    const int k = rem / kTcRows, mm = rem - k * kTcRows;
                    const long long gi = cur + mm;
                    const float val = gi < p.n ? (a ? Vc : Xc)[(long long)k * p.ld + gi] : 0.0f;
                    (a ? stV : stX)[stash_at(mm) + k * kTcPiece] = val;
                }
                __syncthreads();
            }

            if (epi) {
                // ---- my slice of (x, +-v) from the stash
                float ev0 = 0.0f;
                const int sa = stash_at(part);
#pragma unroll
                for (int c = 0; c < kTcCPT; ++c) {
                    const int kc = tc_core(c, q);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const bool in = live && kc * 8 + j < d;
                        x[c][j] = in ? stX[sa + (kc * 8 + j) * kTcPiece] : 0.0f;
                        v[c][j] = in ? stV[sa + (kc * 8 + j) * kTcPiece] : 0.0f;
                    }
                }
#pragma unroll
                for (int c = 0; c < kTcCPT; ++c)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[c][j] *= sign;
                        ev0 += v[c][j] * v[c][j];
                    }
                float e_start = 0.0f, e_end = 0.0f;
                TCX_MARK(1)

                // ---- first A operand: the positions
#pragma unroll
                for (int c = 0; c < kTcCPT; ++c) {
                    if (c < nchunks) {
                        const int kc = tc_core(c, q);
                        if (kc < ncores) split3_store(x[c], a_lane + (uint32_t)(kc * 4));
                        tmem_st_wait();
                        tc_fence_before();
                        mbar_arrive(&bar_chunk[c]);
                    }
                }

                for (int st = 0; st <= L; ++st) {
                    if (POT) {
                        // ---- phase 1: Y = X W + b  ->  G (and the energy at the two ends of the trajectory)
                        const uint32_t par1 = pc & 1u;
                        ++pc;
#pragma unroll
                        for (int c = 0; c < kTcCPT; ++c) {
                            if (c < nchunks) {
                                mbar_wait(&bar_done[c], par1);             // columns of chunk c of Y are complete
                                TCX_EV(2 * st, 0)
                                tc_fence_after();
                                const int kc = tc_core(c, q);
                                if (kc < ncores) {
                                    float y[8];
                                    tmem_ld8(tmem_base + my_lane + (uint32_t)(kc * 8), y);
                                    const float4* t4 = reinterpret_cast<const float4*>(tab);
                                    float a1[8], n2[8], bb[8];
                                    *reinterpret_cast<float4*>(a1) = t4[(0 * P + kc * 8) >> 2]; *reinterpret_cast<float4*>(a1 + 4) = t4[((0 * P + kc * 8) >> 2) + 1];
                                    *reinterpret_cast<float4*>(n2) = t4[(1 * P + kc * 8) >> 2]; *reinterpret_cast<float4*>(n2 + 4) = t4[((1 * P + kc * 8) >> 2) + 1];
                                    *reinterpret_cast<float4*>(bb) = t4[(2 * P + kc * 8) >> 2]; *reinterpret_cast<float4*>(bb + 4) = t4[((2 * P + kc * 8) >> 2) + 1];
                                    if (st == 0 || st == L) {
                                        // (nu+1)/2 log(1 + (y/nu)^2), distributions.py:431, as log2(1 + t) from the
                                        // special-function unit (absolute error 2^-22 per term, below the rounding of
                                        // the fp32 sum over 100 experts); ln 2 is applied to the sum
                                        float e = 0.0f;
#pragma unroll
                                        for (int j = 0; j < 8; ++j) {
                                            const float yy = y[j] + bb[j];
                                            e = fmaf(tab[3 * P + kc * 8 + j], lg2_approx(fmaf(yy * yy, tab[4 * P + kc * 8 + j], 1.0f)), e);
                                        }
                                        e *= 0.693147180559945309f;
                                        if (st == 0) e_start += e;
                                        if (st == L) e_end += e;
                                    }
                                    float g[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float yy = y[j] + bb[j];
                                        g[j] = (a1[j] * yy) * rcp_approx(fmaf(yy, yy, n2[j]));
                                    }
                                    split3_store(g, a_lane + (uint32_t)(kc * 4));
                                }
                                tmem_st_wait();
                                tc_fence_before();
                                mbar_arrive(&bar_chunk[c]);
                                TCX_EV(2 * st, 1 + c)
                            }
                        }
                    }
                    // ---- gradient sweep: kick, drift, next positions
                    const uint32_t par2 = pc & 1u;
                    const uint32_t dcol = tmem_base + my_lane + (POT ? 128u : ((pc & 1u) ? 128u : 0u));
                    ++pc;
                    if (st > 0 && st < L) {
                        // steps 1 .. L-1: the closing half kick of step st and the opening one of step st+1 are one
                        // full kick, then the drift -- two FMAs per dim, then the bf16 split of the new positions
#pragma unroll
                        for (int c = 0; c < kTcCPT; ++c) {
                            if (c < nchunks) {
                                { TCX_T0
                                mbar_wait(&bar_done[c], par2);             // columns of chunk c of the gradient are complete
                                TCX_ACC(0) }
                                TCX_EV(POT ? 2 * st + 1 : st, 0)
                                tc_fence_after();
                                const int kc = tc_core(c, q);
                                if (kc < ncores) {
                                    float g[8];
#ifndef TCX_NOLD
                                    tmem_ld8(dcol + (uint32_t)(kc * 8), g);
#else
#pragma unroll
                                    for (int j = 0; j < 8; ++j) g[j] = x[c][j] * 1e-3f;
#endif
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        v[c][j] = fmaf(neg_eps, g[j], v[c][j]);
                                        x[c][j] = fmaf(eps, v[c][j], x[c][j]);
                                    }
#ifndef TCX_NOSTORE
                                    split3_store(x[c], a_lane + (uint32_t)(kc * 4));
#endif
                                }
                                tmem_st_wait();
                                tc_fence_before();
                                mbar_arrive(&bar_chunk[c]);
                                TCX_EV(POT ? 2 * st + 1 : st, 1 + c)
                            }
                        }
                    } else {
                        // the two ends of the trajectory: energies, half kicks
#pragma unroll
                        for (int c = 0; c < kTcCPT; ++c) {
                            if (c < nchunks) {
                                mbar_wait(&bar_done[c], par2);
                                tc_fence_after();
                                const int kc = tc_core(c, q);
                                if (kc < ncores) {
                                    float g[8];
                                    tmem_ld8(dcol + (uint32_t)(kc * 8), g);
                                    if (!POT) {
                                        float e = 0.0f;                      // E = x.(S x)/2
#pragma unroll
                                        for (int j = 0; j < 8; ++j) e = fmaf(x[c][j], g[j], e);
                                        if (st == 0) e_start += e;
                                        if (st == L) e_end += e;
                                    }
                                    if (L > 0) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) v[c][j] = fmaf(nhe, g[j], v[c][j]);   // opening / closing half kick
                                    }
                                    if (st < L) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) x[c][j] = fmaf(eps, v[c][j], x[c][j]);
                                        split3_store(x[c], a_lane + (uint32_t)(kc * 4));
                                    }
                                }
                                if (st < L) {
                                    tmem_st_wait();
                                    tc_fence_before();
                                    mbar_arrive(&bar_chunk[c]);
                                }
                            }
                        }
                    }
                }
                if (!POT) { e_start *= 0.5f; e_end *= 0.5f; }
                if (L == 0) e_end = e_start;
                float ev1 = 0.0f;
#pragma unroll
                for (int c = 0; c < kTcCPT; ++c)
#pragma unroll
                    for (int j = 0; j < 8; ++j) ev1 += v[c][j] * v[c][j];       // slots beyond my dims hold 0
                s_red[0][q][m] = e_start + 0.5f * ev0;                          // EX + EV, hmc_state.py:80-84
                s_red[1][q][m] = e_end + 0.5f * ev1;
            } else if (mma_warp) {
                // ---- the MMA warp: chunk c of a product is issued as soon as its 512 writers have arrived.
                // One elected lane issues; the descriptors are base + immediate (a single thread executes ~1
                // instruction per 4 cycles: rebuilding six 64-bit descriptors per K step cost 250 cycles per MMA in
                // tools/probe/tc_probe4.cu, against 56 for the MMA itself).
                for (int prod = 0; prod < nprod; ++prod) {
                    const bool y_prod = POT && !(prod & 1);                      // Y = X W: B read MN-major (K = dims)
                    const uint32_t dacc = tmem_base + (POT ? (y_prod ? 0u : 128u) : ((pc & 1u) ? 128u : 0u));
                    const uint32_t idesc = y_prod ? idesc_mn : idesc_k;
                    const uint64_t bhi = y_prod ? b_desc_hi_mn : b_desc_hi_k;
                    const uint32_t bstep = y_prod ? b_kstep_mn : 16u;            // descriptor address units (16 B) per K step
                    for (int c = 0; c < nchunks; ++c) {
                        { TCX_T0
                        mbar_wait(&bar_chunk[c], pc & 1u);
                        TCX_ACC(c) }
                        TCX_EV(prod - 1, 5 + c)
                        tc_fence_after();
                        TCX_T0
                        if (kNSplit && c == nchunks - 1) {
                            // Last K chunk: issued per group of accumulator columns (the column ranges of the epilogue's
                            // chunks: 16, 32, 32, 32) with one commit each, so the epilogue starts on the first columns of
                            // this product while the tensor core still works on the others.
                            if (elect_one()) {
#pragma unroll
                                for (int nc = 0; nc < kTcCPT; ++nc) {
                                    const uint32_t n0 = nc ? 32u * nc - 16u : 0u, nw = nc ? 32u : 16u;
                                    const uint32_t idn = (idesc & ~(0x3Fu << 17)) | ((nw >> 3) << 17);
                                    // 16-byte units per 8 accumulator columns: one core matrix along a row of cores
                                    // (MN-major view) or one row of cores (K-major view)
                                    const uint32_t bn = (n0 >> 3) * (y_prod ? 8u : (uint32_t)ncores * 8u);
#pragma unroll
                                    for (int kk = 0; kk < 2; ++kk) {
                                        const int kg = 2 * c - 1 + kk;
                                        const uint32_t bo = (uint32_t)kg * bstep + bn;
                                        const uint32_t a0t = tmem_base + kTcACol + (uint32_t)kg * 8u, a1t = a0t + kTcPlaneCols, a2t = a1t + kTcPlaneCols;
                                        const uint64_t b0d = bhi | (uint64_t)(b_lo[0] + bo), b1d = bhi | (uint64_t)(b_lo[1] + bo),
                                                       b2d = bhi | (uint64_t)(b_lo[2] + bo);
                                        umma_bf16_ts(dacc + n0, a0t, b0d, idn, 1u);
                                        umma_bf16_ts(dacc + n0, a0t, b1d, idn, 1u);
                                        umma_bf16_ts(dacc + n0, a1t, b0d, idn, 1u);
                                        umma_bf16_ts(dacc + n0, a1t, b1d, idn, 1u);
                                        umma_bf16_ts(dacc + n0, a0t, b2d, idn, 1u);
                                        umma_bf16_ts(dacc + n0, a2t, b0d, idn, 1u);
                                    }
                                    umma_commit(&bar_done[nc]);
                                }
                            }
                        } else if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const int kg = 2 * c - 1 + kk;
                                if (kg >= 0 && kg < ksteps) {
                                    const uint32_t bo = (uint32_t)kg * bstep;
                                    const uint32_t a0t = tmem_base + kTcACol + (uint32_t)kg * 8u, a1t = a0t + kTcPlaneCols, a2t = a1t + kTcPlaneCols;
                                    const uint64_t b0d = bhi | (uint64_t)(b_lo[0] + bo), b1d = bhi | (uint64_t)(b_lo[1] + bo),
                                                   b2d = bhi | (uint64_t)(b_lo[2] + bo);
#ifndef TCX_NOMMA
                                    umma_bf16_ts(dacc, a0t, b0d, idesc, kg > 0 ? 1u : 0u);
                                    umma_bf16_ts(dacc, a0t, b1d, idesc, 1u);
                                    umma_bf16_ts(dacc, a1t, b0d, idesc, 1u);
                                    umma_bf16_ts(dacc, a1t, b1d, idesc, 1u);
                                    umma_bf16_ts(dacc, a0t, b2d, idesc, 1u);
                                    umma_bf16_ts(dacc, a2t, b0d, idesc, 1u);
#endif
                                }
                            }
                            if (c == nchunks - 1) {
#pragma unroll
                                for (int cc = 0; cc < kTcCPT; ++cc)
                                    if (cc < nchunks) umma_commit(&bar_done[cc]);
                            }
                        }
                        __syncwarp();
                        TCX_EV(prod - 1, 9 + c)
                        TCX_ACC(4)
                    }
                    ++pc;
                }
            } else {
                // ---- the helper warps work beside the trajectory on what does not depend on it:
                // the uniforms of this tile's particles (Philox) with the uniform-only half of the race screen ...
                const bool discrete = sampler == MJHMC_SAMPLER_DISCRETE;
                for (int t = ht; t < np; t += kTcHelpers) {
                    const Uniform3 u = draw_uniforms(p, cur + t, attempt, !discrete && p.p_r != 0.0);
                    s_u[0][t] = u.u0; s_u[1][t] = u.u1; s_u[2][t] = u.u2;
                    if (!discrete) s_rd[t] = race_draws(p.p_r, u.u0, u.u1, u.u2, literal_race);
                }
                // ... and the momentum refresh of the R movers of the tile before (their momenta were stored before that
                // tile's closing barrier; nothing reads them before the next iteration)
                if (pend_nr > 0) {
                    bulk_wait_all();               // that tile's bulk stores (issued by these warps) have reached memory
                    asm volatile("fence.proxy.async;" ::: "memory");
                    helper_bar();
                    refresh_jobs(s_rlist[pend_buf], pend_nr, pend_cur, attempt, ht, kTcHelpers);
                }
            }
            pend_nr = 0;
            TCX_MARK(2)
            __syncthreads();
            TCX_MARK(3)

            // ---- decision by the lead thread of each L job (same device code as the register-resident kernel)
            unsigned int take = 0, flip = 0, refresh = 0, choice = 0, ok = 0;
            double dwell = 0.0;
            if (mine) {
                float H = 0.0f, Hl = 0.0f, Hflf = 0.0f;
#pragma unroll
                for (int qq = 0; qq < kTcSplit; ++qq) { H += s_red[0][qq][tid]; Hl += s_red[1][qq][tid]; }
                n_E += 1; n_exec += 1;
                float Hc = 0.0f;
                if (mj) {
                    if (!(cflags & 1u)) n_E += 1;                  // the reference evaluates the FLF state here
                    if (need) {
                        const int fr = s_flf_row[tid];
#pragma unroll
                        for (int qq = 0; qq < kTcSplit; ++qq) Hflf += s_red[1][qq][fr];
                        n_exec += 1;
                    } else {
                        Hflf = Hcc[cur + tid];
                    }
                    const Decision dc = decide_mj_s(p.p_r, s_u[0][tid], s_u[1][tid], s_u[2][tid], s_rd[tid], (double)(H - Hl), (double)(H - Hflf), p.dwell != nullptr || (it + 1 == p.n_iter && p.dwell_last != nullptr));
                    if (dc.fail) report_failure(p, it);
                    else {
                        ok = 1; choice = dc.choice; dwell = dc.dwell;
                        if (choice == 0) { take = 1; Hc = H; cflags = 3u; n_l += 1; }
                        else if (choice == 1) { flip = 1; Hc = Hl; cflags = 2u; n_f += 1; }
                        else { refresh = 1; Hc = Hflf; cflags = 0u; n_r += 1; }
                    }
                    if (!ok) Hc = need ? 0.0f : Hflf;
                    p.ca_out[cur + tid] = (uint8_t)(ok ? cflags : (need ? (cflags & ~2u) : cflags));
                    ((float*)p.Hc_out)[cur + tid] = Hc;
                } else if (sampler == MJHMC_SAMPLER_CONTINUOUS_TIME) {
                    const Decision dc = decide_ct_s(p.p_r, s_u[0][tid], s_u[1][tid], s_u[2][tid], s_rd[tid], (double)(H - Hl), p.dwell != nullptr || (it + 1 == p.n_iter && p.dwell_last != nullptr));
                    if (dc.fail) report_failure(p, it);
                    else {
                        ok = 1; choice = dc.choice; dwell = dc.dwell;
                        if (choice == 1) { take = 2; n_fl += 1; }
                        else if (choice == 0) { flip = 1; n_f += 1; }
                        else { refresh = 1; n_r += 1; }
                    }
                } else {
                    ok = 1; choice = decide_discrete_s(p.p_flip, s_u[0][tid], s_u[1][tid], (double)(H - Hl), s_coin != 0);
                    const bool acc = choice & 1u, fl = choice & 2u;
                    if (acc) take = 2;
                    flip = fl; refresh = (choice & 4u) ? 1u : 0u;
                    n_l += (acc && fl); n_f += (fl && !acc); n_fl += (acc && !fl); n_r += refresh;
                }
                s_code[tid] = take | (flip << 2) | (refresh << 3) | (ok << 4);
                if (ok && refresh) s_rlist[tb][atomicAdd(&s_nr[tb], 1)] = tid;
                if (ok) {
                    if (p.dwell) p.dwell[(long long)it * p.n + cur + tid] = dwell;
                    if (p.choice) p.choice[(long long)it * p.n + cur + tid] = (uint8_t)choice;
                    if (p.dwell_last && sampler != MJHMC_SAMPLER_DISCRETE) p.dwell_last[cur + tid] = dwell;
                }
            }
            const int any_fail = __syncthreads_or(mine && !ok);
            TCX_MARK(4)

            // ---- apply: the new state of the tile's particles is assembled in the stash (a particle that did not take its
            // trajectory is already there) and leaves as one bulk copy (TMA store) per row -- X, V and the sample record --
            // issued by the helper warps while the other warps plan the next tile.  (Per-thread stores: 84 scalar stores
            // with 64-bit row addresses each, 19k cycles per tile.)
            if (is_l) {
                const unsigned int code = s_code[m];
                const unsigned int tk = code & 3u;
                const bool fp = code & 4u, okk = code & 16u;
                const int sm_ = stash_at(m);
                if (okk && tk) {
                    const float vs = ((tk == 2) != fp) ? -1.0f : 1.0f;
#pragma unroll
                    for (int c = 0; c < kTcCPT; ++c) {
                        const int kc = tc_core(c, q);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (kc * 8 + j < d) { stX[sm_ + (kc * 8 + j) * kTcPiece] = x[c][j]; stV[sm_ + (kc * 8 + j) * kTcPiece] = vs * v[c][j]; }
                        }
                    }
                } else if (okk && fp) {
#pragma unroll
                    for (int c = 0; c < kTcCPT; ++c) {
                        const int kc = tc_core(c, q);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (kc * 8 + j < d) stV[sm_ + (kc * 8 + j) * kTcPiece] = -stV[sm_ + (kc * 8 + j) * kTcPiece];
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncthreads();                       // the stash is final; the tables, s_red and the A planes are free for the next tile
            TCX_MARK(5)
            {
                const bool rec = p.samples != nullptr;
                // Boxes that leave by TMA.  A box that reaches past this tile also rewrites the state the stash holds for
                // the first particles of the NEXT tile: unchanged values (or, in the first iteration, the input copied to
                // the output buffer, and a stale sample) that the next tile's own stores replace -- in that order, because
                // the helpers wait for their earlier stores before they issue new ones.  A box that would reach past the end
                // of the CTA's range (the neighbour's particles) is not sent: those columns go out with plain stores; so
                // does a tile with a failed particle (its sample is not recorded).
                int n_box = 0;
                if (maps.ok && !(rec && (any_fail || !maps.smp_ok)))
                    n_box = (int)min((long long)((np + kTcPiece - 1) / kTcPiece), (r1 - cur) / kTcPiece);   // whole boxes inside the range
                if (ht >= 0 && n_box > 0) {
                    bulk_wait_all();
                    helper_bar();
                    if (ht == 0) {
                        for (int piece = 0; piece < n_box; ++piece) {
                            const int c0 = (int)cur + piece * kTcPiece;
                            tma_box_s2g(&maps.xout, c0, stX + piece * (P * kTcPiece));
                            tma_box_s2g(&maps.vout, c0, stV + piece * (P * kTcPiece));
                            if (rec) tma_box_s2g_3d(&maps.smp, c0, it, stX + piece * (P * kTcPiece));
                        }
                    }
                    bulk_commit();
                }
                const bool bulk_out = n_box > 0;
                const int c_plain = n_box * kTcPiece;
                if (c_plain < np) {
                    // the boxes of the tile before may still be in flight, and they cover the first columns of this tile
                    if (ht >= 0) { bulk_wait_all(); asm volatile("fence.proxy.async;" ::: "memory"); }
                    __syncthreads();
                    const int w = np - c_plain;
                    for (int idx = tid; idx < d * w; idx += kTcThreads) {
                        const int k = idx / w, mm = c_plain + (idx - k * w);
                        const long long o = (long long)k * p.ld + cur + mm;
                        const float xv = stX[stash_at(mm) + k * kTcPiece];
                        Xo[o] = xv;
                        Vo[o] = stV[stash_at(mm) + k * kTcPiece];
                        if (rec && (s_code[mm] & 16u)) ((float*)p.samples)[(long long)k * p.s_stride_k + (long long)it * p.s_stride_it + cur + mm] = xv;
                    }
                    __syncthreads();
                }
                // ---- momentum refresh (hmc_state.py:121-129) of the R movers.  A few per tile (the continuous-time
                // samplers): left to the helper warps, beside the trajectory of the next tile.  Many (the batch-wide coin of
                // the discrete samplers): all threads share the Box-Muller pairs now (one thread per particle slice would
                // run the 16 pairs of its slice in warps where a single lane has an R move).
                const int nr = s_nr[tb];                                           // final since the barrier after the decisions
                if (nr > kTcDeferMax) {
                    if (bulk_out && ht >= 0) { bulk_wait_all(); asm volatile("fence.proxy.async;" ::: "memory"); }
                    __syncthreads();                                               // the momenta are in memory (partial refresh reads them)
                    refresh_jobs(s_rlist[tb], nr, cur, attempt, tid, kTcThreads);
                } else if (nr > 0) {
                    pend_nr = nr; pend_buf = tb; pend_cur = cur;
                }
                // ---- the helper warps fetch the next tile's boxes as soon as the stores have read the stash
                if (ht >= 0) {
                    if (bulk_out) bulk_wait_read();
                    helper_bar();
                }
                if (cur + np < r1) stash_bulk = stash_issue(it == 0, cur + np);
            }
            TCX_MARK(6)
#ifdef TCX_TRACE
            if (blockIdx.x == 0 && tile_no == 3 && tid == 0) {
                for (int pr = 0; pr < 12; ++pr) {
                    printf("st %2d:", pr + 4);
                    for (int e = 0; e < 13; ++e) printf(" %6lld", s_ev[pr][e] - s_ev[0][0]);
                    printf("\n");
                }
            }
            ++tile_no;
#endif
            cur += np;
            tb ^= 1;
        }
        // the last tile's R movers: before the next iteration (or the host) reads their momenta
        if (ht >= 0) { bulk_wait_all(); asm volatile("fence.proxy.async;" ::: "memory"); }
        __syncthreads();
        if (pend_nr > 0) {
            refresh_jobs(s_rlist[pend_buf], pend_nr, pend_cur, attempt, tid, kTcThreads);
            pend_nr = 0;
        }
        // this iteration's state (generic-proxy stores) is read by the bulk copies (async proxy) of the next one
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
    }
    if (p.n_iter == 0) {
        for (long long i = r0 + tid; i < r1; i += kTcThreads)
            for (int k = 0; k < d; ++k) {
                ((float*)p.Xout)[(long long)k * p.ld + i] = ((const float*)p.Xin)[(long long)k * p.ld + i];
                ((float*)p.Vout)[(long long)k * p.ld + i] = ((const float*)p.Vin)[(long long)k * p.ld + i];
            }
        if (mj) for (long long i = r0 + tid; i < r1; i += kTcThreads) {
            p.ca_out[i] = p.ca_in[i];
            ((float*)p.Hc_out)[i] = ((const float*)p.Hc_in)[i];
        }
    }

#ifdef TCX_TIMING
    if (blockIdx.x == 0 && (tid == 0 || tid == 130 || tid == 300 || tid == 512))
        printf("tid %d: plan %lld load %lld traj %lld sync %lld decide %lld apply %lld refresh %lld | w0 %lld w1 %lld w2 %lld w3 %lld issue %lld\n", tid, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tw[0], tw[1], tw[2], tw[3], tw[4]);
#endif
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTcTmemCols);
    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, n_E, n_exec};
    flush_counters(p.counters, loc, (unsigned long long)L);
}

bool dense_tc_supported(int kind, int ndims, int nbasis) {
    if (ndims < 1 || ndims > kTcMaxP) return false;
    if (kind == MJHMC_DIST_DENSE_GAUSSIAN) return true;
    return kind == MJHMC_DIST_PRODUCT_OF_T && nbasis == ndims;
}

static void tc_shape(int d, int& P, int& ncores) { P = ((d + 15) >> 4) << 4; ncores = P >> 3; }

long long dense_tc_workspace_bytes(int kind, int ndims) {
    int P, nc;
    tc_shape(ndims, P, nc);
    return 3ll * nc * nc * 128 + (kind == MJHMC_DIST_PRODUCT_OF_T ? (long long)kTcTabs * P * 4 : 0);
}

// Fills the pre-tiled bf16 planes (once per distribution; the sampler launches only read them).
cudaError_t dense_tc_prepare(int kind, const float* Mx, const float* nu, const float* b, int ndims, void* workspace,
                             cudaStream_t stream) {
    int P, nc;
    tc_shape(ndims, P, nc);
    const bool pot = kind == MJHMC_DIST_PRODUCT_OF_T;
    tc_prep_kernel<<<32, 256, 0, stream>>>(Mx, ndims, ndims, P, pot ? nu : nullptr, pot ? b : nullptr, (uint8_t*)workspace);
    return cudaGetLastError();
}

template <bool POT, int PC>
static cudaError_t launch_tc_T(const LaunchParams& p, cudaStream_t stream) {
    int P, nc;
    tc_shape(p.d, P, nc);
    const size_t smem = 3 * (size_t)nc * nc * 128 + (POT ? kTcTabs * P * 4 : 0) + 1024 + 128 + 2 * (size_t)P * kTcRows * 4;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel<POT, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (p.n + 95) / 96;            // ~one tile of jobs per CTA at the least
    if (blocks > sms) blocks = sms;
    if (blocks < 1) blocks = 1;
    // tensor maps of the state arrays and the sample record: (dims x 32 particles) boxes
    TcMaps maps;
    memset(&maps, 0, sizeof maps);
    {
        const long long dims[2] = {p.n, p.d}, strides[1] = {p.ld};
        const int box[2] = {kTcPiece, p.d};
        maps.ok = make_tensor_map_f32(&maps.xin, p.Xin, 2, dims, strides, box) && make_tensor_map_f32(&maps.vin, p.Vin, 2, dims, strides, box) &&
                  make_tensor_map_f32(&maps.xout, p.Xout, 2, dims, strides, box) && make_tensor_map_f32(&maps.vout, p.Vout, 2, dims, strides, box);
        maps.smp_ok = 1;
        if (maps.ok && p.samples && p.n_iter > 0) {
            const long long dims3[3] = {p.n, p.n_iter, p.d}, strides3[2] = {p.s_stride_it, p.s_stride_k};
            const int box3[3] = {kTcPiece, 1, p.d};
            maps.smp_ok = make_tensor_map_f32(&maps.smp, p.samples, 3, dims3, strides3, box3);
        }
    }
    dense_tc_kernel<POT, PC><<<(unsigned)blocks, kTcThreads, smem, stream>>>(p, maps);
    return cudaGetLastError();
}

cudaError_t launch_dense_tc(int kind, const LaunchParams& p, cudaStream_t stream) {
    if (!p.ws) return cudaErrorInvalidValue;                   // the pre-tiled matrix (mjhmc_dense_tc_prepare)
    int P, nc;
    tc_shape(p.d, P, nc);
    if (P == kTcMaxP)
        return kind == MJHMC_DIST_PRODUCT_OF_T ? launch_tc_T<true, kTcMaxP>(p, stream) : launch_tc_T<false, kTcMaxP>(p, stream);
    return kind == MJHMC_DIST_PRODUCT_OF_T ? launch_tc_T<true, 0>(p, stream) : launch_tc_T<false, 0>(p, stream);
}

}  // namespace mjhmc
