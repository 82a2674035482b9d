// K1s: streaming fused sampler for SEPARABLE energies (TestGaussian, diagonal Gaussian, RoughWell), any
// ndims <= 128.  Same arithmetic and same transition code as fused_elementwise.cuh; what changes is where the
// state lives and how it gets on chip:
//
//   * persistent CTAs walk particle TILES; the (ndims x P) X and V boxes of the next tiles are brought into a
//     shared-memory ring by TMA (cp.async.bulk.tensor.2d, one instruction per box, completion on an mbarrier)
//     while the current tile computes -- bytes in flight no longer depend on the register budget of the
//     integrator, which is what capped the register kernel at ~0.6 of the HBM roofline at L = 1;
//   * the dims of a separable energy evolve independently during the L leapfrog steps (hmc_state.py:86-91), only
//     the energy sums couple them, so G warps share one 32-particle column block, each thread owns DT dims of
//     its particle, and the six partial energies meet in shared memory once per iteration.  That lifts the
//     ndims <= 16 limit of the one-thread-per-particle kernel (reference default Gaussian(ndims=100) is diagonal:
//     distributions.py:257-263) without going through the dense-contraction kernel;
//   * the particle's current (x, v) stay in the ring stage across the n_iter iterations of one launch, the
//     proposal lives in registers; results leave with plain coalesced stores.
//
// Replaces the same reference code as fused_elementwise.cuh (hmc_state.py:86-129, markov_jump_hmc.py:116-148,
// :251-290, :355-415, utils.py:15-49).
#pragma once
#include <cuda.h>
#include "fused_elementwise.cuh"

namespace mjhmc {

constexpr int kStreamThreads = 256;
constexpr int kStreamWarps = kStreamThreads / 32;
#ifndef MJ_STREAM_MAX_STAGES
#define MJ_STREAM_MAX_STAGES 4
#endif
constexpr int kStreamMaxStages = MJ_STREAM_MAX_STAGES;
constexpr int kStreamRed = 3;              // partial sums per thread: H, H_L, H_FLF (this thread's dims)

struct StreamCfg {
    int G;                  // warps per 32-particle column block: 1, 2, 4 or 8
    int logG;
    int P;                  // particles per tile = 32 * kStreamWarps / G
    int stages;             // ring depth
    int use_tma;            // 0: cooperative loads (row stride or base not 16-byte aligned)
    long long n_tiles;
};

// Resident CTAs per SM the register allocator must allow: the live arrays are the proposal (x, v, g) of DT dims;
// everything else (Philox, exp, log) is out of line.  More CTAs = more warps to hide the dependent-issue latency of
// the per-particle chains; the TMA ring keeps the loads in flight whatever the register budget is.
template <typename T, int DT, bool LINEAR, bool MJ = false>
__host__ __device__ constexpr int stream_min_blocks() {
    // measured on B200 (profiles/r1_stream_*.txt): at 2 dims per thread 4 CTAs with a few spilled
    // bytes win over 3; energies with a folded linear kick keep no gradient registers and go one size class
    // further; the MarkovJumpHMC body (FLF cache, three holding times) spills at 13 fp64 dims per thread under the
    // 128-register cap of two CTAs and is 8 % faster as one CTA with a 4-stage ring.
    constexpr int bytes = DT * (int)sizeof(T);
    if (LINEAR) return bytes <= 16 ? 4 : (bytes <= 32 ? 3 : (bytes <= (MJ ? 103 : 104) ? 2 : 1));
    return bytes <= 16 ? 4 : (bytes <= 32 ? 3 : (bytes <= 80 ? 2 : 1));
}

namespace tma {
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
        "SMBAR_WAIT:\n"
#ifdef MJ_STREAM_WAIT_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
        "@p bra SMBAR_DONE;\n"
        "bra SMBAR_WAIT;\n"
        "SMBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one (P x ndims) box of a (ndims, n) row-major array: coordinate 0 = first particle, coordinate 1 = first dim
__device__ __forceinline__ void load_box(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
}  // namespace tma

// ---------------------------------------------------------------- energies as the streaming kernel sees them
// Quadratic energies have a LINEAR gradient g_k = j_k x_k, so the kick of hmc_state.py:88,91 folds into one FMA per
// dim, v_k += (-eps j_k) x_k for the two adjacent half kicks between leapfrog steps, -eps/2 j_k at both ends:
// 2 fp64 FMAs per dim and step instead of 4 and no gradient registers.  The folded coefficient is rounded once
// ((eps j) x instead of eps (j x)): a relative 1e-16 per step, inside the fp64 tolerance.
template <typename T, int D>
struct LinTestGaussian {                         // distributions.py:357-362
    static constexpr bool kLinear = true;
    T cf1, inv_2s2;
    __device__ __forceinline__ LinTestGaussian(const LaunchParams& p, int k0, int nd, T*)
        : cf1((T)(-p.eps / (p.dp[0] * p.dp[0]))), inv_2s2((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))) {}
    __device__ __forceinline__ T cf(int) const { return cf1; }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) s += x[k] * x[k];
        return s * inv_2s2;
    }
};

template <typename T, int D>
struct LinDiagGaussian {                         // distributions.py:262-273 with J = diag(j)
    static constexpr bool kLinear = true;
    // -eps * j_k and j_k of this thread's dims (0 on padding dims) in shared memory: the compiler keeps them in
    // registers where it has room and re-reads them (one LDS beside two DFMAs) where it has not -- at 13 dims
    // per thread a register copy spilled and cost 1.6x
    const T* c;
    const T* jj;
    __device__ __forceinline__ LinDiagGaussian(const LaunchParams& p, int k0, int nd, T* scratch)
        : c(scratch), jj(scratch + D) {
        if ((threadIdx.x & 31) == 0) {
            for (int k = 0; k < D; ++k) {
                const T j = k < nd ? ((const T*)p.a0)[k0 + k] : (T)0;
                scratch[k] = (T)(-p.eps) * j;
                scratch[D + k] = j;
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ T cf(int k) const { return c[k]; }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) s += x[k] * (jj[k] * x[k]);
        return s * (T)0.5;
    }
};

template <class Dist, typename T>
__device__ __forceinline__ Dist make_stream_dist(const LaunchParams& p, int k0, int nd, T* scratch) {
    if constexpr (Dist::kLinear) return Dist(p, k0, nd, scratch);
    else return Dist(p, k0, nd);
}

// One trajectory of L leapfrog steps (hmc_state.py:93-100) on this thread's dims; returns the partial energies.
// want_e0: also the potential energy at the starting point (first iteration of a launch), evaluated together with
// the first gradient where the energy shares work with it (dists.cuh: grad_aux).
template <class Dist, typename T, int DT>
__device__ __forceinline__ void stream_trajectory(const Dist& dist, T (&x)[DT], T (&v)[DT], T eps, T neg_half_eps,
                                                  int L, T& e_pot, T& e_kin, bool want_e0, T& e0) {
    if constexpr (Dist::kLinear) {
        if (want_e0) e0 = dist.energy(x);
        if (L > 0) {
#pragma unroll
            for (int k = 0; k < DT; ++k) v[k] += ((T)0.5 * dist.cf(k)) * x[k];
            for (int s = 1; s < L; ++s) {
#pragma unroll
                for (int k = 0; k < DT; ++k) x[k] += eps * v[k];
#pragma unroll
                for (int k = 0; k < DT; ++k) v[k] += dist.cf(k) * x[k];
            }
#pragma unroll
            for (int k = 0; k < DT; ++k) x[k] += eps * v[k];
#pragma unroll
            for (int k = 0; k < DT; ++k) v[k] += ((T)0.5 * dist.cf(k)) * x[k];
        }
        e_pot = dist.energy(x);
    } else {
        T g[DT];
        if (want_e0) {
            const T aux0 = grad_aux<Dist, T, DT>(dist, x, g);
            e0 = energy_after<Dist, T, DT>(dist, x, aux0);
        } else {
            dist.grad(x, g);
        }
        const T aux = leapfrog_L<Dist, T, DT>(dist, x, v, g, eps, neg_half_eps, L);
        e_pot = energy_after<Dist, T, DT>(dist, x, aux);
    }
    e_kin = kinetic<T, DT>(v);
}

template <class Dist, typename T, int DT, int SAMPLER, int LOGG>
__global__ void __launch_bounds__(kStreamThreads, stream_min_blocks<T, DT, Dist::kLinear, SAMPLER == MJHMC_SAMPLER_MARKOV_JUMP>())
stream_sample_kernel(const __grid_constant__ LaunchParams p, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ CUtensorMap tmV, const StreamCfg cfg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];   // TMA without swizzle: 128-byte aligned boxes
    constexpr bool MJ = SAMPLER == MJHMC_SAMPLER_MARKOV_JUMP;
    constexpr bool CT = SAMPLER == MJHMC_SAMPLER_CONTINUOUS_TIME;
    constexpr bool DISCRETE = SAMPLER == MJHMC_SAMPLER_DISCRETE;
    // A stage holds the X box and the V box of one tile: G * DT rows (the dims, zero rows above ndims: the
    // tensor map fills out-of-range rows and columns with zeros) of P particles; G * P = 256.
    constexpr unsigned kBoxElems = (unsigned)kStreamThreads * DT;
    constexpr unsigned kStageBytes = 2u * kBoxElems * (unsigned)sizeof(T);

    // warps per 32-particle column block and particles per tile are compile-time: every shared-memory row offset
    // j * P below is an immediate (with a run-time P the address arithmetic was a quarter of the per-tile code)
    constexpr int G = 1 << LOGG, P = kStreamThreads / G;
    const int d = p.d, stages = cfg.stages;
    T* const ring = (T*)smem_raw;
    unsigned char* const fixed = smem_raw + (size_t)stages * kStageBytes;
    int* const coin = (int*)fixed;                                   // [32] batch-wide R coins (discrete samplers)
    uint64_t* const full = (uint64_t*)(coin + 32);                   // [kStreamMaxStages]
    unsigned int* const done = (unsigned int*)(full + kStreamMaxStages);   // [kStreamMaxStages] warps done with a stage
    unsigned char* const sel = (unsigned char*)(done + kStreamMaxStages);  // [256] decision broadcast   (G > 1)
    T* const coef = (T*)(sel + kStreamThreads);                      // [kStreamWarps][2 * DT] per-warp coefficients
    T* const red = coef + kStreamWarps * 2 * DT;                     // [kStreamRed][G][P]               (G > 1)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cb = warp >> LOGG, g = warp & (G - 1);
    const int c = cb * 32 + lane;                  // particle column inside the tile
    const int k0 = g * DT;                         // first dim of this thread
    const int nd = max(0, min(DT, d - k0));        // dims of this thread
    // the warp with the fewest dims of the column block takes the transition
    const bool decider = (g == G - 1);
    const int off0 = k0 * P + c;                   // this thread's first element inside an X (or V) box

    const Dist dist = make_stream_dist<Dist, T>(p, k0, nd, coef + warp * 2 * DT);
    const T eps = (T)p.eps;
    const T neg_half_eps = (T)(-p.eps / 2.0);
    const int L = p.L, n_iter = p.n_iter;
    const bool literal_race = (p.rng_flags & MJHMC_RNG_FLAG_LITERAL_RACE) != 0;
    const long long n = p.n, ld = p.ld;
    const bool use_tma = cfg.use_tma != 0;
    // G == 1: nothing couples the warps of a tile, so the stage is handed back by the last warp that leaves it
    // (shared counter) instead of a CTA barrier and the warps drift freely over the ring.
    const bool warp_release = (G == 1) && use_tma;

    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0, n_E = 0, n_exec = 0;

    if (tid == 0) {
        for (int s = 0; s < kStreamMaxStages; ++s) { tma::mbar_init(&full[s], 1); done[s] = 0; }
        tma::fence_mbar_init();
    }
    __syncthreads();
    if (use_tma && tid == 0) {
        long long t = blockIdx.x;
        for (int s = 0; s < stages && t < cfg.n_tiles; ++s, t += gridDim.x) {
            T* sx = ring + (size_t)s * 2 * kBoxElems;
            tma::mbar_expect_tx(&full[s], kStageBytes);
            tma::load_box(sx, &tmX, (int)(t * P), 0, &full[s]);
            tma::load_box(sx + kBoxElems, &tmV, (int)(t * P), 0, &full[s]);
        }
    }

    int s = 0;                      // ring stage of the current tile
    uint32_t phase = 0;             // parity of the fill the current tile waits for
    bool first_tile = true;
    unsigned int nx_cflags = 0;
    T nx_Hc = (T)0;
    if (MJ) {
        const long long i0 = (long long)blockIdx.x * P + c;
        if (i0 < n) {
            nx_cflags = p.ca_in[i0];
            if (decider) nx_Hc = ((const T*)p.Hc_in)[i0];
        }
    }
    for (long long tile = blockIdx.x; tile < cfg.n_tiles; tile += gridDim.x) {
        T* const sx = ring + (size_t)s * 2 * kBoxElems;
        T* const bx = sx + off0;                   // this thread's column: dim j at bx[j * P], momentum at bv[j * P]
        T* const bv = bx + kBoxElems;
        if (use_tma) {
            tma::mbar_wait(&full[s], phase);
        } else {
            const T* Xin = (const T*)p.Xin;
            const T* Vin = (const T*)p.Vin;
            for (int k = 0; k < G * DT; ++k)
                for (int cc = tid; cc < P; cc += kStreamThreads) {
                    const long long gi = tile * P + cc;
                    const bool in = gi < n && k < d;
                    sx[k * P + cc] = in ? Xin[(long long)k * ld + gi] : (T)0;
                    sx[kBoxElems + k * P + cc] = in ? Vin[(long long)k * ld + gi] : (T)0;
                }
            __syncthreads();
        }

        const long long i = tile * P + c;
        const bool live = i < n;
        const int nst = live ? nd : 0;             // rows this thread stores
        // FLF cache of this tile (loaded one tile ahead: a plain global load here would expose a DRAM round trip)
        unsigned int cflags = nx_cflags;
        T Hc = nx_Hc;
        if (MJ) {
            const long long tn = tile + gridDim.x;
            const long long in = tn * P + c;
            if (tn < cfg.n_tiles && in < n) {
                nx_cflags = p.ca_in[in];
                if (decider) nx_Hc = ((const T*)p.Hc_in)[in];
            }
        }
        double dwell = 0.0;
        bool failed = false;
        T ex = (T)0, ev = (T)0;                    // partial energies of the current state (this thread's dims)
        T xt[DT], vt[DT];

        if (n_iter == 0) {
            T* gx = (T*)p.Xout + i + (long long)k0 * ld;
            T* gv = (T*)p.Vout + i + (long long)k0 * ld;
#pragma unroll
            for (int j = 0; j < DT; ++j) {
                if (j < nst) { *gx = bx[j * P]; *gv = bv[j * P]; }
                gx += ld; gv += ld;
            }
        }

        for (int it = 0; it < n_iter; ++it) {
            const unsigned long long attempt = p.attempt0 + (unsigned long long)it;
            const bool last = it + 1 == n_iter;
            const bool active = live && !failed;

            if (DISCRETE && (it & 31) == 0 && (n_iter > 32 || first_tile)) {
                __syncthreads();
                if (warp == 0 && it + lane < n_iter)
                    coin[lane] = draw_coin(p, attempt + (unsigned long long)lane) < p.p_r;   // markov_jump_hmc.py:138
                __syncthreads();
            }

            // the draws of this attempt do not depend on the trajectory: the Philox rounds run on the integer pipe
            // before the energies are known
            double u0 = 0.0, u1 = 0.0, u2 = 0.0;
            RaceDraws rd = {};
            if (decider && active) {
                const bool need_u2 = !DISCRETE && p.p_r != 0.0;
                const Uniform3 u = draw_uniforms(p, i, attempt, need_u2);
                u0 = u.u0; u1 = u.u1; u2 = u.u2;
                if (!DISCRETE) rd = race_draws(p.p_r, u0, u1, u2, literal_race);
            }
            const bool need_dwell = !DISCRETE && (p.dwell != nullptr || (last && p.dwell_last != nullptr));

            T exf = (T)0, evf = (T)0, exl = (T)0, evl = (T)0;
            if (active) {
#pragma unroll
                for (int j = 0; j < DT; ++j) { xt[j] = bx[j * P]; vt[j] = bv[j * P]; }
                if (it == 0) ev = kinetic<T, DT>(vt);
                T e0 = (T)0;
                // ---- FLF state (hmc_state.py:109-119): only its energy is ever read
                if (MJ && !(cflags & kCacheValid)) {
#pragma unroll
                    for (int j = 0; j < DT; ++j) vt[j] = -vt[j];
                    stream_trajectory<Dist, T, DT>(dist, xt, vt, eps, neg_half_eps, L, exf, evf, false, e0);
#pragma unroll
                    for (int j = 0; j < DT; ++j) { xt[j] = bx[j * P]; vt[j] = bv[j * P]; }
                }
                // ---- L state (hmc_state.py:93-100); the energy of the starting point with its first gradient
                stream_trajectory<Dist, T, DT>(dist, xt, vt, eps, neg_half_eps, L, exl, evl, it == 0, e0);
                if (it == 0) ex = e0;
            }

            // ---- the partial energies of the G threads of a particle meet here
            T H = ex + ev, Hl = exl + evl, Hf = exf + evf;         // hmc_state.py:80-84
            if (G > 1) {
                T* r = red + g * P + c;
                r[0 * kStreamThreads] = H;
                r[1 * kStreamThreads] = Hl;
                if (MJ) r[2 * kStreamThreads] = Hf;
                __syncthreads();
                if (decider) {
                    H = Hl = Hf = (T)0;
                    for (int gg = 0; gg < G; ++gg) {              // fixed order: shard- and schedule-independent
                        const T* rr = red + gg * P + c;
                        H += rr[0 * kStreamThreads];
                        Hl += rr[1 * kStreamThreads];
                        if (MJ) Hf += rr[2 * kStreamThreads];
                    }
                }
            }

            // ---- transition (decider thread of the particle)
            unsigned int choice = 0;
            if (decider && active) {
                n_E += 1;
                n_exec += 1;
                if (MJ) {
                    T Hflf = Hc;
                    if (!(cflags & kCacheRef)) n_E += 1;           // the reference evaluates the FLF state here
                    if (!(cflags & kCacheValid)) { Hflf = Hf; n_exec += 1; }
                    const Decision dc = decide_mj_s(p.p_r, u0, u1, u2, rd, (double)(H - Hl), (double)(H - Hflf), need_dwell);
                    if (dc.fail) { report_failure(p, it); failed = true; }
                    else {
                        choice = dc.choice; dwell = dc.dwell;
                        if (choice == 0) { Hc = H; cflags = kCacheRef | kCacheValid; n_l += 1; }      // :399
                        else if (choice == 1) { Hc = Hl; cflags = kCacheValid; n_f += 1; }            // :410
                        else { cflags = 0; n_r += 1; }                                                // :409
                    }
                } else if (CT) {
                    const Decision dc = decide_ct_s(p.p_r, u0, u1, u2, rd, (double)(H - Hl), need_dwell);
                    if (dc.fail) { report_failure(p, it); failed = true; }
                    else {
                        choice = dc.choice; dwell = dc.dwell;
                        if (choice == 1) n_fl += 1; else if (choice == 0) n_f += 1; else n_r += 1;
                    }
                } else {
                    choice = decide_discrete_s(p.p_flip, u0, u1, (double)(H - Hl), coin[it & 31] != 0);
                    const bool acc = choice & 1u, flip = choice & 2u;
                    n_l += (acc && flip);
                    n_f += (flip && !acc);
                    n_fl += (acc && !flip);
                    if (choice & 4u) n_r += 1;
                }
            }
            if (G > 1) {
                if (decider) sel[c] = (unsigned char)(choice | (failed ? 8u : 0u) | (cflags << 4));
                __syncthreads();
                const unsigned int sc = sel[c];
                choice = sc & 7u; failed = (sc & 8u) != 0; cflags = sc >> 4;
            }

            // ---- apply the operator to this thread's dims (xt, vt become the new state)
            if (live) {
                bool moved, negate, refresh;
                if (MJ) { moved = choice == 0; negate = choice == 1; refresh = choice == 2; }
                else if (CT) { moved = choice == 1; negate = choice != 2; refresh = choice == 2; }   // F L z / F z
                else {
                    const bool acc = choice & 1u, flip = choice & 2u;
                    moved = acc; negate = acc != flip; refresh = (choice & 4u) != 0;   // F L z, then the flip
                }
                if (!active || failed) { moved = false; negate = false; refresh = false; }
                if (moved) { ex = exl; ev = evl; }
                else {
#pragma unroll
                    for (int j = 0; j < DT; ++j) { xt[j] = bx[j * P]; vt[j] = bv[j * P]; }
                }
                if (negate) {
#pragma unroll
                    for (int j = 0; j < DT; ++j) vt[j] = -vt[j];
                }
                if (refresh) {
                    double z0 = 0.0, z1 = 0.0;
#pragma unroll
                    for (int j = 0; j < DT; ++j) {
                        const int k = k0 + j;
                        if (j < nd) {
                            if (!(k & 1) || j == 0) normal_pair(p, i, attempt, k >> 1, d, z0, z1);
                            const T z = (T)((k & 1) ? z1 : z0);
                            vt[j] = vt[j] * (T)p.r_keep + z * (T)p.r_mix;           // hmc_state.py:126
                        }
                    }
                    ev = kinetic<T, DT>(vt);
                }
                if (active && !failed) {
                    if (p.samples) {                                   // record (markov_jump_hmc.py:169,334)
                        T* gs = (T*)p.samples + (long long)it * p.s_stride_it + i + (long long)k0 * p.s_stride_k;
#pragma unroll
                        for (int j = 0; j < DT; ++j) {
                            if (j < nst) *gs = xt[j];
                            gs += p.s_stride_k;
                        }
                    }
                    if (decider) {
                        if (!DISCRETE && p.dwell) p.dwell[(long long)it * n + i] = dwell;
                        if (p.choice) p.choice[(long long)it * n + i] = (uint8_t)choice;
                    }
                }
                if (!last) {
                    if (active && !failed) {
#pragma unroll
                        for (int j = 0; j < DT; ++j) { bx[j * P] = xt[j]; bv[j * P] = vt[j]; }
                    }
                } else {
                    T* gx = (T*)p.Xout + i + (long long)k0 * ld;
                    T* gv = (T*)p.Vout + i + (long long)k0 * ld;
#pragma unroll
                    for (int j = 0; j < DT; ++j) {
                        if (j < nst) { *gx = xt[j]; *gv = vt[j]; }
                        gx += ld; gv += ld;
                    }
                }
            }
        }

        if (live && decider) {
            if (MJ) {
                p.ca_out[i] = (uint8_t)cflags;
                ((T*)p.Hc_out)[i] = Hc;
            }
            if (!DISCRETE && p.dwell_last) p.dwell_last[i] = dwell;
        }

        // ---- the stage is free: refill it with the tile `stages` steps ahead
        if (n_iter > 1) tma::fence_proxy_async();      // generic-proxy writes to the stage precede the TMA refill
        bool refill;
        if (warp_release) {
            __syncwarp();
            refill = false;
            if (lane == 0) {
                __threadfence_block();
                refill = (atomicAdd(&done[s], 1u) & (kStreamWarps - 1)) == kStreamWarps - 1;   // the last warp out
            }
        } else {
            __syncthreads();
            refill = use_tma && tid == 0;
        }
        if (refill) {
            const long long t = tile + (long long)stages * gridDim.x;
            if (t < cfg.n_tiles) {
                tma::mbar_expect_tx(&full[s], kStageBytes);
                tma::load_box(sx, &tmX, (int)(t * P), 0, &full[s]);
                tma::load_box(sx + kBoxElems, &tmV, (int)(t * P), 0, &full[s]);
            }
        }
        if (++s == stages) { s = 0; phase ^= 1u; }
        first_tile = false;
    }

    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, n_E, n_exec};
    flush_counters(p.counters, loc, (unsigned long long)L);
}

// ---------------------------------------------------------------- host side
struct StreamPlan { int G, logG, DT; };

// ndims -> (warps per particle column, dims per thread); DT == 0: not supported.
// One thread per particle up to 16 dims (no barrier inside a tile); above that the smallest power of two of
// warps that brings the dims per thread to <= 16.
inline StreamPlan stream_plan(int d) {
    static const int dts[] = {2, 4, 6, 8, 10, 13, 16};
    StreamPlan pl{8, 3, 0};
    if (d <= 0 || d > 128) return pl;
    for (int lg = 0; lg <= 3; ++lg) {
        const int G = 1 << lg, per = (d + G - 1) / G;
        if (per <= 16) {
            pl.G = G; pl.logG = lg;
            for (int t : dts) if (per <= t) { pl.DT = t; break; }
            break;
        }
    }
    return pl;
}

typedef cudaError_t (*stream_launch_fn)(const LaunchParams&, const StreamPlan&, int dtype, cudaStream_t);

void stream_note_launch(int use_tma, int stages, long long grid, int per_sm, int G, int DT, size_t smem);
cudaError_t stream_make_maps(const LaunchParams& p, int dtype, int P, int rows, CUtensorMap* mx, CUtensorMap* mv,
                             int* use_tma);
int stream_sm_count();

template <class Dist, typename T, int DT, int SAMPLER, int LOGG>
cudaError_t launch_stream_g(const LaunchParams& p, const StreamPlan& pl, int dtype, cudaStream_t stream) {
    StreamCfg cfg;
    cfg.G = pl.G; cfg.logG = pl.logG;
    cfg.P = 32 * kStreamWarps / pl.G;
    cfg.n_tiles = (p.n + cfg.P - 1) / cfg.P;
    if (cfg.n_tiles == 0) return cudaSuccess;
    CUtensorMap mx, mv;
    cudaError_t e = stream_make_maps(p, dtype, cfg.P, pl.G * DT, &mx, &mv, &cfg.use_tma);
    if (e != cudaSuccess) return e;
    const size_t stage_bytes = (size_t)2 * kStreamThreads * DT * sizeof(T);      // G * DT rows of P particles, X and V
    const size_t fixed = 32 * sizeof(int) + kStreamMaxStages * (sizeof(uint64_t) + sizeof(unsigned int)) + kStreamThreads +
                         (size_t)kStreamWarps * 2 * DT * sizeof(T) +
                         (pl.G > 1 ? (size_t)kStreamRed * kStreamThreads * sizeof(T) : 0);
    // as many CTAs per SM as the registers allow, but at least two ring stages each
    int nb = stream_min_blocks<T, DT, Dist::kLinear, SAMPLER == MJHMC_SAMPLER_MARKOV_JUMP>(), stages = 0;
    for (; nb >= 1; --nb) {
        // 228 KB of shared memory per SM, 1 KB of it reserved per resident CTA (probed on B200 with
        // mjhmc_stream_probe_blocks: 115712 dynamic bytes is the most two CTAs can have each); 512 bytes of slack
        // for the static shared memory of flush_counters
        const size_t budget = (size_t)233472 / nb - 1024 - 512 - fixed;
        stages = (int)(budget / stage_bytes);
        if (stages >= 2 || nb == 1) break;
    }
    if (stages < 1) return cudaErrorInvalidConfiguration;
    if (stages > kStreamMaxStages) stages = kStreamMaxStages;
    if (!cfg.use_tma) stages = 1;
    cfg.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    auto kern = stream_sample_kernel<Dist, T, DT, SAMPLER, LOGG>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStreamThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    long long grid = (long long)per_sm * stream_sm_count();
    if (grid > cfg.n_tiles) grid = cfg.n_tiles;
    stream_note_launch(cfg.use_tma, stages, grid, per_sm, pl.G, DT, smem);
    kern<<<(unsigned)grid, kStreamThreads, smem, stream>>>(p, mx, mv, cfg);
    return cudaGetLastError();
}

// several warps per particle only exist for ndims > 16, i.e. at 10, 13 or 16 dims per thread (stream_plan)
template <class Dist, typename T, int DT, int SAMPLER>
cudaError_t launch_stream_s(const LaunchParams& p, const StreamPlan& pl, int dtype, cudaStream_t stream) {
    if (pl.logG == 0) return launch_stream_g<Dist, T, DT, SAMPLER, 0>(p, pl, dtype, stream);
    if constexpr (DT >= 10) {
        if (pl.logG == 1) return launch_stream_g<Dist, T, DT, SAMPLER, 1>(p, pl, dtype, stream);
        if (pl.logG == 2) return launch_stream_g<Dist, T, DT, SAMPLER, 2>(p, pl, dtype, stream);
        if (pl.logG == 3) return launch_stream_g<Dist, T, DT, SAMPLER, 3>(p, pl, dtype, stream);
    }
    return cudaErrorInvalidConfiguration;
}

template <class Dist, typename T, int DT>
cudaError_t launch_stream(const LaunchParams& p, const StreamPlan& pl, int dtype, cudaStream_t stream) {
    switch (p.sampler) {
        case MJHMC_SAMPLER_DISCRETE:
            return launch_stream_s<Dist, T, DT, MJHMC_SAMPLER_DISCRETE>(p, pl, dtype, stream);
        case MJHMC_SAMPLER_CONTINUOUS_TIME:
            return launch_stream_s<Dist, T, DT, MJHMC_SAMPLER_CONTINUOUS_TIME>(p, pl, dtype, stream);
        default:
            return launch_stream_s<Dist, T, DT, MJHMC_SAMPLER_MARKOV_JUMP>(p, pl, dtype, stream);
    }
}

template <typename T, int DT>
stream_launch_fn pick_stream_dist(int kind) {
    switch (kind) {
        case MJHMC_DIST_TEST_GAUSSIAN: return &launch_stream<LinTestGaussian<T, DT>, T, DT>;
        case MJHMC_DIST_DIAG_GAUSSIAN: return &launch_stream<LinDiagGaussian<T, DT>, T, DT>;
        case MJHMC_DIST_ROUGH_WELL:    return &launch_stream<RoughWellD<T, DT>, T, DT>;
        default: return nullptr;
    }
}

}  // namespace mjhmc
