// K1+K2+K3: fused leapfrog integrator + transition for register-resident energies.
//
// One thread owns one particle.  Its position, momentum and gradient stay in
// registers across the L leapfrog steps AND across the n_iter sampling iterations
// of one launch; HBM sees one coalesced read of (X, V) at launch start, one
// coalesced write per iteration of the recorded sample column (+ dwelling time),
// and one write of (X, V) at the end.  EX, EV and dEdX of the reference state
// (samplers/hmc_state.py:28-39) are functions of (X, V) and are recomputed on chip.
//
// Replaces, per iteration:
//   HMCState.leapfrog/L/F/FLF/R            samplers/hmc_state.py:86-129
//   HMCBase.sampling_iteration             samplers/markov_jump_hmc.py:116-148
//   ContinuousTimeHMC.sampling_iteration   samplers/markov_jump_hmc.py:251-290
//   MarkovJumpHMC.sampling_iteration       samplers/markov_jump_hmc.py:355-415
//   draw_from / min_idx                    misc/utils.py:15-49
#pragma once
#include "dists.cuh"

namespace mjhmc {

constexpr int kFusedThreads = 128;       // particles per CTA (one thread each)

// Resident CTAs per SM the register allocator must allow (measured on B200, profiles/r1_variants.md):
// small ndims want occupancy (the leapfrog loop is a latency-bound fp64 chain), large ndims want
// registers (x, v, g and the proposal are 6*D live values).
template <typename T, int D>
__host__ __device__ constexpr int fused_min_blocks() {
#ifdef MJ_MINBLOCKS_OVERRIDE
    return MJ_MINBLOCKS_OVERRIDE;
#else
    return sizeof(T) == 8 ? (D <= 4 ? 6 : (D <= 10 ? 3 : 2)) : (D <= 4 ? 8 : (D <= 10 ? 6 : 3));
#endif
}

template <typename T, int D>
__device__ __forceinline__ T kinetic(const T (&v)[D]) {
    T s = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) s += v[k] * v[k];
    return s * (T)0.5;                       // hmc_state.py:49-50
}

// hmc_state.py:86-100 -- L leapfrog steps; g enters as dEdX(x) and leaves as dEdX(x').  Returns what the last
// gradient evaluation shares with the energy at x' (energy_after, dists.cuh).
template <class Dist, typename T, int D>
__device__ __forceinline__ T leapfrog_L(const Dist& dist, T (&x)[D], T (&v)[D], T (&g)[D],
                                        T eps, T neg_half_eps, int L) {
    T aux;
    if (L <= 0) { T g0[D]; return grad_aux<Dist, T, D>(dist, x, g0); }
#ifndef MJ_SPLIT_KICKS
    // the closing half kick of step s and the opening half kick of step s+1 use the same gradient: one full kick
    // (v - eps g instead of (v - eps/2 g) - eps/2 g: one rounding fewer, one fp64 FMA per dim and step fewer;
    // +5 % on the Funnel, neutral on RoughWell whose loop is the sine; -DMJ_SPLIT_KICKS restores the literal form)
    const T neg_eps = neg_half_eps + neg_half_eps;
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] += neg_half_eps * g[k];
    for (int s = 1; s < L; ++s) {
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] += eps * v[k];
        dist.grad(x, g);
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] += neg_eps * g[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] += eps * v[k];
    aux = grad_aux<Dist, T, D>(dist, x, g);
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] += neg_half_eps * g[k];
#else
    for (int s = 0; s < L; ++s) {
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] += neg_half_eps * g[k];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] += eps * v[k];
        aux = grad_aux<Dist, T, D>(dist, x, g);
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] += neg_half_eps * g[k];
    }
#endif
    return aux;
}

// Momentum refresh v = v sqrt(1 - beta) + z sqrt(beta) (hmc_state.py:121-129) for the lanes with `mine` set; called by
// the whole warp.  PHILOX mode: the Box-Muller pairs of the refreshing lanes are shared out over all 32 lanes (job =
// (refreshing lane, pair)) and handed back through shared memory -- one lane in seven takes an R move on the Funnel
// of BASELINE config 5, and with every lane drawing its own D / 2 pairs the warp ran five Philox + log + sqrt +
// sincospi passes at 15 % occupancy: 39 % of all instructions (profiles/r2_fused_funnel10d_cthmc_v0.txt).
// Same counters, same arithmetic, same results as draw_normals.  zs: this CTA's [D][kFusedThreads] staging array.
template <typename T, int D>
__device__ __forceinline__ void refresh_momentum(const LaunchParams& p, long long i, unsigned long long attempt, int d,
                                                 bool mine, T (&v)[D], T (*zs)[128]) {
    const unsigned mask = __ballot_sync(0xffffffffu, mine);
    if (mask == 0u) return;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    if (D <= 2 || p.rng_mode == MJHMC_RNG_INJECT) {          // one pair per lane: nothing to share out
        if (mine) {
            T z[D];
            draw_normals<T, D>(p, i, attempt, d, z);
#pragma unroll
            for (int k = 0; k < D; ++k) v[k] = v[k] * (T)p.r_keep + z[k] * (T)p.r_mix;
        }
        return;
    }
    const int npairs = (d + 1) >> 1, njobs = __popc(mask) * npairs;
    for (int job = lane; job < njobs; job += 32) {
        const int r = job / npairs, pr = job - r * npairs;
        const int owner = __fns(mask, 0, r + 1);                   // lane of the r-th refreshing particle
        double z0, z1;
        normal_pair(p, i - lane + owner, attempt, pr, d, z0, z1);
        zs[2 * pr][wbase + owner] = (T)z0;
        if (2 * pr + 1 < D) zs[2 * pr + 1][wbase + owner] = (2 * pr + 1 < d) ? (T)z1 : (T)0;
    }
    __syncwarp();
    if (mine) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const T z = k < d ? zs[k][threadIdx.x] : (T)0;
            v[k] = v[k] * (T)p.r_keep + z * (T)p.r_mix;
        }
    }
    __syncwarp();
}

// FLF cache flags (one byte per particle).  bit0 is the reference's cache_active
// (hmc_state.py:41-44,131-148: set by an L move, cleared by F and R moves).  bit1 says the
// cached energy is valid; it is also set by an F move, because the FLF state of F z is F L z,
// whose energy is the H_L just computed for z -- the bitwise-identical trajectory the reference
// re-integrates at the next iteration (its own commented-out line markov_jump_hmc.py:400-401).
// The evaluation counters follow bit0, i.e. they count what the reference would evaluate; the
// trajectories actually integrated are counted separately (MJHMC_CNT_EXEC).
constexpr unsigned kCacheRef = 1u, kCacheValid = 2u;

template <class Dist, typename T, int D>
__global__ void __launch_bounds__(kFusedThreads, fused_min_blocks<T, D>())
fused_sample_kernel(const __grid_constant__ LaunchParams p) {
    __shared__ int s_coin[2];          // batch-wide R coin of the discrete samplers, one draw per CTA
    __shared__ T s_z[D][kFusedThreads];  // normals on their way from the lane that drew them to the lane that owns them

    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < p.n;
    const Dist dist(p);
    const int d = p.d;
    const T eps = (T)p.eps;
    const T neg_half_eps = (T)(-p.eps / 2.0);
    const int L = p.L;
    const int sampler = p.sampler;

    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0, n_E = 0, n_exec = 0;

    T x[D], v[D], g[D];
    {
        const T* Xin = (const T*)p.Xin;
        const T* Vin = (const T*)p.Vin;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            x[k] = (live && k < d) ? Xin[(long long)k * p.ld + i] : (T)0;
            v[k] = (live && k < d) ? Vin[(long long)k * p.ld + i] : (T)0;
        }
    }
    const T aux0 = grad_aux<Dist, T, D>(dist, x, g);
    T EX = energy_after<Dist, T, D>(dist, x, aux0);
    T EV = kinetic<T, D>(v);

    unsigned int cflags = 0;
    T Hc = (T)0;
    if (live && sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
        cflags = p.ca_in[i];
        Hc = ((const T*)p.Hc_in)[i];
    }
    double dwell = 0.0;
    bool failed = false;

    for (int it = 0; it < p.n_iter; ++it) {
        const unsigned long long attempt = p.attempt0 + (unsigned long long)it;
        const bool active = live && !failed;
        bool need_r = false;
        unsigned int choice = 0;
        // the holding time is stored per iteration (dwelling-time record) or only after the last one (dwell_last)
        const bool need_dwell = p.dwell != nullptr || (it + 1 == p.n_iter && p.dwell_last != nullptr);
        const T H = EX + EV;                                   // hmc_state.py:80-84
        T xt[D], vt[D], gt[D];
        T Hflf = Hc;
        if (sampler == MJHMC_SAMPLER_DISCRETE) {
            if (threadIdx.x == 0) s_coin[it & 1] = draw_coin(p, attempt) < p.p_r;   // :138
            __syncthreads();
        }

        if (active) {
            // ---- FLF state (hmc_state.py:109-119): only its energy is ever read
            if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
                if (!(cflags & kCacheRef)) n_E += 1;           // the reference evaluates it here
                if (!(cflags & kCacheValid)) {
#pragma unroll
                    for (int k = 0; k < D; ++k) { xt[k] = x[k]; vt[k] = -v[k]; gt[k] = g[k]; }
                    const T aux = leapfrog_L<Dist, T, D>(dist, xt, vt, gt, eps, neg_half_eps, L);
                    const T EVf = kinetic<T, D>(vt);
                    const T EXf = energy_after<Dist, T, D>(dist, xt, aux);
                    Hflf = EXf + EVf;
                    n_exec += 1;
                }
            }
            // ---- L state (hmc_state.py:93-100)
#pragma unroll
            for (int k = 0; k < D; ++k) { xt[k] = x[k]; vt[k] = v[k]; gt[k] = g[k]; }
            const T auxl = leapfrog_L<Dist, T, D>(dist, xt, vt, gt, eps, neg_half_eps, L);
            const T EVl = kinetic<T, D>(vt);
            const T EXl = energy_after<Dist, T, D>(dist, xt, auxl);
            const T Hl = EXl + EVl;
            n_E += 1;
            n_exec += 1;

            if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
                const Decision dc = decide_mj(p, i, attempt, (double)(H - Hl), (double)(H - Hflf), need_dwell);
                if (dc.fail) { report_failure(p, it); failed = true; }
                else {
                    choice = dc.choice; dwell = dc.dwell;
                    if (choice == 0) {
                        Hc = H; cflags = kCacheRef | kCacheValid;   // :399
#pragma unroll
                        for (int k = 0; k < D; ++k) { x[k] = xt[k]; v[k] = vt[k]; g[k] = gt[k]; }
                        EX = EXl; EV = EVl;
                        n_l += 1;
                    } else if (choice == 1) {
#pragma unroll
                        for (int k = 0; k < D; ++k) v[k] = -v[k];
                        Hc = Hl; cflags = kCacheValid;              // :410 clears cache_active; FLF(F z) = F L z
                        n_f += 1;
                    } else {
                        need_r = true;                              // refreshed below, by the whole warp
                        cflags = 0;                                 // :409
                        n_r += 1;
                    }
                }
            } else if (sampler == MJHMC_SAMPLER_CONTINUOUS_TIME) {
                // proposal is F L z (markov_jump_hmc.py:258)
                const Decision dc = decide_ct(p, i, attempt, (double)(H - Hl), need_dwell);
                if (dc.fail) { report_failure(p, it); failed = true; }
                else {
                    choice = dc.choice; dwell = dc.dwell;
                    if (choice == 1) {
#pragma unroll
                        for (int k = 0; k < D; ++k) { x[k] = xt[k]; v[k] = -vt[k]; g[k] = gt[k]; }
                        EX = EXl; EV = EVl;
                        n_fl += 1;
                    } else if (choice == 0) {
#pragma unroll
                        for (int k = 0; k < D; ++k) v[k] = -v[k];
                        n_f += 1;
                    } else {
                        need_r = true;
                        n_r += 1;
                    }
                }
            } else {
                const Decision dc = decide_discrete(p, i, attempt, (double)(H - Hl), s_coin[it & 1] != 0);
                choice = dc.choice;
                const bool acc = choice & 1u, flip = choice & 2u;
                if (acc) {
#pragma unroll
                    for (int k = 0; k < D; ++k) { x[k] = xt[k]; v[k] = -vt[k]; g[k] = gt[k]; }
                    EX = EXl; EV = EVl;
                }
                if (flip) {
#pragma unroll
                    for (int k = 0; k < D; ++k) v[k] = -v[k];
                }
                if (choice & 4u) {                                  // one coin for the whole batch :138
                    need_r = true;
                    n_r += 1;
                }
                n_l += (acc && flip);
                n_f += (flip && !acc);
                n_fl += (acc && !flip);
            }

        }
        // ---- R moves: V = V sqrt(1 - beta) + randn sqrt(beta) (hmc_state.py:126), shared by the warp
        refresh_momentum<T, D>(p, i, attempt, d, need_r, v, s_z);
        if (need_r) EV = kinetic<T, D>(v);
        if (active) {
            if (!failed) {
                // ---- record (markov_jump_hmc.py:169,334)
                if (p.samples) {
                    T* S = (T*)p.samples + (long long)it * p.s_stride_it + i;
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k < d) S[(long long)k * p.s_stride_k] = x[k];
                }
                if (p.dwell) p.dwell[(long long)it * p.n + i] = dwell;
                if (p.choice) p.choice[(long long)it * p.n + i] = (uint8_t)choice;
                if (p.energy) p.energy[(long long)it * p.n + i] = (double)(EX + EV);      // state.H() after the iteration
            }
        }
    }

    if (live) {
        T* Xout = (T*)p.Xout;
        T* Vout = (T*)p.Vout;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (k < d) {
                Xout[(long long)k * p.ld + i] = x[k];
                Vout[(long long)k * p.ld + i] = v[k];
            }
        }
        if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
            p.ca_out[i] = (uint8_t)cflags;
            ((T*)p.Hc_out)[i] = Hc;
        }
        if (p.dwell_last && sampler != MJHMC_SAMPLER_DISCRETE) p.dwell_last[i] = dwell;
    }

    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, n_E, n_exec};
    flush_counters(p.counters, loc, (unsigned long long)L);
}

// ---------------------------------------------------------------------------------------------------------------
// Same sampler with the particle state kept in SHARED memory between trajectories (ndims >= 6).
//
// In fused_sample_kernel a thread holds the state (x, v, g) AND the proposal (xt, vt, gt) in registers: 6 D values,
// 120 registers at D = 10 in fp64, 168 in all -- three CTAs (12 warps) per SM, and the Funnel of BASELINE config 5 sat
// on fixed-latency dependency stalls with 45 % of the issue slots used (profiles/r2_fused_funnel10d_cthmc_v2_*.txt).
// Here the registers hold one trajectory only; the state lives in this thread's own column of a [3][D][128] shared
// array (conflict-free: lane = column), is loaded when a trajectory starts (30 LDS against ~900 instructions of
// trajectory) and written back when the proposal is taken.  Same arithmetic in the same order: results are
// bit-identical to the register kernel (tests/test_gpu_screen.py compares the two).
template <typename T, int D>
__host__ __device__ constexpr bool fused_stash() {
#ifdef MJ_NO_STASH
    return false;
#else
    return sizeof(T) == 8 && D >= 10;
#endif
}
// Measured on the Funnel 10-d point (4M particles, ContinuousTimeHMC, profiles/r2_variants_funnel_stash.txt): the
// register allocation, not the warp count, is what the shared-memory state buys.  With the register budget of
// fused_sample_kernel (168, three CTAs) the trajectory no longer shares registers with a second copy of the state
// and runs 7 % faster; asking for four, five or six CTAs (128 / 96 / 80 registers) spills the trajectory itself and
// gives +2 %, +0 %, -25 %.  (One warp with four independent DFMA chains already saturates the fp64 pipe of its SM
// sub-partition -- tools/probe/fp64_probe.cu: latency 8.3 cycles, one warp-DFMA per 2 cycles -- so occupancy is not
// what this kernel lacks.)
template <typename T, int D>
__host__ __device__ constexpr int stash_min_blocks() {
#ifdef MJ_STASH_MINBLOCKS
    return MJ_STASH_MINBLOCKS;
#else
    return fused_min_blocks<T, D>();
#endif
}
template <typename T, int D>
constexpr size_t stash_smem_bytes() { return (size_t)4 * D * kFusedThreads * sizeof(T); }     // x, v, g, z

template <class Dist, typename T, int D>
__global__ void __launch_bounds__(kFusedThreads, stash_min_blocks<T, D>())
fused_stash_kernel(const __grid_constant__ LaunchParams p) {
    extern __shared__ __align__(16) unsigned char stash_raw[];
    typedef T (*Rows)[kFusedThreads];
    const Rows s_x = (Rows)stash_raw, s_v = s_x + D, s_g = s_v + D, s_z = s_g + D;
    __shared__ int s_coin[2];

    const int t = threadIdx.x;
    const long long i = (long long)blockIdx.x * blockDim.x + t;
    const bool live = i < p.n;
    const Dist dist(p);
    const int d = p.d;
    const T eps = (T)p.eps;
    const T neg_half_eps = (T)(-p.eps / 2.0);
    const int L = p.L;
    const int sampler = p.sampler;

    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0, n_E = 0, n_exec = 0;

    T xt[D], vt[D], gt[D];                 // the trajectory being integrated; between trajectories: scratch
    {
        const T* Xin = (const T*)p.Xin;
        const T* Vin = (const T*)p.Vin;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            xt[k] = (live && k < d) ? Xin[(long long)k * p.ld + i] : (T)0;
            vt[k] = (live && k < d) ? Vin[(long long)k * p.ld + i] : (T)0;
        }
    }
    const T aux0 = grad_aux<Dist, T, D>(dist, xt, gt);
    T EX = energy_after<Dist, T, D>(dist, xt, aux0);
    T EV = kinetic<T, D>(vt);
#pragma unroll
    for (int k = 0; k < D; ++k) { s_x[k][t] = xt[k]; s_v[k][t] = vt[k]; s_g[k][t] = gt[k]; }

    unsigned int cflags = 0;
    T Hc = (T)0;
    if (live && sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
        cflags = p.ca_in[i];
        Hc = ((const T*)p.Hc_in)[i];
    }
    double dwell = 0.0;
    bool failed = false;

    for (int it = 0; it < p.n_iter; ++it) {
        const unsigned long long attempt = p.attempt0 + (unsigned long long)it;
        const bool active = live && !failed;
        // what the decision does to the state: take the proposal, flip the sign of the momentum it ends with
        // (F L z = (x', -v'), F z = (x, -v)), refresh the momentum
        bool moved = false, negate = false, need_r = false;
        unsigned int choice = 0;
        const bool need_dwell = p.dwell != nullptr || (it + 1 == p.n_iter && p.dwell_last != nullptr);
        const T H = EX + EV;                                   // hmc_state.py:80-84
        T EXl = (T)0, EVl = (T)0;
        T Hflf = Hc;
        if (sampler == MJHMC_SAMPLER_DISCRETE) {
            if (t == 0) s_coin[it & 1] = draw_coin(p, attempt) < p.p_r;   // :138
            __syncthreads();
        }

        if (active) {
            // ---- FLF state (hmc_state.py:109-119): only its energy is ever read
            if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
                if (!(cflags & kCacheRef)) n_E += 1;           // the reference evaluates it here
                if (!(cflags & kCacheValid)) {
#pragma unroll
                    for (int k = 0; k < D; ++k) { xt[k] = s_x[k][t]; vt[k] = -s_v[k][t]; gt[k] = s_g[k][t]; }
                    const T aux = leapfrog_L<Dist, T, D>(dist, xt, vt, gt, eps, neg_half_eps, L);
                    const T EVf = kinetic<T, D>(vt);
                    const T EXf = energy_after<Dist, T, D>(dist, xt, aux);
                    Hflf = EXf + EVf;
                    n_exec += 1;
                }
            }
            // ---- L state (hmc_state.py:93-100)
#pragma unroll
            for (int k = 0; k < D; ++k) { xt[k] = s_x[k][t]; vt[k] = s_v[k][t]; gt[k] = s_g[k][t]; }
            const T auxl = leapfrog_L<Dist, T, D>(dist, xt, vt, gt, eps, neg_half_eps, L);
            EVl = kinetic<T, D>(vt);
            EXl = energy_after<Dist, T, D>(dist, xt, auxl);
            const T Hl = EXl + EVl;
            n_E += 1;
            n_exec += 1;

            if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
                const Decision dc = decide_mj(p, i, attempt, (double)(H - Hl), (double)(H - Hflf), need_dwell);
                if (dc.fail) { report_failure(p, it); failed = true; }
                else {
                    choice = dc.choice; dwell = dc.dwell;
                    if (choice == 0) { moved = true; Hc = H; cflags = kCacheRef | kCacheValid; n_l += 1; }   // :399
                    else if (choice == 1) { negate = true; Hc = Hl; cflags = kCacheValid; n_f += 1; }        // :410
                    else { need_r = true; cflags = 0; n_r += 1; }                                            // :409
                }
            } else if (sampler == MJHMC_SAMPLER_CONTINUOUS_TIME) {
                const Decision dc = decide_ct(p, i, attempt, (double)(H - Hl), need_dwell);                  // F L z :258
                if (dc.fail) { report_failure(p, it); failed = true; }
                else {
                    choice = dc.choice; dwell = dc.dwell;
                    if (choice == 1) { moved = true; negate = true; n_fl += 1; }
                    else if (choice == 0) { negate = true; n_f += 1; }
                    else { need_r = true; n_r += 1; }
                }
            } else {
                const Decision dc = decide_discrete(p, i, attempt, (double)(H - Hl), s_coin[it & 1] != 0);
                choice = dc.choice;
                const bool acc = choice & 1u, flip = choice & 2u;
                moved = acc; negate = acc != flip;                  // F L z, then the flip
                if (choice & 4u) { need_r = true; n_r += 1; }       // one coin for the whole batch :138
                n_l += (acc && flip);
                n_f += (flip && !acc);
                n_fl += (acc && !flip);
            }

            // ---- the new state: (xt, vt) become the current position and momentum
            if (moved) {
#pragma unroll
                for (int k = 0; k < D; ++k) { s_x[k][t] = xt[k]; s_g[k][t] = gt[k]; }
                EX = EXl; EV = EVl;
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) { xt[k] = s_x[k][t]; vt[k] = s_v[k][t]; }
            }
            if (negate) {
#pragma unroll
                for (int k = 0; k < D; ++k) vt[k] = -vt[k];
            }
        }
        // ---- R moves: V = V sqrt(1 - beta) + randn sqrt(beta) (hmc_state.py:126), shared by the warp
        refresh_momentum<T, D>(p, i, attempt, d, need_r, vt, s_z);
        if (need_r) EV = kinetic<T, D>(vt);
        if (moved || negate || need_r) {
#pragma unroll
            for (int k = 0; k < D; ++k) s_v[k][t] = vt[k];
        }
        if (active && !failed) {
            // ---- record (markov_jump_hmc.py:169,334)
            if (p.samples) {
                T* S = (T*)p.samples + (long long)it * p.s_stride_it + i;
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) S[(long long)k * p.s_stride_k] = xt[k];
            }
            if (p.dwell) p.dwell[(long long)it * p.n + i] = dwell;
            if (p.choice) p.choice[(long long)it * p.n + i] = (uint8_t)choice;
            if (p.energy) p.energy[(long long)it * p.n + i] = (double)(EX + EV);      // state.H() after the iteration
        }
    }

    if (live) {
        T* Xout = (T*)p.Xout;
        T* Vout = (T*)p.Vout;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (k < d) {
                Xout[(long long)k * p.ld + i] = s_x[k][t];
                Vout[(long long)k * p.ld + i] = s_v[k][t];
            }
        }
        if (sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
            p.ca_out[i] = (uint8_t)cflags;
            ((T*)p.Hc_out)[i] = Hc;
        }
        if (p.dwell_last && sampler != MJHMC_SAMPLER_DISCRETE) p.dwell_last[i] = dwell;
    }

    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, n_E, n_exec};
    flush_counters(p.counters, loc, (unsigned long long)L);
}

// Host-side launcher for one (Dist, T, D) instantiation.
template <class Dist, typename T, int D>
cudaError_t launch_fused(const LaunchParams& p, cudaStream_t stream) {
    const long long blocks = (p.n + kFusedThreads - 1) / kFusedThreads;
    if (blocks == 0) return cudaSuccess;
    if constexpr (fused_stash<T, D>()) {
        if (!(p.rng_flags & MJHMC_RNG_FLAG_REGISTER_STATE)) {
            static bool configured = false;
            if (!configured) {
                const cudaError_t e = cudaFuncSetAttribute(fused_stash_kernel<Dist, T, D>,
                                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           (int)stash_smem_bytes<T, D>());
                if (e != cudaSuccess) return e;
                configured = true;
            }
            fused_stash_kernel<Dist, T, D><<<(unsigned)blocks, kFusedThreads, stash_smem_bytes<T, D>(), stream>>>(p);
            return cudaGetLastError();
        }
    }
    fused_sample_kernel<Dist, T, D><<<(unsigned)blocks, kFusedThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// Dispatch table entry point implemented per translation unit (fused_inst_*.cu).
typedef cudaError_t (*fused_launch_fn)(const LaunchParams&, cudaStream_t);
fused_launch_fn find_fused_f64(int dist_kind, int D);
fused_launch_fn find_fused_f32(int dist_kind, int D);

template <typename T, int D>
fused_launch_fn pick_dist(int kind) {
    switch (kind) {
        case MJHMC_DIST_TEST_GAUSSIAN:  return &launch_fused<TestGaussianD<T, D>, T, D>;
        case MJHMC_DIST_DIAG_GAUSSIAN:  return &launch_fused<DiagGaussianD<T, D>, T, D>;
        case MJHMC_DIST_ROUGH_WELL:     return &launch_fused<RoughWellD<T, D>, T, D>;
        case MJHMC_DIST_FUNNEL:         return &launch_fused<FunnelD<T, D, false>, T, D>;
        case MJHMC_DIST_FUNNEL_LITERAL: return &launch_fused<FunnelD<T, D, true>, T, D>;
        case MJHMC_DIST_MULTIMODAL:     return &launch_fused<MultimodalD<T, D>, T, D>;
        default: return nullptr;
    }
}

// Register-kernel template dimension for a runtime ndims (0 = none).
inline int fused_template_dim(int d) {
    static const int dims[] = {1, 2, 3, 4, 6, 8, 10, 16};
    for (int t : dims) if (d <= t) return t;
    return 0;
}

}  // namespace mjhmc
