// Shared device helpers: Philox4x32-10 stream, draw access, reductions.
// sm_100a only.  Reference call sites replaced: np.random.* in
// samplers/hmc_state.py:126, samplers/markov_jump_hmc.py:125,132,138, misc/utils.py:42.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/mjhmc_b200.h"

// Once-per-iteration transcendental code (exp/log/sincospi in fp64) is kept out of line so its
// register needs do not set the occupancy of the leapfrog loop.
#ifndef MJ_INLINE_COLD
#define MJ_COLD static __device__ __noinline__
#else
#define MJ_COLD static __device__ __forceinline__
#endif

namespace mjhmc {

constexpr int kMaxRegDims = 16;          // largest ndims with a register-resident fused kernel

// Kernel-side view of one launch (passed by value as a __grid_constant__ parameter).
struct LaunchParams {
    // state
    const void* Xin; const void* Vin; void* Xout; void* Vout;
    const void* Hc_in; void* Hc_out; const uint8_t* ca_in; uint8_t* ca_out;
    long long n, ld;
    // outputs
    void* samples; long long s_stride_k, s_stride_it;
    double* dwell; double* dwell_last; uint8_t* choice; double* energy;
    unsigned long long* counters;
    // hyper-parameters
    int sampler, L, n_iter, d;
    double eps, p_flip, p_r, r_keep, r_mix;   // r_keep = sqrt(1-beta), r_mix = sqrt(beta)
    // rng
    int rng_mode, rng_flags;
    unsigned long long seed, attempt0, particle0;
    const double* Z; const double* U; const double* U0; long long inj_ld;
    // distribution
    double dp[4];
    double coef[12];                 // distribution-specific constants precomputed on the host (api.cu)
    const void* a0; const void* a1; const void* a2; int nbasis;
    const void* ws;                  // fp32 dense kinds: the pre-tiled matrix planes (mjhmc_dense_tc_prepare)
};

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    // numpy legacy double: (a>>5, b>>6) -> [0,1)
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ uint4 philox_at(unsigned long long seed, unsigned long long particle,
                                           unsigned long long attempt, uint32_t slot) {
    return philox4x32_10(make_uint4((uint32_t)particle, (uint32_t)(particle >> 32), (uint32_t)attempt, slot),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// The three per-particle uniforms of one attempt (slots: see oracle/philox.py).
struct Uniform3 { double u0, u1, u2; };

__device__ __forceinline__ Uniform3 draw_uniforms(const LaunchParams& p, long long i, unsigned long long attempt,
                                                  bool need_u2) {
    Uniform3 r;
    if (p.rng_mode == MJHMC_RNG_INJECT) {
        const double* base = p.U + (attempt * 3ull) * (unsigned long long)p.inj_ld + p.particle0 + i;
        r.u0 = base[0];
        r.u1 = base[p.inj_ld];
        r.u2 = need_u2 ? base[2 * p.inj_ld] : 0.0;
    } else {
        const unsigned long long g = p.particle0 + (unsigned long long)i;
        const uint4 w = philox_at(p.seed, g, attempt, 0u);
        r.u0 = u53(w.x, w.y);
        r.u1 = u53(w.z, w.w);
        if (need_u2) {
            const uint4 w1 = philox_at(p.seed, g, attempt, 1u);
            r.u2 = u53(w1.x, w1.y);
        } else {
            r.u2 = 0.0;
        }
    }
    return r;
}

__device__ __forceinline__ double draw_coin(const LaunchParams& p, unsigned long long attempt) {
    if (p.rng_mode == MJHMC_RNG_INJECT) return p.U0[attempt];
    const uint4 w = philox_at(p.seed, 0xFFFFFFFFFFFFFFFFull, attempt, 0u);
    return u53(w.x, w.y);
}

// Box-Muller pair j of particle i, attempt a: normals 2j and 2j+1 (z1 unused when 2j+1 == d).
MJ_COLD void normal_pair(const LaunchParams& p, long long i, unsigned long long attempt,
                                            int j, int d, double& z0, double& z1) {
    if (p.rng_mode == MJHMC_RNG_INJECT) {
        const double* base = p.Z + (attempt * (unsigned long long)d + 2ull * j) * (unsigned long long)p.inj_ld
                             + p.particle0 + i;
        z0 = base[0];
        z1 = (2 * j + 1 < d) ? base[p.inj_ld] : 0.0;
    } else {
        const uint4 w = philox_at(p.seed, p.particle0 + (unsigned long long)i, attempt, 2u + (uint32_t)j);
        const double u1 = u53(w.x, w.y), u2 = u53(w.z, w.w);
        const double r = sqrt(-2.0 * log(1.0 - u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        z0 = r * c;
        z1 = r * s;
    }
}

// Standard normals z[0..d) for particle i, attempt a (rows >= d are zero padding).
template <typename T, int D>
__device__ __forceinline__ void draw_normals(const LaunchParams& p, long long i, unsigned long long attempt,
                                             int d, T (&z)[D]) {
#pragma unroll
    for (int j = 0; j < (D + 1) / 2; ++j) {
        double z0 = 0.0, z1 = 0.0;
        if (2 * j < d) normal_pair(p, i, attempt, j, d, z0, z1);
        z[2 * j] = (T)z0;
        if (2 * j + 1 < D) z[2 * j + 1] = (2 * j + 1 < d) ? (T)z1 : (T)0;
    }
}

// np.random.exponential(scale = 1/rate) == (1/rate) * -log(1 - u); zero rate -> inf (utils.py:38-42)
__device__ __forceinline__ double exp_draw(double rate, double u) {
    return rate == 0.0 ? INFINITY : (1.0 / rate) * (-log(1.0 - u));
}

// transition rate exp(H - H')**.5 (markov_jump_hmc.py:341-347)
__device__ __forceinline__ double jump_rate(double ediff) { return sqrt(exp(ediff)); }

// ---------------------------------------------------------------- transitions
// The operator choice of one particle for one attempt.  `fail` = a non-finite rate
// (misc/utils.py:41-48), which the host turns into the batch-wide back-off / ValueError.
struct Decision { unsigned int choice; double dwell; bool fail; };

// MarkovJumpHMC (markov_jump_hmc.py:366-396): choice 0 = L, 1 = F, 2 = R.
// ediff_l = H - H_L, ediff_flf = H - H_FLF.  The *_u forms take the uniforms of the attempt as arguments
// (the streaming kernel draws them before the trajectory so the Philox rounds overlap the fp64 work).
__device__ __forceinline__ Decision decide_mj_u(double p_r, double u0, double u1, double u2,
                                                double ediff_l, double ediff_flf) {
    Decision dc; dc.choice = 0; dc.dwell = 0.0; dc.fail = false;
    const double rl = jump_rate(ediff_l);
    const double rflf = jump_rate(ediff_flf);
    if (!(isfinite(rl) && isfinite(rflf))) { dc.fail = true; return dc; }
    const double rf = rflf - (rl < rflf ? rl : rflf);              // :368
    const double tl = exp_draw(rl, u0);
    const double tf = exp_draw(rf, u1);
    const double tr = exp_draw(p_r, u2);
    dc.dwell = tl;                                                 // min_idx([l, f, r]): first minimum wins
    if (tf < dc.dwell) { dc.choice = 1; dc.dwell = tf; }
    if (tr < dc.dwell) { dc.choice = 2; dc.dwell = tr; }
    return dc;
}
// ---- certified single-precision screening of the holding-time race
// The operator choice is the index of the first minimum of three exponential draws (1/rate) * -log(1 - u)
// (utils.py:15-49).  Evaluated literally that is three fp64 logs, three fp64 divisions and an fp64 exp + sqrt per
// rate: ~400 instructions, 40 % of the Funnel / ContinuousTimeHMC kernel of BASELINE config 5
// (profiles/r2_fused_funnel10d_cthmc_v1_shared_refresh.txt).  Only the ORDER of the three times decides the move, and only the
// winner's time is ever stored.  So each time is first enclosed in a single-precision interval [lo, hi] with a
// rigorous error budget (MUFU lg2 / ex2 / rcp, ~60 instructions); where the intervals separate, the choice is the
// one the fp64 code would make, and the winner's time -- if the caller stores it -- is evaluated with exactly the
// reference expression.  Where they overlap (a few lanes in 1e5), the literal fp64 code decides (decide_*_u).
// Error budgets (all bounds hold for the fp64 VALUES the literal code computes, which differ from the exact real
// numbers by < 1e-15 relative, far inside the slack):
//   w = -log(1 - u):  x = float(1 - u) carries 2^-24 relative error (|dw| <= 6e-8); __logf: absolute error
//       2^-21.41 for x in [0.5, 2], 3 ulp otherwise  =>  |w~ - w| <= 6e-7 (1 + w~).
//   r = sqrt(exp(e)), |e| <= 64:  a = float(e log2(e) / 2), |da| <= 46.2 * 2^-24 => 1.9e-6 relative in 2^a;
//       ex2.approx 2^-22  =>  |r~ / r - 1| <= 2.2e-6, enclosed with 4e-6.  e < -64: r in [0, 1.3e-14].
//       e > 64 or NaN: not screened (the literal code also detects the non-finite rate, utils.py:41-48).
//   t = w / r:  __fdividef 2 ulp, every product with (1 +- 1e-6) absorbs the fp32 roundings of the step.
struct Encl { float lo, hi; };

__device__ __forceinline__ Encl encl_neg_log1m(double u) {
    const float x = __double2float_rn(1.0 - u);
    const float w = -__logf(x);
    const float e = fmaf(w, 6e-7f, 6e-7f);
    Encl r; r.lo = fmaxf(w - e, 0.0f); r.hi = w + e;
    return r;
}
__device__ __forceinline__ bool encl_jump_rate(double ediff, Encl& r) {
    if (!(ediff <= 64.0)) return false;
    if (ediff < -64.0) { r.lo = 0.0f; r.hi = 1.3e-14f; return true; }      // exp(-32) = 1.27e-14
    const float a = __double2float_rn(ediff * 0.72134752044448170368);      // log2(e) / 2
    float v;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(a));
    r.lo = v * (1.0f - 4e-6f); r.hi = v * (1.0f + 4e-6f);
    return true;
}
__device__ __forceinline__ Encl encl_time(Encl w, Encl r) {
    Encl t;
    t.lo = r.hi > 0.0f ? __fdividef(w.lo, r.hi) * (1.0f - 1e-6f) : INFINITY;
    t.hi = r.lo > 0.0f ? __fdividef(w.hi, r.lo) * (1.0f + 1e-6f) : INFINITY;
    return t;
}
// p_r is a launch constant; screened only inside a range where float(p_r) and __fdividef are harmless
__device__ __forceinline__ bool encl_refresh_time(double p_r, double u2, Encl& t) {
    if (p_r == 0.0) { t.lo = t.hi = INFINITY; return true; }                // exp_draw: zero rate -> inf
    if (!(p_r >= 1e-30 && p_r <= 1e30)) return false;
    const float pf = __double2float_rn(p_r);
    Encl r; r.lo = pf * (1.0f - 1e-6f); r.hi = pf * (1.0f + 1e-6f);
    t = encl_time(encl_neg_log1m(u2), r);
    return true;
}
// index of the first minimum of (a, b, c) the way min_idx finds it (a strict `<` replaces the incumbent), or -1 when
// the enclosures do not decide
__device__ __forceinline__ int certain_first_min(Encl a, Encl b, Encl c) {
    if (b.lo > a.hi && c.lo > a.hi) return 0;
    if (b.hi < a.lo && c.lo > b.hi) return 1;
    if (c.hi < a.lo && c.hi < b.lo) return 2;
    return -1;
}

// The part of the screen that depends on the uniforms only (the streaming kernel evaluates it before the trajectory,
// where it fills latency slots; after the energies meet only the rates, three divisions and the comparisons remain).
struct RaceDraws { Encl w0, w1, tr; bool ok; };
__device__ __forceinline__ RaceDraws race_draws(double p_r, double u0, double u1, double u2, bool literal) {
    RaceDraws rd;
    rd.w0 = encl_neg_log1m(u0);
    rd.w1 = encl_neg_log1m(u1);
    rd.ok = !literal && encl_refresh_time(p_r, u2, rd.tr);
    return rd;
}

// need_dwell: the caller stores the holding time of this attempt (dwelling-time record or the last iteration of a
// launch); otherwise only the choice is produced.
__device__ __forceinline__ Decision decide_mj_screened(double p_r, double u0, double u1, double u2, const RaceDraws& rd,
                                                       double ediff_l, double ediff_flf, bool need_dwell) {
#ifndef MJ_NO_SCREEN
    Encl rl, rflf;
    if (rd.ok && encl_jump_rate(ediff_l, rl) && encl_jump_rate(ediff_flf, rflf)) {
        Encl rf;                                                           // rflf - min(rl, rflf) = max(rflf - rl, 0)
        rf.lo = fmaxf(rflf.lo - rl.hi, 0.0f) * (1.0f - 1e-6f);
        rf.hi = fmaxf(rflf.hi - rl.lo, 0.0f) * (1.0f + 1e-6f);
        const Encl tl = encl_time(rd.w0, rl);
        const Encl tf = encl_time(rd.w1, rf);
        const int c = certain_first_min(tl, tf, rd.tr);
        if (c >= 0) {
            Decision dc; dc.choice = (unsigned int)c; dc.dwell = 0.0; dc.fail = false;
            if (need_dwell) {
                double rate = p_r, u = u2;
                if (c != 2) {
                    const double erl = jump_rate(ediff_l);
                    rate = erl; u = u0;
                    if (c == 1) { const double erflf = jump_rate(ediff_flf); rate = erflf - (erl < erflf ? erl : erflf); u = u1; }
                }
                dc.dwell = exp_draw(rate, u);
            }
            return dc;
        }
    }
#endif
    return decide_mj_u(p_r, u0, u1, u2, ediff_l, ediff_flf);
}
MJ_COLD Decision decide_mj(const LaunchParams& p, long long i, unsigned long long attempt,
                           double ediff_l, double ediff_flf, bool need_dwell) {
    const Uniform3 u = draw_uniforms(p, i, attempt, p.p_r != 0.0);
    const RaceDraws rd = race_draws(p.p_r, u.u0, u.u1, u.u2, (p.rng_flags & MJHMC_RNG_FLAG_LITERAL_RACE) != 0);
    return decide_mj_screened(p.p_r, u.u0, u.u1, u.u2, rd, ediff_l, ediff_flf, need_dwell);
}

// ContinuousTimeHMC (markov_jump_hmc.py:261-275): choice 0 = F, 1 = FL, 2 = R.
__device__ __forceinline__ Decision decide_ct_u(double p_r, double u0, double u1, double u2, double ediff_fl) {
    Decision dc; dc.choice = 0; dc.dwell = 0.0; dc.fail = false;
    const double rfl = jump_rate(ediff_fl);
    if (!isfinite(rfl)) { dc.fail = true; return dc; }
    const double tfl = exp_draw(rfl, u0);
    const double tf = exp_draw(1.0, u1);
    const double tr = exp_draw(p_r, u2);
    dc.dwell = tf;                                                 // min_idx([f, fl, r]) :271
    if (tfl < dc.dwell) { dc.choice = 1; dc.dwell = tfl; }
    if (tr < dc.dwell) { dc.choice = 2; dc.dwell = tr; }
    return dc;
}
__device__ __forceinline__ Decision decide_ct_screened(double p_r, double u0, double u1, double u2, const RaceDraws& rd,
                                                       double ediff_fl, bool need_dwell) {
#ifndef MJ_NO_SCREEN
    Encl rfl;
    if (rd.ok && encl_jump_rate(ediff_fl, rfl)) {
        const Encl tfl = encl_time(rd.w0, rfl);
        const int c = certain_first_min(rd.w1, tfl, rd.tr);               // rate 1: t_f = (1.0 / 1.0) * w1
        if (c >= 0) {
            Decision dc; dc.choice = (unsigned int)c; dc.dwell = 0.0; dc.fail = false;
            if (need_dwell) {
                double rate = 1.0, u = u1;
                if (c == 1) { rate = jump_rate(ediff_fl); u = u0; }
                if (c == 2) { rate = p_r; u = u2; }
                dc.dwell = exp_draw(rate, u);
            }
            return dc;
        }
    }
#endif
    return decide_ct_u(p_r, u0, u1, u2, ediff_fl);
}
MJ_COLD Decision decide_ct(const LaunchParams& p, long long i, unsigned long long attempt,
                           double ediff_fl, bool need_dwell) {
    const Uniform3 u = draw_uniforms(p, i, attempt, p.p_r != 0.0);
    const RaceDraws rd = race_draws(p.p_r, u.u0, u.u1, u.u2, (p.rng_flags & MJHMC_RNG_FLAG_LITERAL_RACE) != 0);
    return decide_ct_screened(p.p_r, u.u0, u.u1, u.u2, rd, ediff_fl, need_dwell);
}

// HMCBase / HMC / ControlHMC (markov_jump_hmc.py:106-148): bit0 = FL accepted, bit1 = flipped,
// bit2 = the batch-wide R coin fired.
__device__ __forceinline__ unsigned int decide_discrete_u(double p_flip, double u0, double u1, double ediff,
                                                          bool coin_fired) {
    // u0 < leap_prob = min(1, exp(ediff)) (markov_jump_hmc.py:106-114,125).  exp is evaluated only where the
    // elementary bounds 1 + e <= exp(e) <= 1 / (1 - e)  (e < 0) leave the comparison open; the relative guard
    // of 2^-40 on both bounds is far wider than the rounding of the three operations, so the outcome is the
    // one of the exact comparison.  A warp of small |e| skips the ~90-instruction fp64 exp altogether.
    bool acc = true;                                               // e >= 0: leap_prob = 1 > u0
    if (!(ediff >= 0.0)) {
        const double guard = 9.094947017729282e-13;                // 2^-40
        const double lo = 1.0 + ediff;
        if (!(u0 < lo - guard * fabs(lo))) {
            const double hi = __drcp_rn(1.0 - ediff);
            acc = (u0 > hi + guard * hi) ? false : (u0 < exp(ediff));
        }
    }
    return (acc ? 1u : 0u) | (u1 < p_flip ? 2u : 0u) | (coin_fired ? 4u : 0u);
}
MJ_COLD Decision decide_discrete(const LaunchParams& p, long long i, unsigned long long attempt,
                                 double ediff, bool coin_fired) {
    Decision dc; dc.dwell = 0.0; dc.fail = false;
    const Uniform3 u = draw_uniforms(p, i, attempt, false);
    dc.choice = decide_discrete_u(p.p_flip, u.u0, u.u1, ediff, coin_fired);
    return dc;
}

// Streaming-kernel forms: the uniforms are drawn before the trajectory (the Philox rounds overlap the fp64 work),
// the decision runs out of line once the energies are known.
static __device__ __noinline__ Decision decide_mj_s(double p_r, double u0, double u1, double u2, RaceDraws rd,
                                                    double ediff_l, double ediff_flf, bool need_dwell) {
    return decide_mj_screened(p_r, u0, u1, u2, rd, ediff_l, ediff_flf, need_dwell);
}
static __device__ __noinline__ Decision decide_ct_s(double p_r, double u0, double u1, double u2, RaceDraws rd,
                                                    double ediff_fl, bool need_dwell) {
    return decide_ct_screened(p_r, u0, u1, u2, rd, ediff_fl, need_dwell);
}
static __device__ __noinline__ unsigned int decide_discrete_s(double p_flip, double u0, double u1, double ediff,
                                                              bool coin_fired) {
    return decide_discrete_u(p_flip, u0, u1, ediff, coin_fired);
}

__device__ __forceinline__ void report_failure(const LaunchParams& p, int it) {
    atomicMin(p.counters + (size_t)(blockIdx.x % MJHMC_COUNTER_STRIPES) * MJHMC_N_COUNTERS + MJHMC_CNT_FAIL,
              (unsigned long long)it);
}

// ---------------------------------------------------------------- counters
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-reduce the six 32-bit local counters {l, f, fl, r, E, exec} (REDUX per warp, shared
// atomics per CTA) and add them to this CTA's counter stripe.  dEdX = L * E and the executed
// gradient evaluations = L * exec are formed here (L is constant inside one launch).
__device__ __forceinline__ void flush_counters(unsigned long long* counters, const unsigned int (&loc)[6],
                                               unsigned long long L) {
    __shared__ unsigned int sm[6];
    if (threadIdx.x < 6) sm[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const unsigned int s = __reduce_add_sync(0xffffffffu, loc[c]);
        if (lane == 0 && s) atomicAdd(&sm[c], s);
    }
    __syncthreads();
    unsigned long long* row = counters + (size_t)(blockIdx.x % MJHMC_COUNTER_STRIPES) * MJHMC_N_COUNTERS;
    if (threadIdx.x < 6) {
        const unsigned long long s = sm[threadIdx.x];
        const int slot[6] = {MJHMC_CNT_L, MJHMC_CNT_F, MJHMC_CNT_FL, MJHMC_CNT_R, MJHMC_CNT_E, MJHMC_CNT_EXEC};
        if (s) atomicAdd(row + slot[threadIdx.x], threadIdx.x == 5 ? s * L : s);
        if (threadIdx.x == 4 && s) atomicAdd(row + MJHMC_CNT_DEDX, s * L);
    }
}

// Q coefficients of sin(pi f) = f Q(f^2) on f^2 in [0, 1/4] (dists.cuh: scaled_sin_halfturns)
static const double kSinQ64[9] = {3.14159265358979312e+00, -5.16771278004996937e+00, 2.55016403987730067e+00,
                                  -5.99264529318944694e-01, 8.21458865731028998e-02, -7.37043050591694362e-03,
                                  4.66299816189839390e-04, -2.19034970746260181e-05, 7.69782676822419091e-07};
static const double kSinQ32[5] = {3.14159264007720340e+00, -5.16771007666831483e+00, 2.55007738652891192e+00,
                                  -5.98290411283427526e-01, 7.76559122760138720e-02};

// Fills LaunchParams::coef for the RoughWell gradient: -(2 pi / scale2) * Q.
inline void fill_roughwell_coef(double* coef, int dtype, double scale2) {
    const double c = -2.0 * 3.14159265358979323846 / scale2;
    if (dtype == MJHMC_F64) for (int j = 0; j < 9; ++j) coef[j] = c * kSinQ64[j];
    else for (int j = 0; j < 5; ++j) coef[j] = c * kSinQ32[j];
}

}  // namespace mjhmc
