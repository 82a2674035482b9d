// C ABI of libmjhmc_b200.so (include/mjhmc_b200.h): argument checking, parameter
// packing and kernel dispatch.  No torch types, no hidden allocation, no CPU path.
#include <string>
#include <cstdio>
#include <cstring>
#include <climits>
#include "fused_elementwise.cuh"
#include "unfused.h"
#include "analysis.h"
#include "dense.h"
#include "stream.h"

namespace mjhmc {
fused_launch_fn find_fused_f64_g0(int, int); fused_launch_fn find_fused_f64_g1(int, int);
fused_launch_fn find_fused_f64_g2(int, int); fused_launch_fn find_fused_f64_g3(int, int);
fused_launch_fn find_fused_f32_g0(int, int); fused_launch_fn find_fused_f32_g1(int, int);
fused_launch_fn find_fused_f32_g2(int, int); fused_launch_fn find_fused_f32_g3(int, int);

static thread_local std::string g_err;
static int fail(const char* fmt, const char* a = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a);
    g_err = buf;
    return -1;
}
static int check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -2;
}

static fused_launch_fn find_fused(int dtype, int kind, int D) {
    fused_launch_fn f = nullptr;
    if (dtype == MJHMC_F64) {
        if (!f) f = find_fused_f64_g0(kind, D);
        if (!f) f = find_fused_f64_g1(kind, D);
        if (!f) f = find_fused_f64_g2(kind, D);
        if (!f) f = find_fused_f64_g3(kind, D);
    } else {
        if (!f) f = find_fused_f32_g0(kind, D);
        if (!f) f = find_fused_f32_g1(kind, D);
        if (!f) f = find_fused_f32_g2(kind, D);
        if (!f) f = find_fused_f32_g3(kind, D);
    }
    return f;
}

static bool is_elementwise(int kind) {
    return (kind >= MJHMC_DIST_TEST_GAUSSIAN && kind <= MJHMC_DIST_FUNNEL_LITERAL) || kind == MJHMC_DIST_MULTIMODAL;
}
static bool is_dense(int kind) { return kind == MJHMC_DIST_DENSE_GAUSSIAN || kind == MJHMC_DIST_PRODUCT_OF_T; }

static int fill_hp(LaunchParams& p, const mjhmc_hp* hp) {
    if (hp->sampler < MJHMC_SAMPLER_DISCRETE || hp->sampler > MJHMC_SAMPLER_MARKOV_JUMP) return fail("bad sampler kind");
    if (hp->num_leapfrog_steps < 0) return fail("num_leapfrog_steps < 0");
    if (!(hp->beta >= 0.0 && hp->beta <= 1.0)) return fail("beta must be in [0, 1]");
    if (hp->sampler != MJHMC_SAMPLER_DISCRETE && !(hp->p_r >= 0.0 && hp->p_r < INFINITY))
        return fail("p_r must be a finite non-negative rate (Infinite rate)");
    p.sampler = hp->sampler;
    p.L = hp->num_leapfrog_steps;
    p.eps = hp->epsilon;
    p.p_flip = hp->p_flip;
    p.p_r = hp->p_r;
    p.r_keep = sqrt(1.0 - hp->beta);
    p.r_mix = sqrt(hp->beta);
    return 0;
}

static int fill_rng(LaunchParams& p, const mjhmc_rng* rng) {
    p.rng_mode = rng->mode;
    p.rng_flags = rng->flags;
    p.seed = rng->seed;
    p.attempt0 = rng->attempt0;
    p.particle0 = rng->particle0;
    p.Z = rng->Z; p.U = rng->U; p.U0 = rng->U0; p.inj_ld = rng->inj_ld;
    if (rng->mode == MJHMC_RNG_INJECT) {
        if (!rng->U) return fail("INJECT mode needs U");
        if (rng->inj_ld <= 0) return fail("INJECT mode needs inj_ld");
    } else if (rng->mode != MJHMC_RNG_PHILOX) {
        return fail("bad rng mode");
    }
    return 0;
}

static void fill_outputs(LaunchParams& p, const mjhmc_outputs* o) {
    p.samples = o->samples; p.s_stride_k = o->stride_k; p.s_stride_it = o->stride_it;
    p.dwell = o->dwell; p.dwell_last = o->dwell_last; p.choice = o->choice; p.energy = o->energy;
    p.counters = (unsigned long long*)o->counters;
}

static int fill_dist(LaunchParams& p, const mjhmc_dist* dist) {
    p.d = dist->ndims;
    p.nbasis = dist->nbasis;
    for (int k = 0; k < 4; ++k) p.dp[k] = dist->p[k];
    p.a0 = dist->a0; p.a1 = dist->a1; p.a2 = dist->a2; p.ws = dist->ws;
    if (dist->kind == MJHMC_DIST_ROUGH_WELL) fill_roughwell_coef(p.coef, dist->dtype, dist->p[1]);
    return 0;
}

static DistParams dist_params(const mjhmc_dist* dist) {
    DistParams dp;
    dp.kind = dist->kind; dp.d = dist->ndims; dp.nbasis = dist->nbasis;
    for (int k = 0; k < 4; ++k) dp.p[k] = dist->p[k];
    dp.a0 = dist->a0; dp.a1 = dist->a1; dp.a2 = dist->a2;
    for (int k = 0; k < 12; ++k) dp.coef[k] = 0.0;
    if (dist->kind == MJHMC_DIST_ROUGH_WELL) fill_roughwell_coef(dp.coef, dist->dtype, dist->p[1]);
    return dp;
}

static int check_dist(const mjhmc_dist* dist) {
    if (!dist) return fail("dist is NULL");
    if (dist->dtype != MJHMC_F32 && dist->dtype != MJHMC_F64) return fail("bad dtype");
    if (dist->ndims <= 0) return fail("ndims must be positive");
    if (dist->kind < MJHMC_DIST_TEST_GAUSSIAN || dist->kind > MJHMC_DIST_MULTIMODAL) return fail("bad distribution kind");
    if ((dist->kind == MJHMC_DIST_DIAG_GAUSSIAN || is_dense(dist->kind)) && !dist->a0) return fail("distribution needs a0");
    if (dist->kind == MJHMC_DIST_PRODUCT_OF_T && (!dist->a1 || !dist->a2 || dist->nbasis <= 0)) return fail("ProductOfT needs a1, a2, nbasis");
    return 0;
}

}  // namespace mjhmc

using namespace mjhmc;

extern "C" {

const char* mjhmc_last_error(void) { return g_err.c_str(); }
int mjhmc_abi_version(void) { return MJHMC_ABI_VERSION; }

int mjhmc_fused_supported(const mjhmc_dist* dist) {
    if (check_dist(dist)) return 0;
    if (is_elementwise(dist->kind)) {
        const int D = fused_template_dim(dist->ndims);
        if (D && find_fused(dist->dtype, dist->kind, D)) return 1;
        return stream_supported(dist->dtype, dist->kind, dist->ndims) ? 1 : 0;
    }
    return dense_supported(dist->dtype, dist->kind, dist->ndims, dist->nbasis) ? 1 : 0;
}

int mjhmc_stream_supported(const mjhmc_dist* dist) {
    if (check_dist(dist)) return 0;
    return stream_supported(dist->dtype, dist->kind, dist->ndims) ? 1 : 0;
}

void mjhmc_stream_set_tma(int32_t enabled) { stream_set_tma(enabled); }
int mjhmc_stream_probe_blocks(int64_t smem_bytes) { return stream_probe_blocks(smem_bytes); }
void mjhmc_stream_last_launch(int64_t* out7_host) { if (out7_host) stream_last_launch((long long*)out7_host); }

static int sample_impl(const mjhmc_dist* dist, const mjhmc_hp* hp, const mjhmc_rng* rng,
                       const mjhmc_state* in, const mjhmc_state* out, int32_t n_iter,
                       const mjhmc_outputs* o, void* stream_, bool force_stream) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (check_dist(dist)) return -1;
    if (!hp || !rng || !in || !out || !o) return fail("NULL argument");
    if (!o->counters) return fail("outputs.counters is required");
    if (n_iter < 0) return fail("n_iter < 0");
    if (in->n != out->n || in->ld != out->ld) return fail("in/out state shapes differ");
    if (in->n < 0 || in->ld < in->n) return fail("bad n / ld");
    if (!in->X || !in->V || !out->X || !out->V) { if (in->n) return fail("state X/V is NULL"); }
    LaunchParams p;
    memset(&p, 0, sizeof p);
    if (fill_hp(p, hp)) return -1;
    if (fill_rng(p, rng)) return -1;
    if (rng->mode == MJHMC_RNG_INJECT) {
        const bool needs_z = hp->sampler == MJHMC_SAMPLER_DISCRETE ? hp->p_r > 0.0 : hp->p_r != 0.0;
        if (needs_z && !rng->Z) return fail("INJECT mode needs Z");
        if (hp->sampler == MJHMC_SAMPLER_DISCRETE && !rng->U0) return fail("INJECT mode needs U0 for discrete samplers");
    }
    if (hp->sampler == MJHMC_SAMPLER_MARKOV_JUMP && in->n &&
        (!in->H_cache || !in->cache_active || !out->H_cache || !out->cache_active))
        return fail("MarkovJumpHMC needs H_cache and cache_active");
    fill_outputs(p, o);
    if (int rc = fill_dist(p, dist)) return rc;
    p.Xin = in->X; p.Vin = in->V; p.Xout = out->X; p.Vout = out->V;
    p.Hc_in = in->H_cache; p.Hc_out = out->H_cache; p.ca_in = in->cache_active; p.ca_out = out->cache_active;
    p.n = in->n; p.ld = in->ld; p.n_iter = n_iter;
    if (p.n == 0) return 0;
    if (o->energy && (force_stream || !is_elementwise(dist->kind) || !fused_template_dim(dist->ndims)))
        return fail("outputs.energy is written by the register-resident and unfused kernels only");
    if (force_stream) {
        if (!stream_supported(dist->dtype, dist->kind, dist->ndims))
            return fail("no streaming kernel for this distribution / ndims (separable energies, ndims <= 128)");
        return check(launch_stream_kernel(dist->dtype, dist->kind, p, stream), "stream_sample_kernel");
    }
    if (is_elementwise(dist->kind)) {
        const int D = fused_template_dim(dist->ndims);
        fused_launch_fn fn = D ? find_fused(dist->dtype, dist->kind, D) : nullptr;
        if (fn) return check(fn(p, stream), "fused_sample_kernel");
        if (stream_supported(dist->dtype, dist->kind, dist->ndims))
            return check(launch_stream_kernel(dist->dtype, dist->kind, p, stream), "stream_sample_kernel");
        return fail("no fused kernel for this distribution / ndims (use the unfused path)");
    }
    if (!dense_supported(dist->dtype, dist->kind, dist->ndims, dist->nbasis))
        return fail("no fused dense kernel for this distribution / shape (use the unfused path)");
    return check(launch_dense(dist->dtype, dist->kind, p, stream), "dense_sample_kernel");
}

int mjhmc_sample_fused(const mjhmc_dist* dist, const mjhmc_hp* hp, const mjhmc_rng* rng,
                       const mjhmc_state* in, const mjhmc_state* out, int32_t n_iter,
                       const mjhmc_outputs* o, void* stream) {
    return sample_impl(dist, hp, rng, in, out, n_iter, o, stream, false);
}

int mjhmc_sample_stream(const mjhmc_dist* dist, const mjhmc_hp* hp, const mjhmc_rng* rng,
                        const mjhmc_state* in, const mjhmc_state* out, int32_t n_iter,
                        const mjhmc_outputs* o, void* stream) {
    return sample_impl(dist, hp, rng, in, out, n_iter, o, stream, true);
}

int mjhmc_energy(const mjhmc_dist* dist, const void* X, int64_t n, int64_t ld, void* E, void* stream) {
    if (check_dist(dist)) return -1;
    if (n < 0 || ld < n) return fail("bad n / ld");
    if (n && (!X || !E)) return fail("NULL argument");
    return check(launch_energy(dist->dtype, dist_params(dist), X, n, ld, E, (cudaStream_t)stream), "energy_kernel");
}

int mjhmc_gradient(const mjhmc_dist* dist, const void* X, int64_t n, int64_t ld, void* G, void* stream) {
    if (check_dist(dist)) return -1;
    if (n < 0 || ld < n) return fail("bad n / ld");
    if (n && (!X || !G)) return fail("NULL argument");
    return check(launch_gradient(dist->dtype, dist_params(dist), X, n, ld, G, (cudaStream_t)stream), "gradient_kernel");
}

int mjhmc_kinetic(int32_t dtype, int32_t ndims, const void* V, int64_t n, int64_t ld, void* EV, void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (n < 0 || ld < n || ndims <= 0) return fail("bad n / ld / ndims");
    if (n && (!V || !EV)) return fail("NULL argument");
    return check(launch_kinetic(dtype, ndims, V, n, ld, EV, (cudaStream_t)stream), "kinetic_kernel");
}

int mjhmc_kick_drift(int32_t dtype, int32_t ndims, void* X, void* V, const void* G, int64_t n, int64_t ld,
                     double epsilon, void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (n < 0 || ld < n || ndims <= 0) return fail("bad n / ld / ndims");
    if (n && (!X || !V || !G)) return fail("NULL argument");
    return check(launch_kick(dtype, ndims, X, V, G, n, ld, epsilon, true, (cudaStream_t)stream), "kick_drift");
}

int mjhmc_kick(int32_t dtype, int32_t ndims, void* V, const void* G, int64_t n, int64_t ld, double epsilon,
               void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (n < 0 || ld < n || ndims <= 0) return fail("bad n / ld / ndims");
    if (n && (!V || !G)) return fail("NULL argument");
    return check(launch_kick(dtype, ndims, nullptr, V, G, n, ld, epsilon, false, (cudaStream_t)stream), "kick");
}

int mjhmc_transition(int32_t dtype, int32_t ndims, const mjhmc_hp* hp, const mjhmc_rng* rng, int64_t n, int64_t ld,
                     const mjhmc_full_state* cur, const mjhmc_full_state* prop, const void* H_flf, void* H_cache,
                     uint8_t* cache_active, const mjhmc_outputs* o, void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (!hp || !rng || !cur || !prop || !o) return fail("NULL argument");
    if (!o->counters) return fail("outputs.counters is required");
    if (n < 0 || ld < n || ndims <= 0) return fail("bad n / ld / ndims");
    LaunchParams p;
    memset(&p, 0, sizeof p);
    if (fill_hp(p, hp)) return -1;
    if (fill_rng(p, rng)) return -1;
    fill_outputs(p, o);
    p.d = ndims; p.n = n; p.ld = ld; p.n_iter = 1;
    p.Hc_out = H_cache; p.ca_out = cache_active;
    if (hp->sampler == MJHMC_SAMPLER_MARKOV_JUMP && n && (!H_flf || !H_cache || !cache_active))
        return fail("MarkovJumpHMC transition needs H_flf, H_cache, cache_active");
    FullPtrs c{cur->X, cur->V, cur->G, cur->EX, cur->EV};
    FullPtrs q{prop->X, prop->V, prop->G, prop->EX, prop->EV};
    if (n && (!c.X || !c.V || !c.G || !c.EX || !c.EV || !q.X || !q.V || !q.G || !q.EX || !q.EV))
        return fail("state arrays must not be NULL");
    return check(launch_transition(dtype, p, c, q, H_flf, (cudaStream_t)stream), "transition_kernel");
}

// Counter block housekeeping runs as two tiny kernels: a reset that needs no host synchronisation, and a fold that
// writes the totals straight into mapped pinned host memory.  (cudaMemcpy of the block went through the device->host
// copy engine, where it queued behind the 64 MB sample copies of the pipelined sample(): every chunk waited up to 2 ms
// for its 2 KB of counters.)
__global__ void counters_reset_kernel(long long* counters) {
    const int i = threadIdx.x;
    if (i < MJHMC_COUNTER_ROWS * MJHMC_N_COUNTERS) {
        const int row = i / MJHMC_N_COUNTERS, c = i - row * MJHMC_N_COUNTERS;
        counters[i] = (row < MJHMC_COUNTER_STRIPES && c == MJHMC_CNT_FAIL) ? INT64_MAX : 0;
    }
}
__global__ void counters_fold_kernel(const long long* counters, long long* out) {
    const int c = threadIdx.x;
    if (c < MJHMC_N_COUNTERS) {
        long long v = c == MJHMC_CNT_FAIL ? INT64_MAX : 0;
        for (int s = 0; s < MJHMC_COUNTER_STRIPES; ++s) {
            const long long h = counters[s * MJHMC_N_COUNTERS + c];
            if (c == MJHMC_CNT_FAIL) v = h < v ? h : v; else v += h;
        }
        out[c] = v;
        __threadfence_system();
    }
}

int mjhmc_counters_reset(int64_t* counters, void* stream_) {
    if (!counters) return fail("NULL argument");
    static_assert(MJHMC_COUNTER_ROWS * MJHMC_N_COUNTERS <= 512, "one CTA resets the block");
    counters_reset_kernel<<<1, 512, 0, (cudaStream_t)stream_>>>((long long*)counters);
    return check(cudaGetLastError(), "counters_reset");
}

int mjhmc_counters_read(const int64_t* counters, int64_t* out_host, void* stream_) {
    if (!counters || !out_host) return fail("NULL argument");
    static thread_local long long* mapped = nullptr;           // mapped pinned, one per host thread
    if (!mapped && check(cudaHostAlloc((void**)&mapped, MJHMC_N_COUNTERS * sizeof(long long), cudaHostAllocMapped | cudaHostAllocPortable), "counters_read"))
        return -2;
    long long* dev_view = nullptr;
    if (check(cudaHostGetDevicePointer((void**)&dev_view, mapped, 0), "counters_read")) return -2;
    cudaStream_t stream = (cudaStream_t)stream_;
    counters_fold_kernel<<<1, 32, 0, stream>>>((const long long*)counters, dev_view);
    if (check(cudaGetLastError(), "counters_read")) return -2;
    if (check(cudaStreamSynchronize(stream), "counters_read")) return -2;
    for (int c = 0; c < MJHMC_N_COUNTERS; ++c) out_host[c] = mapped[c];
    return 0;
}

int64_t mjhmc_dense_tc_workspace_bytes(const mjhmc_dist* dist) {
    if (check_dist(dist)) return -1;
    if (dist->dtype != MJHMC_F32 || !dense_tc_supported(dist->kind, dist->ndims, dist->nbasis)) { fail("no tensor-core workspace for this distribution"); return -1; }
    return dense_tc_workspace_bytes(dist->kind, dist->ndims);
}

int mjhmc_dense_tc_prepare(const mjhmc_dist* dist, void* stream) {
    if (check_dist(dist)) return -1;
    if (dist->dtype != MJHMC_F32 || !dense_tc_supported(dist->kind, dist->ndims, dist->nbasis))
        return fail("the tensor-core workspace is for the fp32 dense Gaussian / ProductOfT (ndims <= 112)");
    if (!dist->ws) return fail("dist->ws (workspace) is NULL");
    return check(dense_tc_prepare(dist->kind, (const float*)dist->a0, (const float*)dist->a1, (const float*)dist->a2, dist->ndims,
                                  (void*)dist->ws, (cudaStream_t)stream), "tc_prep_kernel");
}

int64_t mjhmc_resample_scratch_bytes(int64_t m) { return m < 0 ? -1 : resample_scratch_bytes(m); }

int mjhmc_resample(int32_t dtype, int32_t ndims, const double* dwell, int64_t m, const double* r, int64_t m_out,
                   const void* samples, int64_t ld_in, void* out, int64_t ld_out, int64_t* idx_out, void* scratch,
                   void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (m < 0 || m_out < 0 || ndims <= 0 || ld_in < m || ld_out < m_out) return fail("bad sizes");
    if (m && m_out && (!dwell || !r || !samples || !out || !scratch)) return fail("NULL argument");
    return check(launch_resample(dtype, ndims, dwell, m, r, m_out, samples, ld_in, out, ld_out,
                                 (long long*)idx_out, scratch, (cudaStream_t)stream), "resample");
}

int mjhmc_autocorr(int32_t dtype, int32_t ndims, const void* samples, int64_t stride_k, int64_t stride_it, int64_t n,
                   int32_t T, int32_t n_lags, int32_t circular, double* ac, void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (n < 0 || T < 0 || n_lags < 0 || ndims <= 0) return fail("bad sizes");
    if (n && T && n_lags && (!samples || !ac)) return fail("NULL argument");
    return check(launch_autocorr(dtype, ndims, samples, stride_k, stride_it, n, T, n_lags, circular, ac,
                                 (cudaStream_t)stream), "autocorr");
}

int64_t mjhmc_autocorr_fft_scratch_bytes(int32_t T) { return autocorr_fft_supported(T, 1) ? autocorr_fft_scratch_bytes(T) : -1; }

int mjhmc_autocorr_fft(int32_t dtype, int32_t ndims, const void* samples, int64_t stride_k, int64_t stride_it, int64_t n,
                       int32_t T, int32_t n_lags, double* ac, void* scratch, void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (n < 0 || T < 0 || n_lags < 0 || n_lags > T || ndims <= 0) return fail("bad sizes");
    if (!autocorr_fft_supported(T, 1)) return fail("the FFT autocorrelation needs T = 2^m, 16 <= T <= 4096 (use mjhmc_autocorr)");
    if (n && n_lags && (!samples || !ac || !scratch)) return fail("NULL argument");
    return check(launch_autocorr_fft(dtype, ndims, samples, stride_k, stride_it, n, T, n_lags, ac, (double*)scratch,
                                     (cudaStream_t)stream), "autocorr_fft");
}

int mjhmc_ladder_visits(const uint8_t* choice, int64_t n_iter, int64_t n, int32_t K, int32_t* state, int64_t* visits,
                        void* stream) {
    if (n_iter < 0 || n < 0 || K < 0) return fail("bad sizes");
    if (n_iter && n && (!choice || !state || !visits)) return fail("NULL argument");
    return check(launch_ladder_visits(choice, n_iter, n, K, state, (long long*)visits, (cudaStream_t)stream), "ladder_visits");
}

int mjhmc_moments(int32_t dtype, const void* x, int64_t count, double* out, void* stream) {
    if (dtype != MJHMC_F32 && dtype != MJHMC_F64) return fail("bad dtype");
    if (count < 0) return fail("bad count");
    if (count && (!x || !out)) return fail("NULL argument");
    return check(launch_moments(dtype, x, count, out, (cudaStream_t)stream), "moments");
}

}  // extern "C"
