/*
 * mjhmc_b200 -- C ABI of the B200-native particle-parallel sampler loop.
 *
 * The reference (rueberger/MJHMC) is pure Python: there is no FFI in it.  The
 * drop-in boundary is therefore its Python class surface (kept by the host
 * package `mjhmc_b200`), and this C ABI sits *underneath* that surface.  Every
 * entry point below names the reference code it replaces (file:line relative to
 * /root/reference/mjhmc/).  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add to call these directly.
 *
 * Conventions
 *   - plain C symbols, plain pointers and sizes; no torch types
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`
 *   - the caller owns every buffer; the library allocates nothing persistent
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as
 *     void*); nothing synchronises except where stated
 *   - return 0 on success, <0 on error; mjhmc_last_error() gives the message
 *   - state arrays are (ndims, n) row-major with row stride `ld` (elements):
 *     the particle index is the fast axis, exactly the reference layout
 *     (samplers/hmc_state.py:20-39)
 */
#ifndef MJHMC_B200_H
#define MJHMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MJHMC_ABI_VERSION 2

/* dtype of the state arrays X, V, samples, H cache (arithmetic type of the path) */
enum { MJHMC_F32 = 0, MJHMC_F64 = 1 };

/* energy models with a fused device implementation (misc/distributions.py) */
enum {
    MJHMC_DIST_TEST_GAUSSIAN  = 0, /* distributions.py:357-362  p[0]=sigma                          */
    MJHMC_DIST_DIAG_GAUSSIAN  = 1, /* distributions.py:262-273 with diagonal J; a0 = diag(J) [ndims] */
    MJHMC_DIST_ROUGH_WELL     = 2, /* distributions.py:295-304  p[0]=scale1 p[1]=scale2              */
    MJHMC_DIST_FUNNEL         = 3, /* tf_distributions.py:143-147 (Neal's funnel)  p[0]=scale        */
    MJHMC_DIST_FUNNEL_LITERAL = 4, /* tf_distributions.py:158-165 as written        p[0]=scale        */
    MJHMC_DIST_DENSE_GAUSSIAN = 5, /* distributions.py:268-273 full J; a0 = (J+J^T)/2 [ndims x ndims];
                                      fp32: ws = workspace filled by mjhmc_dense_tc_prepare */
    MJHMC_DIST_PRODUCT_OF_T   = 6, /* distributions.py:420-433; a0=W [ndims x nbasis], a1=nu, a2=b;
                                      fp32: ws = workspace filled by mjhmc_dense_tc_prepare */
    MJHMC_DIST_MULTIMODAL     = 7  /* distributions.py:314-335  two Gaussians at -/+ 2*separation along dim 0; p[0]=separation */
};

/* sampler classes (samplers/markov_jump_hmc.py) */
enum {
    MJHMC_SAMPLER_DISCRETE        = 0, /* HMCBase / HMC / ControlHMC  :116-148 (differ only in p_flip, p_r, beta) */
    MJHMC_SAMPLER_CONTINUOUS_TIME = 1, /* ContinuousTimeHMC           :251-290 */
    MJHMC_SAMPLER_MARKOV_JUMP     = 2  /* MarkovJumpHMC               :355-415 */
};

enum { MJHMC_RNG_PHILOX = 0, MJHMC_RNG_INJECT = 1 };

/* counters block: int64[MJHMC_N_COUNTERS], accumulated (+=) by the kernels */
enum {
    MJHMC_CNT_L = 0, MJHMC_CNT_F = 1, MJHMC_CNT_FL = 2, MJHMC_CNT_R = 3, /* markov_jump_hmc.py:82-87 */
    MJHMC_CNT_E = 4, MJHMC_CNT_DEDX = 5,                                 /* distributions.py:44-48,62-75 */
    MJHMC_CNT_FAIL = 6,  /* first iteration (relative to the launch) with a non-finite rate, else INT64_MAX */
    MJHMC_CNT_EXEC = 7,  /* gradient evaluations actually executed on the device: < DEDX when the energy of
                            an FLF state is taken from the cache instead of being re-integrated */
    MJHMC_N_COUNTERS = 8
};
/* the kernels stripe their atomics over this many counter rows; mjhmc_counters_reduce folds them */
#define MJHMC_COUNTER_STRIPES 32
/* rows of the counter block: the stripes plus one row of work-queue heads for the persistent kernels */
#define MJHMC_COUNTER_ROWS (MJHMC_COUNTER_STRIPES + 1)

typedef struct mjhmc_dist {
    int32_t kind;       /* MJHMC_DIST_*  */
    int32_t dtype;      /* MJHMC_F32/F64 -- dtype of a0/a1/a2 and of the state */
    int32_t ndims;
    int32_t nbasis;     /* ProductOfT only */
    double  p[4];       /* scalar parameters, see MJHMC_DIST_* */
    const void *a0, *a1, *a2;   /* device parameter arrays, see MJHMC_DIST_* */
    const void *ws;             /* fp32 DENSE_GAUSSIAN / PRODUCT_OF_T: caller-owned device workspace of
                                   mjhmc_dense_tc_workspace_bytes(dist) bytes, filled once by mjhmc_dense_tc_prepare */
} mjhmc_dist;

/* hyper-parameters after the host-side derivation of markov_jump_hmc.py:67-80,189,197-200,221-223 */
typedef struct mjhmc_hp {
    int32_t sampler;            /* MJHMC_SAMPLER_* */
    int32_t num_leapfrog_steps;
    double  epsilon;
    double  beta;               /* as seen by HMCState.R (hmc_state.py:126): 1 for Control/CT/MJ */
    double  p_flip;
    double  p_r;                /* probability (discrete) or rate (CT/MJ) */
} mjhmc_hp;

/* random stream (DESIGN.md "Random streams"): counter-based Philox4x32-10, or
 * pre-drawn arrays indexed (attempt, slot, particle) for trajectory parity with
 * the reference's np.random call sites (hmc_state.py:126; markov_jump_hmc.py:125,132,138; utils.py:42) */
/* mjhmc_rng.flags.  LITERAL_RACE: evaluate all three exponential holding times of the continuous-time samplers in
 * fp64 exactly as misc/utils.py:15-49 writes them.  Default (0): the kernels first enclose the three times in
 * single-precision intervals and evaluate only the winner in fp64 when the intervals separate (csrc/common.cuh);
 * the choices and the stored holding times are bit-identical either way (tests/test_gpu_screen.py compares them). */
#define MJHMC_RNG_FLAG_LITERAL_RACE 1
/* REGISTER_STATE: run the register-resident form of the fused kernel also where the library would keep the particle
 * state in shared memory between trajectories (ndims >= 6, csrc/fused_elementwise.cuh: fused_stash_kernel).  Same
 * results bit for bit; the flag exists so the tests can compare the two. */
#define MJHMC_RNG_FLAG_REGISTER_STATE 2

typedef struct mjhmc_rng {
    int32_t  mode;              /* MJHMC_RNG_* */
    int32_t  flags;             /* MJHMC_RNG_FLAG_* */
    uint64_t seed;              /* PHILOX key */
    uint64_t attempt0;          /* attempt index of the first iteration of this launch */
    uint64_t particle0;         /* global index of local particle 0 (shard offset) */
    const double *Z;            /* INJECT: normals  [n_attempts][ndims][inj_ld] */
    const double *U;            /* INJECT: uniforms [n_attempts][3][inj_ld]     */
    const double *U0;           /* INJECT: batch coin [n_attempts]              */
    int64_t  inj_ld;            /* row length of Z and U (global particle count) */
} mjhmc_rng;

/* particle state.  Only X and V live in HBM between launches; EX, EV and dEdX of
 * hmc_state.py:28-39 are pure functions of (X, V) and are recomputed on chip.
 * The FLF cache of hmc_state.py:41-44,131-148 is one scalar + one flag per
 * particle (only H() of the cached state is read: markov_jump_hmc.py:367). */
typedef struct mjhmc_state {
    void    *X, *V;             /* (ndims, n), row stride ld */
    void    *H_cache;           /* (n,) dtype; MarkovJumpHMC only, else NULL */
    uint8_t *cache_active;      /* (n,) flags  MarkovJumpHMC only, else NULL: bit0 = the reference's cache_active,
                                   bit1 = H_cache holds a valid FLF energy (also set after an F move) */
    int64_t  n;                 /* particles in this shard */
    int64_t  ld;
} mjhmc_state;

/* where one launch writes its per-iteration results (any pointer may be NULL) */
typedef struct mjhmc_outputs {
    void    *samples;           /* dtype; element (k, it, i) at k*stride_k + it*stride_it + i */
    int64_t  stride_k, stride_it;
    double  *dwell;             /* (n_iter, n): dwelling time of every iteration (markov_jump_hmc.py:274,395) */
    double  *dwell_last;        /* (n,): dwelling time of the last iteration -> sampler.dwelling_times */
    uint8_t *choice;            /* (n_iter, n): operator taken; MJ 0=L 1=F 2=R, CT 0=F 1=FL 2=R,
                                   discrete bit0=accepted bit1=flipped bit2=R fired */
    int64_t *counters;          /* [MJHMC_COUNTER_ROWS][MJHMC_N_COUNTERS], +=; reset before every launch */
    double  *energy;            /* (n_iter, n): H() of the state after every iteration (hmc_state.py:80-84), what
                                   experiments/spectral.py:33-45,186-208 reads with sampler.state.H() after each
                                   sampling_iteration().  Register-resident and unfused kernels only (NULL elsewhere) */
} mjhmc_outputs;

const char *mjhmc_last_error(void);
int         mjhmc_abi_version(void);

/* 1 if (dist, dtype, ndims) has a fused kernel, 0 if the caller must use the unfused pieces */
int mjhmc_fused_supported(const mjhmc_dist *dist);

/* Replaces HMCBase.sample / sampling_iteration (markov_jump_hmc.py:116-173),
 * ContinuousTimeHMC.sampling_iteration (:251-290), MarkovJumpHMC.sampling_iteration
 * (:355-415) together with HMCState.L/F/FLF/R (hmc_state.py:86-129), draw_from and
 * min_idx (misc/utils.py:15-49): runs `n_iter` sampling iterations for all particles of
 * `in`, keeping each particle's position, momentum and gradient on chip across the
 * iterations, and writes the final state to `out` (may alias `in`).
 * The infinite-rate condition (utils.py:41-48) is reported in counters[MJHMC_CNT_FAIL];
 * the caller implements the batch-wide back-off (markov_jump_hmc.py:376-389). */
int mjhmc_sample_fused(const mjhmc_dist *dist, const mjhmc_hp *hp, const mjhmc_rng *rng,
                       const mjhmc_state *in, const mjhmc_state *out,
                       int32_t n_iter, const mjhmc_outputs *o, void *stream);

/* Same contract as mjhmc_sample_fused, forced onto the streaming kernel for SEPARABLE energies
 * (MJHMC_DIST_TEST_GAUSSIAN, _DIAG_GAUSSIAN, _ROUGH_WELL; ndims <= 128): persistent CTAs walk particle tiles whose
 * X / V boxes arrive through a TMA ring, several threads share one particle (the dims of hmc_state.py:86-91
 * evolve independently; only the energy sums of :46-50 couple them).  mjhmc_sample_fused picks it by itself
 * when ndims > 16; call it directly for short trajectories (L of a few steps), where the path is HBM-bound.
 * `in` and `out` must not alias unless they are identical.  mjhmc_stream_supported: 1 if a kernel exists.
 * mjhmc_stream_set_tma(0) forces the non-TMA loader (it is also taken when X / V are not 16-byte aligned
 * or ld * sizeof(dtype) is not a multiple of 16). */
int mjhmc_stream_supported(const mjhmc_dist *dist);
int mjhmc_sample_stream(const mjhmc_dist *dist, const mjhmc_hp *hp, const mjhmc_rng *rng,
                        const mjhmc_state *in, const mjhmc_state *out,
                        int32_t n_iter, const mjhmc_outputs *o, void *stream);
void mjhmc_stream_set_tma(int32_t enabled);
/* launch geometry of the calling thread's last streaming launch:
 * {TMA used, ring stages, grid, CTAs per SM, warps per particle column, dims per thread, dynamic smem bytes} */
void mjhmc_stream_last_launch(int64_t *out7_host);
/* developer probe: resident 256-thread CTAs per SM of a trivial kernel with `smem_bytes` of dynamic shared memory */
int mjhmc_stream_probe_blocks(int64_t smem_bytes);

/* Replaces Distribution.E_val / dEdX_val (distributions.py:62-81) for the built-in
 * energies and HMCState.update_EV (hmc_state.py:49-50).  E, EV: (n,) dtype; G: (ndims, n). */
int mjhmc_energy(const mjhmc_dist *dist, const void *X, int64_t n, int64_t ld, void *E, void *stream);
int mjhmc_gradient(const mjhmc_dist *dist, const void *X, int64_t n, int64_t ld, void *G, void *stream);
int mjhmc_kinetic(int32_t dtype, int32_t ndims, const void *V, int64_t n, int64_t ld, void *EV, void *stream);

/* Unfused pieces for energies that only exist as host callables (LambdaDistribution,
 * user subclasses): the gradient is evaluated by the caller between these launches.
 *   kick_drift : V += -eps/2 * G ; X += eps * V      (hmc_state.py:88-89)
 *   kick       : V += -eps/2 * G                     (hmc_state.py:91)
 * Arrays are (ndims, n) with row stride ld. */
int mjhmc_kick_drift(int32_t dtype, int32_t ndims, void *X, void *V, const void *G,
                     int64_t n, int64_t ld, double epsilon, void *stream);
int mjhmc_kick(int32_t dtype, int32_t ndims, void *V, const void *G,
               int64_t n, int64_t ld, double epsilon, void *stream);

/* The reference's full HMCState (hmc_state.py:20-39) as kept by the unfused path, where
 * EX / EV / dEdX come from host callables and so cannot be recomputed on chip. */
typedef struct mjhmc_full_state {
    void *X, *V, *G;            /* (ndims, n) row stride ld; G = dEdX(X) */
    void *EX, *EV;              /* (n,) */
} mjhmc_full_state;

/* Unfused transition: the tail of one sampling iteration (same device code as the tail of
 * mjhmc_sample_fused) given a proposal computed by the caller.
 *   cur   : current state, updated in place (HMCState.update, hmc_state.py:63-72)
 *   prop  : the L state (hmc_state.py:93-100, before any F)
 *   H_flf : (n,) energy of the FLF state of the uncached particles (MarkovJumpHMC; ignored where
 *           cache_active; NULL for the other samplers)
 *   H_cache / cache_active : the FLF cache (MarkovJumpHMC), updated in place
 * Writes one sample column block / dwell / choice through `o` (iteration 0) and adds the
 * l/f/fl/r counters; E/dEdX counters belong to the caller's callables. */
int mjhmc_transition(int32_t dtype, int32_t ndims, const mjhmc_hp *hp, const mjhmc_rng *rng,
                     int64_t n, int64_t ld, const mjhmc_full_state *cur, const mjhmc_full_state *prop,
                     const void *H_flf, void *H_cache, uint8_t *cache_active,
                     const mjhmc_outputs *o, void *stream);

/* fp32 dense-contraction energies on the tcgen05 tensor cores (replaces the np.dot / Theano contractions of
 * distributions.py:268-273 and :420-433 for fp32 states): the kernel consumes the matrix (S, or W of ProductOfT)
 * pre-tiled into 8-row x 16-byte core matrices and split into three bf16 planes (+ the per-expert tables of
 * ProductOfT).  The caller allocates mjhmc_dense_tc_workspace_bytes(dist) bytes, points dist->ws at them and calls
 * prepare once per distribution; sampler launches only read the workspace. */
int64_t mjhmc_dense_tc_workspace_bytes(const mjhmc_dist *dist);
int mjhmc_dense_tc_prepare(const mjhmc_dist *dist, void *stream);

/* Folds the striped counter rows into host int64[MJHMC_N_COUNTERS] (a one-warp kernel writing into mapped pinned
 * memory, then a stream synchronisation: no device->host copy engine involved). */
int mjhmc_counters_read(const int64_t *counters, int64_t *out_host, void *stream);
/* Resets a striped counter block (zeros, FAIL = INT64_MAX) in stream order; does not synchronise. */
int mjhmc_counters_reset(int64_t *counters, void *stream);

/* Replaces the resampling loop of ContinuousTimeHMC.sample (markov_jump_hmc.py:321-328):
 * idx[j] = first i with cumsum(dwell)[i] > r[j]  for sorted r; then out[:, j] = samples[:, idx[j]].
 * dwell: (m,) double; r: (m_out,) double sorted ascending, already scaled by sum(dwell);
 * samples: (ndims, m) dtype row stride ld_in; out: (ndims, m_out) row stride ld_out.
 * scratch: device buffer of mjhmc_resample_scratch_bytes(m) bytes. */
int64_t mjhmc_resample_scratch_bytes(int64_t m);
int mjhmc_resample(int32_t dtype, int32_t ndims, const double *dwell, int64_t m,
                   const double *r, int64_t m_out, const void *samples, int64_t ld_in,
                   void *out, int64_t ld_out, int64_t *idx_out, void *scratch, void *stream);

/* Replaces autocor.fft_autocor (misc/autocor.py:37-49, circular = 1) and slow_autocorrelation (:177-211,
 * circular = 0) as direct products:
 *   circular: ac[tau] += sum_{k,i,t}        x[k,i,t] * x[k,i,(t+tau) mod T]
 *   linear  : ac[tau] += sum_{k,i,t<T-tau}  x[k,i,t] * x[k,i,t+tau]
 * NOT normalised (the caller divides after the cross-GPU all-reduce).
 * samples element (k, t, i) at k*stride_k + t*stride_it + i.  ac: (n_lags,) double. */
int mjhmc_autocorr(int32_t dtype, int32_t ndims, const void *samples, int64_t stride_k, int64_t stride_it,
                   int64_t n, int32_t T, int32_t n_lags, int32_t circular, double *ac, void *stream);

/* The same circular sums as mjhmc_autocorr(circular = 1) the way the reference computes them (misc/autocor.py:37-49:
 * FFT along time, |.|^2, inverse FFT): one forward FFT per pair of series, the power spectrum summed over all series,
 * ONE inverse transform per call -- O(T log T) per series instead of O(T n_lags).  T must be a power of two,
 * 16 <= T <= 4096 (mjhmc_autocorr_fft_scratch_bytes returns -1 otherwise); n_lags <= T.  ac += sums (not normalised).
 * scratch: device buffer of mjhmc_autocorr_fft_scratch_bytes(T) bytes. */
int64_t mjhmc_autocorr_fft_scratch_bytes(int32_t T);
int mjhmc_autocorr_fft(int32_t dtype, int32_t ndims, const void *samples, int64_t stride_k, int64_t stride_it, int64_t n,
                       int32_t T, int32_t n_lags, double *ac, void *scratch, void *stream);

/* Replaces the counter-polling loop of experiments/spectral.py:107-131 (ladder_heatmap): every particle walks its
 * state ladder (samplers/algebraic_hmc.py:485-519: L moves k2 by +1 / -1 depending on the flip bit k1, F toggles k1,
 * R starts a new ladder at [0, 0]) through the operator choices a MarkovJumpHMC launch recorded (0 = L, 1 = F, 2 = R)
 * and counts the visits of node (k1, k2):  visits[k1 * (2 K + 1) + k2 + K] += 1 for |k2| <= K, visits[2 (2 K + 1)] counts
 * the steps outside the window.  state: (n, 2) int32 ladder position carried between calls (zeros to start).
 * choice: (n_iter, n) uint8.  visits: int64[2 (2 K + 1) + 1]. */
int mjhmc_ladder_visits(const uint8_t *choice, int64_t n_iter, int64_t n, int32_t K, int32_t *state, int64_t *visits,
                        void *stream);

/* Replaces the Welford loop of online_variance (misc/gen_mj_init.py:76-98) for one chunk of samples:
 * out[0] += sum x, out[1] += sum x^2 over `count` contiguous elements (double accumulation). */
int mjhmc_moments(int32_t dtype, const void *x, int64_t count, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MJHMC_B200_H */
