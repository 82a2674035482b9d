// Runtime-ndims kernels: energies / gradients of the built-in distributions as
// stand-alone launches (Distribution.E / dEdX, misc/distributions.py:62-81), the
// leapfrog pieces and the transition used when the energy is a host callable
// (LambdaDistribution, user subclasses: the "unfused" path), kinetic energy.
#include "common.cuh"
#include "dists.cuh"
#include "unfused.h"

namespace mjhmc {

constexpr int kThreads = 256;
static inline unsigned grid_for(long long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// ------------------------------------------------------------------ energies
template <typename T>
__global__ void __launch_bounds__(kThreads) energy_kernel(DistParams dp, const T* __restrict__ X, long long n,
                                                          long long ld, T* __restrict__ E) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = dp.d;
    T e = (T)0;
    switch (dp.kind) {
        case MJHMC_DIST_TEST_GAUSSIAN: {
            for (int k = 0; k < d; ++k) { const T x = X[k * ld + i]; e += x * x; }
            e *= (T)(1.0 / (2.0 * dp.p[0] * dp.p[0]));
        } break;
        case MJHMC_DIST_DIAG_GAUSSIAN: {
            const T* j = (const T*)dp.a0;
            for (int k = 0; k < d; ++k) { const T x = X[k * ld + i]; e += x * (j[k] * x); }
            e *= (T)0.5;
        } break;
        case MJHMC_DIST_ROUGH_WELL: {
            const T inv_2s1sq = (T)(1.0 / (2.0 * dp.p[0] * dp.p[0]));
            const T c_pi = (T)(2.0 / dp.p[1]);
            for (int k = 0; k < d; ++k) { const T x = X[k * ld + i]; e += x * x * inv_2s1sq + t_cospi<T>(x * c_pi); }
        } break;
        case MJHMC_DIST_FUNNEL:
        case MJHMC_DIST_FUNNEL_LITERAL: {
            const T inv_s2 = (T)(1.0 / (dp.p[0] * dp.p[0]));
            const T nk = (T)(d - 1);
            const T x0 = X[i];
            T s = (T)0;
            for (int k = 1; k < d; ++k) { const T x = X[k * ld + i]; s += x * x; }
            const T ex = t_exp<T>(-x0);
            if (dp.kind == MJHMC_DIST_FUNNEL_LITERAL) e = -(nk * x0 * x0 * inv_s2 + ex * s);
            else e = x0 * x0 * ((T)0.5 * inv_s2) + (T)0.5 * ex * s + (T)0.5 * nk * x0;
        } break;
        case MJHMC_DIST_MULTIMODAL: {                           // distributions.py:323-327
            const T s0 = (T)(2.0 * dp.p[0]);
            const T x0 = X[i];
            T rest = (T)0;
            for (int k = 1; k < d; ++k) { const T x = X[k * ld + i]; rest += x * x; }
            const T a = (x0 + s0) * (x0 + s0) + rest, b = (x0 - s0) * (x0 - s0) + rest;
            e = -t_log<T>(t_exp<T>(-a) + t_exp<T>(-b));
        } break;
        case MJHMC_DIST_DENSE_GAUSSIAN: {
            const T* S = (const T*)dp.a0;                       // (J + J^T)/2, row-major
            for (int k = 0; k < d; ++k) {
                T acc = (T)0;
                for (int j = 0; j < d; ++j) acc += S[k * d + j] * X[j * ld + i];
                e += X[k * ld + i] * acc;
            }
            e *= (T)0.5;
        } break;
        case MJHMC_DIST_PRODUCT_OF_T: {
            const T* W = (const T*)dp.a0; const T* nu = (const T*)dp.a1; const T* b = (const T*)dp.a2;
            const int nb = dp.nbasis;
            for (int j = 0; j < nb; ++j) {
                T y = (T)0;
                for (int k = 0; k < d; ++k) y += X[k * ld + i] * W[k * nb + j];
                y += b[j];
                const T r = y / nu[j];
                e += (nu[j] + (T)1) * (T)0.5 * (T)log(1.0 + (double)(r * r));
            }
        } break;
    }
    E[i] = e;
}

// ------------------------------------------------------------------ gradients
template <typename T>
__global__ void __launch_bounds__(kThreads) gradient_kernel(DistParams dp, const T* __restrict__ X, long long n,
                                                            long long ld, T* __restrict__ G) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = dp.d;
    switch (dp.kind) {
        case MJHMC_DIST_TEST_GAUSSIAN: {
            const T inv_s2 = (T)(1.0 / (dp.p[0] * dp.p[0]));
            for (int k = 0; k < d; ++k) G[k * ld + i] = X[k * ld + i] * inv_s2;
        } break;
        case MJHMC_DIST_DIAG_GAUSSIAN: {
            const T* j = (const T*)dp.a0;
            for (int k = 0; k < d; ++k) G[k * ld + i] = j[k] * X[k * ld + i];
        } break;
        case MJHMC_DIST_ROUGH_WELL: {
            const T inv_s1sq = (T)(1.0 / (dp.p[0] * dp.p[0]));
            const T c_pi = (T)(2.0 / dp.p[1]);
            T sc[9];
            for (int j = 0; j < (sizeof(T) == 8 ? 9 : 5); ++j) sc[j] = (T)dp.coef[j];
            for (int k = 0; k < d; ++k) { const T x = X[k * ld + i]; G[k * ld + i] = x * inv_s1sq + scaled_sin_halfturns(x, c_pi, sc); }
        } break;
        case MJHMC_DIST_FUNNEL:
        case MJHMC_DIST_FUNNEL_LITERAL: {
            const T inv_s2 = (T)(1.0 / (dp.p[0] * dp.p[0]));
            const T nk = (T)(d - 1);
            const T x0 = X[i];
            T s = (T)0;
            for (int k = 1; k < d; ++k) { const T x = X[k * ld + i]; s += x * x; }
            const T ex = t_exp<T>(-x0);
            if (dp.kind == MJHMC_DIST_FUNNEL_LITERAL) {
                G[i] = (T)-2 * nk * x0 * inv_s2 + ex * s;
                for (int k = 1; k < d; ++k) G[k * ld + i] = (T)-2 * X[k * ld + i] * ex;
            } else {
                G[i] = x0 * inv_s2 - (T)0.5 * ex * s + (T)0.5 * nk;
                for (int k = 1; k < d; ++k) G[k * ld + i] = X[k * ld + i] * ex;
            }
        } break;
        case MJHMC_DIST_MULTIMODAL: {                           // distributions.py:329-335
            const T s0 = (T)(2.0 * dp.p[0]);
            const T x0 = X[i];
            const T c = t_exp<T>((T)4 * s0 * x0);
            const T den = c + (T)1;
            G[i] = ((T)2 * ((x0 - s0) * c + s0 + x0)) / den;
            for (int k = 1; k < d; ++k) { const T x = X[k * ld + i]; G[k * ld + i] = ((T)2 * (x * c + x)) / den; }
        } break;
        case MJHMC_DIST_DENSE_GAUSSIAN: {
            const T* S = (const T*)dp.a0;
            for (int k = 0; k < d; ++k) {
                T acc = (T)0;
                for (int j = 0; j < d; ++j) acc += S[k * d + j] * X[j * ld + i];
                G[k * ld + i] = acc;
            }
        } break;
        default: break;
    }
}

// ProductOfT gradient in two passes through a (nbasis, n) scratch:  Y = W^T X + b -> G_j, then dEdX = W G.
template <typename T>
__global__ void __launch_bounds__(kThreads) pot_expert_kernel(DistParams dp, const T* __restrict__ X, long long n,
                                                              long long ld, T* __restrict__ Y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* W = (const T*)dp.a0; const T* nu = (const T*)dp.a1; const T* b = (const T*)dp.a2;
    const int nb = dp.nbasis, d = dp.d;
    for (int j = 0; j < nb; ++j) {
        T y = (T)0;
        for (int k = 0; k < d; ++k) y += X[k * ld + i] * W[k * nb + j];
        y += b[j];
        Y[(long long)j * n + i] = (nu[j] + (T)1) * y / (nu[j] * nu[j] + y * y);
    }
}
template <typename T>
__global__ void __launch_bounds__(kThreads) pot_backproject_kernel(DistParams dp, const T* __restrict__ Y, long long n,
                                                                   long long ld, T* __restrict__ G) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* W = (const T*)dp.a0;
    const int nb = dp.nbasis, d = dp.d;
    for (int k = 0; k < d; ++k) {
        T acc = (T)0;
        for (int j = 0; j < nb; ++j) acc += W[k * nb + j] * Y[(long long)j * n + i];
        G[k * ld + i] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) kinetic_kernel(int d, const T* __restrict__ V, long long n, long long ld,
                                                           T* __restrict__ EV) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T s = (T)0;
    for (int k = 0; k < d; ++k) { const T v = V[k * ld + i]; s += v * v; }
    EV[i] = s * (T)0.5;
}

// ------------------------------------------------------------------ leapfrog pieces (hmc_state.py:86-91)
template <typename T, bool DRIFT>
__global__ void __launch_bounds__(kThreads) kick_kernel(int d, T* __restrict__ X, T* __restrict__ V,
                                                        const T* __restrict__ G, long long n, long long ld,
                                                        T eps, T neg_half_eps) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const long long k = idx / n, i = idx - k * n;
    const long long o = k * ld + i;
    T v = V[o];
    v += neg_half_eps * G[o];
    V[o] = v;
    if (DRIFT) X[o] += eps * v;
}

// ------------------------------------------------------------------ transition
template <typename T>
__global__ void __launch_bounds__(kThreads) transition_kernel(const __grid_constant__ LaunchParams p, FullPtrs cur,
                                                              FullPtrs prop, const T* __restrict__ H_flf) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0;
    if (i < p.n) {
        T* X = (T*)cur.X; T* V = (T*)cur.V; T* G = (T*)cur.G; T* EX = (T*)cur.EX; T* EV = (T*)cur.EV;
        const T* Xp = (const T*)prop.X; const T* Vp = (const T*)prop.V; const T* Gp = (const T*)prop.G;
        const T* EXp = (const T*)prop.EX; const T* EVp = (const T*)prop.EV;
        const int d = p.d;
        const long long ld = p.ld;
        const T H = EX[i] + EV[i];
        const T Hl = EXp[i] + EVp[i];
        const unsigned long long attempt = p.attempt0;
        int take = 0;        // 0 none, +1 proposal as is, -1 proposal with flipped momentum
        bool flip = false, refresh = false;
        unsigned int choice = 0;
        double dwell = 0.0;
        bool failed = false;
        if (p.sampler == MJHMC_SAMPLER_MARKOV_JUMP) {
            uint8_t* ca = p.ca_out; T* Hc = (T*)p.Hc_out;
            const T Hflf = ca[i] ? Hc[i] : H_flf[i];
            const Decision dc = decide_mj(p, i, attempt, (double)(H - Hl), (double)(H - Hflf), p.dwell != nullptr || p.dwell_last != nullptr);
            if (dc.fail) { report_failure(p, 0); failed = true; }
            else {
                choice = dc.choice; dwell = dc.dwell;
                if (choice == 0) { Hc[i] = H; ca[i] = 3; take = 1; n_l = 1; }
                else if (choice == 1) { flip = true; ca[i] = 0; n_f = 1; }
                else { refresh = true; ca[i] = 0; n_r = 1; }
            }
        } else if (p.sampler == MJHMC_SAMPLER_CONTINUOUS_TIME) {
            const Decision dc = decide_ct(p, i, attempt, (double)(H - Hl), p.dwell != nullptr || p.dwell_last != nullptr);
            if (dc.fail) { report_failure(p, 0); failed = true; }
            else {
                choice = dc.choice; dwell = dc.dwell;
                if (choice == 1) { take = -1; n_fl = 1; }
                else if (choice == 0) { flip = true; n_f = 1; }
                else { refresh = true; n_r = 1; }
            }
        } else {
            const Decision dc = decide_discrete(p, i, attempt, (double)(H - Hl), draw_coin(p, attempt) < p.p_r);
            choice = dc.choice;
            const bool acc = choice & 1u;
            flip = choice & 2u;
            refresh = choice & 4u;
            if (acc) take = -1;
            n_l = acc && flip; n_f = flip && !acc; n_fl = acc && !flip; n_r = refresh;
        }
        if (!failed) {
            if (take != 0) {
                for (int k = 0; k < d; ++k) {
                    X[k * ld + i] = Xp[k * ld + i];
                    V[k * ld + i] = take > 0 ? Vp[k * ld + i] : -Vp[k * ld + i];
                    G[k * ld + i] = Gp[k * ld + i];
                }
                EX[i] = EXp[i]; EV[i] = EVp[i];
            }
            if (flip) for (int k = 0; k < d; ++k) V[k * ld + i] = -V[k * ld + i];
            if (refresh) {
                T s = (T)0;
                for (int j = 0; 2 * j < d; ++j) {
                    double z0, z1;
                    normal_pair(p, i, attempt, j, d, z0, z1);
                    T v = V[(2 * j) * ld + i] * (T)p.r_keep + (T)z0 * (T)p.r_mix;
                    V[(2 * j) * ld + i] = v; s += v * v;
                    if (2 * j + 1 < d) {
                        v = V[(2 * j + 1) * ld + i] * (T)p.r_keep + (T)z1 * (T)p.r_mix;
                        V[(2 * j + 1) * ld + i] = v; s += v * v;
                    }
                }
                EV[i] = s * (T)0.5;
            }
            if (p.samples) {
                T* S = (T*)p.samples + i;
                for (int k = 0; k < d; ++k) S[(long long)k * p.s_stride_k] = X[k * ld + i];
            }
            if (p.dwell) p.dwell[i] = dwell;
            if (p.dwell_last && p.sampler != MJHMC_SAMPLER_DISCRETE) p.dwell_last[i] = dwell;
            if (p.choice) p.choice[i] = (uint8_t)choice;
            if (p.energy) p.energy[i] = (double)(EX[i] + EV[i]);
        }
    }
    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, 0u, 0u};
    flush_counters(p.counters, loc, 0ull);
}

// ------------------------------------------------------------------ host launchers
template <typename T>
static cudaError_t energy_T(const DistParams& dp, const void* X, long long n, long long ld, void* E, cudaStream_t s) {
    energy_kernel<T><<<grid_for(n), kThreads, 0, s>>>(dp, (const T*)X, n, ld, (T*)E);
    return cudaGetLastError();
}
cudaError_t launch_energy(int dtype, const DistParams& dp, const void* X, long long n, long long ld, void* E,
                          cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    return dtype == MJHMC_F64 ? energy_T<double>(dp, X, n, ld, E, s) : energy_T<float>(dp, X, n, ld, E, s);
}

template <typename T>
static cudaError_t gradient_T(const DistParams& dp, const void* X, long long n, long long ld, void* G, cudaStream_t s) {
    if (dp.kind == MJHMC_DIST_PRODUCT_OF_T) {
        T* Y = nullptr;
        cudaError_t e = cudaMallocAsync((void**)&Y, sizeof(T) * (size_t)dp.nbasis * (size_t)n, s);
        if (e != cudaSuccess) return e;
        pot_expert_kernel<T><<<grid_for(n), kThreads, 0, s>>>(dp, (const T*)X, n, ld, Y);
        pot_backproject_kernel<T><<<grid_for(n), kThreads, 0, s>>>(dp, Y, n, ld, (T*)G);
        e = cudaGetLastError();
        cudaFreeAsync(Y, s);
        return e;
    }
    gradient_kernel<T><<<grid_for(n), kThreads, 0, s>>>(dp, (const T*)X, n, ld, (T*)G);
    return cudaGetLastError();
}
cudaError_t launch_gradient(int dtype, const DistParams& dp, const void* X, long long n, long long ld, void* G,
                            cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    return dtype == MJHMC_F64 ? gradient_T<double>(dp, X, n, ld, G, s) : gradient_T<float>(dp, X, n, ld, G, s);
}

cudaError_t launch_kinetic(int dtype, int d, const void* V, long long n, long long ld, void* EV, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    if (dtype == MJHMC_F64) kinetic_kernel<double><<<grid_for(n), kThreads, 0, s>>>(d, (const double*)V, n, ld, (double*)EV);
    else kinetic_kernel<float><<<grid_for(n), kThreads, 0, s>>>(d, (const float*)V, n, ld, (float*)EV);
    return cudaGetLastError();
}

cudaError_t launch_kick(int dtype, int d, void* X, void* V, const void* G, long long n, long long ld, double eps,
                        bool drift, cudaStream_t s) {
    if (n == 0 || d == 0) return cudaSuccess;
    const unsigned grid = grid_for(n * d);
    if (dtype == MJHMC_F64) {
        if (drift) kick_kernel<double, true><<<grid, kThreads, 0, s>>>(d, (double*)X, (double*)V, (const double*)G, n, ld, eps, -eps / 2.0);
        else kick_kernel<double, false><<<grid, kThreads, 0, s>>>(d, nullptr, (double*)V, (const double*)G, n, ld, eps, -eps / 2.0);
    } else {
        if (drift) kick_kernel<float, true><<<grid, kThreads, 0, s>>>(d, (float*)X, (float*)V, (const float*)G, n, ld, (float)eps, (float)(-eps / 2.0));
        else kick_kernel<float, false><<<grid, kThreads, 0, s>>>(d, nullptr, (float*)V, (const float*)G, n, ld, (float)eps, (float)(-eps / 2.0));
    }
    return cudaGetLastError();
}

cudaError_t launch_transition(int dtype, const LaunchParams& p, const FullPtrs& cur, const FullPtrs& prop,
                              const void* H_flf, cudaStream_t s) {
    if (p.n == 0) return cudaSuccess;
    if (dtype == MJHMC_F64) transition_kernel<double><<<grid_for(p.n), kThreads, 0, s>>>(p, cur, prop, (const double*)H_flf);
    else transition_kernel<float><<<grid_for(p.n), kThreads, 0, s>>>(p, cur, prop, (const float*)H_flf);
    return cudaGetLastError();
}

}  // namespace mjhmc
