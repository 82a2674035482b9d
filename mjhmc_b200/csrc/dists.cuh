// Analytic energies and gradients, register-resident form (one particle = one thread).
// Each functor evaluates a full ndims-vector held in registers; `d` (<= D) is the
// runtime dimension, rows k >= d are padding and hold zeros.
//
//   TestGaussian  misc/distributions.py:357-362   E = sum x^2 / (2 sigma^2),  g = x / sigma^2
//   DiagGaussian  misc/distributions.py:262-273   J = diag(j):  E = sum x (j x) / 2,  g = j x
//   RoughWell     misc/distributions.py:295-304   E = sum x^2/(2 s1^2) + cos(2 pi x / s2)
//                                                 g = x/s1^2 - sin(2 pi x / s2) 2 pi / s2
//   Funnel        misc/tf_distributions.py:143-147 (Neal) and :158-165 (literal graph)
//
// Sums run k = 0..d-1 sequentially, the order numpy's axis-0 reduction uses.
#pragma once
#include "common.cuh"

namespace mjhmc {

template <typename T> __device__ __forceinline__ T t_sin(T x);
template <> __device__ __forceinline__ double t_sin<double>(double x) { return sin(x); }
template <> __device__ __forceinline__ float t_sin<float>(float x) { return sinf(x); }
template <typename T> __device__ __forceinline__ T t_cos(T x);
template <> __device__ __forceinline__ double t_cos<double>(double x) { return cos(x); }
template <> __device__ __forceinline__ float t_cos<float>(float x) { return cosf(x); }
template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }

template <typename T, int D>
struct TestGaussianD {
    static constexpr int kind = MJHMC_DIST_TEST_GAUSSIAN;
    T inv_s2, inv_2s2;
    int d;
    __device__ __forceinline__ explicit TestGaussianD(const LaunchParams& p)
        : inv_s2((T)(1.0 / (p.dp[0] * p.dp[0]))), inv_2s2((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))), d(p.d) {}
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const {
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = x[k] * inv_s2;
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) s += x[k] * x[k];
        return s * inv_2s2;
    }
};

template <typename T, int D>
struct DiagGaussianD {
    static constexpr int kind = MJHMC_DIST_DIAG_GAUSSIAN;
    T j[D];
    __device__ __forceinline__ explicit DiagGaussianD(const LaunchParams& p) {
#pragma unroll
        for (int k = 0; k < D; ++k) j[k] = (k < p.d) ? ((const T*)p.a0)[k] : (T)0;
    }
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const {
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = j[k] * x[k];
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) s += x[k] * (j[k] * x[k]);
        return s * (T)0.5;
    }
};

template <typename T, int D>
struct RoughWellD {
    static constexpr int kind = MJHMC_DIST_ROUGH_WELL;
    T inv_s1sq, inv_2s1sq, c;   // c = 2 pi / scale2
    int d;
    __device__ __forceinline__ explicit RoughWellD(const LaunchParams& p)
        : inv_s1sq((T)(1.0 / (p.dp[0] * p.dp[0]))), inv_2s1sq((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))),
          c((T)(2.0 * 3.14159265358979323846 / p.dp[1])), d(p.d) {}
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const {
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = x[k] * inv_s1sq - t_sin<T>(x[k] * c) * c;
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k < d) s += x[k] * x[k] * inv_2s1sq + t_cos<T>(x[k] * c);
        return s;
    }
};

// LITERAL = false: Neal's funnel  E = x0^2/(2 s^2) + exp(-x0)/2 sum_k xk^2 + (d-1) x0/2
// LITERAL = true : the TF graph as written  E = -[(d-1) x0^2/s^2 + exp(-x0) sum_k xk^2]
template <typename T, int D, bool LITERAL>
struct FunnelD {
    static constexpr int kind = LITERAL ? MJHMC_DIST_FUNNEL_LITERAL : MJHMC_DIST_FUNNEL;
    T inv_s2, nk;
    __device__ __forceinline__ explicit FunnelD(const LaunchParams& p)
        : inv_s2((T)(1.0 / (p.dp[0] * p.dp[0]))), nk((T)(p.d - 1)) {}
    __device__ __forceinline__ T sumsq(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 1; k < D; ++k) s += x[k] * x[k];
        return s;
    }
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const {
        const T e = t_exp<T>(-x[0]);
        const T s = sumsq(x);
        if (LITERAL) {
            g[0] = (T)-2 * nk * x[0] * inv_s2 + e * s;
#pragma unroll
            for (int k = 1; k < D; ++k) g[k] = (T)-2 * x[k] * e;
        } else {
            g[0] = x[0] * inv_s2 - (T)0.5 * e * s + (T)0.5 * nk;
#pragma unroll
            for (int k = 1; k < D; ++k) g[k] = x[k] * e;
        }
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        const T e = t_exp<T>(-x[0]);
        const T s = sumsq(x);
        if (LITERAL) return -(nk * x[0] * x[0] * inv_s2 + e * s);
        return x[0] * x[0] * ((T)0.5 * inv_s2) + (T)0.5 * e * s + (T)0.5 * nk * x[0];
    }
};

}  // namespace mjhmc
