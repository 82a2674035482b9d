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

template <typename T> __device__ __forceinline__ T t_cospi(T x);
template <> __device__ __forceinline__ double t_cospi<double>(double x) { return cospi(x); }
template <> __device__ __forceinline__ float t_cospi<float>(float x) { return cospif(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }
template <typename T> __device__ __forceinline__ T t_exp(T x);
// exp(x) in fp64 for the leapfrog loop of the Funnel (tf_distributions.py:161-165: one exp per gradient): k = rint(x
// log2 e) by the magic-number add, r = x - k ln 2 (two-term Cody-Waite with FMAs, |r| <= 0.3466), exp(r) by a degree-11
// polynomial interpolated at Chebyshev nodes (truncation 1.6e-17 relative), 2^k added to the exponent field.  Maximum
// error 0.6 ulp against mpmath (tools/probe/exp_poly_check.py).  The library exp() is the same scheme, but its
// constants are literals that ptxas re-materialises with two UMOVs each on every call: 22 of its ~90 instructions,
// 16 % of everything the Funnel kernel issued (profiles/r2_fused_funnel10d_cthmc_v1_shared_refresh.txt).  Here they are operands
// read straight from the constant bank.  |x| >= 700 and NaN take the library path.
static __constant__ double kExpC[15] = {
    1.4426950408889634, -0.6931471805599453, -2.3190468138462996e-17,      // log2 e, -ln2 hi, -ln2 lo
    1.0, 1.0, 0.5000000000000019, 0.1666666666666668, 0.04166666666648795, 0.008333333333319589,
    0.0013888888952352863, 0.00019841269890076403, 2.4801485441561313e-05, 2.755724088722987e-06,
    2.763265472252779e-07, 2.5110049204818658e-08};
static __device__ __noinline__ double exp_out_of_range(double x) { return exp(x); }
__device__ __forceinline__ double exp_poly(double x) {
    const double magic = 6755399441055744.0;          // 1.5 * 2^52
    const double t = fma(x, kExpC[0], magic);
    const double k = t - magic;
    double r = fma(k, kExpC[1], x);
    r = fma(k, kExpC[2], r);
#ifndef MJ_EXP_ESTRIN
    double s = kExpC[14];
    s = fma(s, r, kExpC[13]);
    s = fma(s, r, kExpC[12]);
    s = fma(s, r, kExpC[11]);
    s = fma(s, r, kExpC[10]);
    s = fma(s, r, kExpC[9]);
    s = fma(s, r, kExpC[8]);
    s = fma(s, r, kExpC[7]);
    s = fma(s, r, kExpC[6]);
    s = fma(s, r, kExpC[5]);
    s = fma(s, r, kExpC[4]);
    s = fma(s, r, kExpC[3]);
#else
    // Estrin form of the same polynomial: 1 + (r + r^2 T(r)), T of degree 9 in pairs -- dependency depth 6 instead of
    // 11 (the exp is the critical path of a Funnel leapfrog step), three multiplications more; 0.7 ulp
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p0 = fma(kExpC[6], r, kExpC[5]), p1 = fma(kExpC[8], r, kExpC[7]), p2 = fma(kExpC[10], r, kExpC[9]);
    const double p3 = fma(kExpC[12], r, kExpC[11]), p4 = fma(kExpC[14], r, kExpC[13]);
    const double q0 = fma(p1, r2, p0), q1 = fma(p3, r2, p2);
    double s = fma(p4, r8, fma(q1, r4, q0));
    s = 1.0 + fma(s, r2, r);
#endif
    double res = __hiloint2double(__double2hiint(s) + (__double2loint(t) << 20), __double2loint(s));
    if (!(fabs(x) < 700.0)) res = exp_out_of_range(x);
    return res;
}
template <> __device__ __forceinline__ double t_exp<double>(double x) {
#ifdef MJ_LIB_EXP
    return exp(x);
#else
    return exp_poly(x);
#endif
}
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }

// sin(pi u) for u = x w in half-turns, branch free and table free (the fp64 leapfrog loop of RoughWell is
// nothing but this function).  The product x w is never rounded: k = rint(u) comes out of fma(x, w, magic) by the
// magic-number add and f = u - k in [-1/2, 1/2] out of a second FMA (one instruction fewer than rounding u first),
// sin(pi u) = (-1)^k sin(pi f); the parity of k is the low mantissa bit of (u + magic) and is
// XOR-ed into the sign of f on the integer pipe.  sin(pi f) = f Q(f^2) with Q fitted at Chebyshev
// nodes on [0, 1/4] (max relative error 3.9e-17 before rounding in fp64, 6.6e-9 in fp32).
// `c` holds the Q coefficients pre-multiplied by the caller's scale (host side, api.cu), so the
// return value is scale * sin(pi u).  Valid for |u| < 2^51 (fp64) / 2^22 (fp32); NaN/Inf -> NaN.
constexpr int kSinCoefF64 = 9, kSinCoefF32 = 5;
__device__ __forceinline__ double scaled_sin_halfturns(double x, double w, const double* __restrict__ c) {
    const double magic = 6755399441055744.0;          // 1.5 * 2^52
    const double y = fma(x, w, magic);
    const double f = fma(x, w, -(y - magic));
    const int sign = __double2loint(y) << 31;
    const double fs = __hiloint2double(__double2hiint(f) ^ sign, __double2loint(f));
    const double z = f * f;
    double q = c[8];
    q = fma(q, z, c[7]);
    q = fma(q, z, c[6]);
    q = fma(q, z, c[5]);
    q = fma(q, z, c[4]);
    q = fma(q, z, c[3]);
    q = fma(q, z, c[2]);
    q = fma(q, z, c[1]);
    q = fma(q, z, c[0]);
    return q * fs;
}
__device__ __forceinline__ float scaled_sin_halfturns(float x, float w, const float* __restrict__ c) {
    const float magic = 12582912.0f;                   // 1.5 * 2^23
    const float y = fmaf(x, w, magic);
    const float f = fmaf(x, w, -(y - magic));
    const float fs = __int_as_float(__float_as_int(f) ^ (__float_as_int(y) << 31));
    const float z = f * f;
    float q = c[4];
    q = fmaf(q, z, c[3]);
    q = fmaf(q, z, c[2]);
    q = fmaf(q, z, c[1]);
    q = fmaf(q, z, c[0]);
    return q * fs;
}

// cos(pi u) for u in half-turns with the same reduction: cos(pi (k + f)) = (-1)^k cos(pi f), cos(pi f) = C(f^2) with
// C interpolated at Chebyshev nodes on [0, 1/4] (max abs error 2.8e-16 against the exact cos(pi u) in fp64 arithmetic).
// A third of the instructions of cospi(); it is the RoughWell energy (distributions.py:295-299), evaluated once per
// trajectory.  Valid for |u| < 2^51.
__device__ __forceinline__ double cos_halfturns(double x, double w) {
    const double magic = 6755399441055744.0;          // 1.5 * 2^52
    const double y = fma(x, w, magic);
    const double f = fma(x, w, -(y - magic));
    const int sign = __double2loint(y) << 31;
    const double z = f * f;
    double q = 4.14956435394258569e-06;
    q = fma(q, z, -1.04566553387487379e-04);
    q = fma(q, z, 1.92955627248171321e-03);
    q = fma(q, z, -2.58068887370022024e-02);
    q = fma(q, z, 2.35330630129532842e-01);
    q = fma(q, z, -1.33526276884344708e+00);
    q = fma(q, z, 4.05871212641649670e+00);
    q = fma(q, z, -4.93480220054467633e+00);
    q = fma(q, z, 1.0);
    return __hiloint2double(__double2hiint(q) ^ sign, __double2loint(q));
}
// scale * sin(pi u) and cos(pi u) of the same argument: one range reduction for both (the RoughWell energy is asked
// for where a gradient has just been evaluated -- both ends of a trajectory).  Operation for operation the two
// functions above, so the values are bitwise theirs.
__device__ __forceinline__ double scaled_sincos_halfturns(double x, double w, const double* __restrict__ c, double& cosv) {
    const double magic = 6755399441055744.0;
    const double y = fma(x, w, magic);
    const double f = fma(x, w, -(y - magic));
    const int sign = __double2loint(y) << 31;
    const double fs = __hiloint2double(__double2hiint(f) ^ sign, __double2loint(f));
    const double z = f * f;
    double q = c[8];
    q = fma(q, z, c[7]);
    q = fma(q, z, c[6]);
    q = fma(q, z, c[5]);
    q = fma(q, z, c[4]);
    q = fma(q, z, c[3]);
    q = fma(q, z, c[2]);
    q = fma(q, z, c[1]);
    q = fma(q, z, c[0]);
    double r = 4.14956435394258569e-06;
    r = fma(r, z, -1.04566553387487379e-04);
    r = fma(r, z, 1.92955627248171321e-03);
    r = fma(r, z, -2.58068887370022024e-02);
    r = fma(r, z, 2.35330630129532842e-01);
    r = fma(r, z, -1.33526276884344708e+00);
    r = fma(r, z, 4.05871212641649670e+00);
    r = fma(r, z, -4.93480220054467633e+00);
    r = fma(r, z, 1.0);
    cosv = __hiloint2double(__double2hiint(r) ^ sign, __double2loint(r));
    return q * fs;
}
template <typename T> __device__ __forceinline__ T rw_cos_halfturns(T x, T w);
template <> __device__ __forceinline__ double rw_cos_halfturns<double>(double x, double w) {
#ifdef MJ_LIB_COSPI
    return cospi(x * w);
#else
    return cos_halfturns(x, w);
#endif
}
template <> __device__ __forceinline__ float rw_cos_halfturns<float>(float x, float w) { return cospif(x * w); }

template <typename T, int D>
struct TestGaussianD {
    static constexpr int kind = MJHMC_DIST_TEST_GAUSSIAN;
    T inv_s2, inv_2s2;
    int d;
    __device__ __forceinline__ explicit TestGaussianD(const LaunchParams& p)
        : inv_s2((T)(1.0 / (p.dp[0] * p.dp[0]))), inv_2s2((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))), d(p.d) {}
    // dims k0 .. k0+nd-1 of the particle (streaming kernel: several threads share one particle)
    __device__ __forceinline__ TestGaussianD(const LaunchParams& p, int k0, int nd)
        : inv_s2((T)(1.0 / (p.dp[0] * p.dp[0]))), inv_2s2((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))), d(nd) {}
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
    __device__ __forceinline__ DiagGaussianD(const LaunchParams& p, int k0, int nd) {
#pragma unroll
        for (int k = 0; k < D; ++k) j[k] = (k < nd) ? ((const T*)p.a0)[k0 + k] : (T)0;
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
    static constexpr bool kLinear = false;       // stream_separable.cuh: no folded-kick form
    static constexpr int kNC = sizeof(T) == 8 ? kSinCoefF64 : kSinCoefF32;
    T inv_s1sq, inv_2s1sq, c_pi;      // c_pi = 2 / scale2 (argument of sin / cos in half-turns)
    T sc[kNC];                        // Q coefficients times -2 pi / scale2
    int d;
    __device__ __forceinline__ explicit RoughWellD(const LaunchParams& p)
        : inv_s1sq((T)(1.0 / (p.dp[0] * p.dp[0]))), inv_2s1sq((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))),
          c_pi((T)(2.0 / p.dp[1])), d(p.d) {
#pragma unroll
        for (int j = 0; j < kNC; ++j) sc[j] = (T)p.coef[j];
    }
    __device__ __forceinline__ RoughWellD(const LaunchParams& p, int k0, int nd)
        : inv_s1sq((T)(1.0 / (p.dp[0] * p.dp[0]))), inv_2s1sq((T)(1.0 / (2.0 * p.dp[0] * p.dp[0]))),
          c_pi((T)(2.0 / p.dp[1])), d(nd) {
#pragma unroll
        for (int j = 0; j < kNC; ++j) sc[j] = (T)p.coef[j];
    }
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const {
        // x/s1^2 - sin(2 pi x / s2) 2 pi / s2, the sine evaluated in half-turns with the factor folded in
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = x[k] * inv_s1sq + scaled_sin_halfturns(x[k], c_pi, sc);
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k < d) s += x[k] * x[k] * inv_2s1sq + rw_cos_halfturns<T>(x[k], c_pi);
        return s;
    }
#ifndef MJ_LIB_COSPI
    // gradient and energy at the same point (grad_aux / energy_after below): the auxiliary value IS the energy
    static constexpr bool kGradAux = sizeof(T) == 8;
    __device__ __forceinline__ T grad_aux(const T (&x)[D], T (&g)[D]) const {
        if constexpr (sizeof(T) == 8) {
            T s = (T)0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double cv;
                g[k] = x[k] * inv_s1sq + scaled_sincos_halfturns(x[k], c_pi, sc, cv);
                if (k < d) s += x[k] * x[k] * inv_2s1sq + cv;
            }
            return s;
        } else {
            grad(x, g);
            return (T)0;
        }
    }
    __device__ __forceinline__ T energy_with(const T (&x)[D], T e) const { return e; }
#endif
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
    // grad_aux returns exp(-x0): the energy at the end of a trajectory is asked for at the point of the last gradient
    // (grad_aux / energy_after below), one exp per trajectory fewer
    static constexpr bool kGradAux = true;
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const { grad_aux(x, g); }
    __device__ __forceinline__ T grad_aux(const T (&x)[D], T (&g)[D]) const {
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
        return e;
    }
    __device__ __forceinline__ T energy_with(const T (&x)[D], T e) const {
        const T s = sumsq(x);
        if (LITERAL) return -(nk * x[0] * x[0] * inv_s2 + e * s);
        return x[0] * x[0] * ((T)0.5 * inv_s2) + (T)0.5 * e * s + (T)0.5 * nk * x[0];
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const { return energy_with(x, t_exp<T>(-x[0])); }
};

// MultimodalGaussian (distributions.py:314-335): the separation vector is (2 sep, 0, ..., 0), so
//   E = -log( exp(-|x + s|^2) + exp(-|x - s|^2) ),   c = exp(4 s.x),   g = 2 ((x - s) c + s + x) / (c + 1)
// written as in the reference (the two exponentials under the log, the common factor c over all dims).
template <typename T, int D>
struct MultimodalD {
    static constexpr int kind = MJHMC_DIST_MULTIMODAL;
    static constexpr bool kLinear = false;
    T s0;
    __device__ __forceinline__ explicit MultimodalD(const LaunchParams& p) : s0((T)(2.0 * p.dp[0])) {}
    __device__ __forceinline__ void grad(const T (&x)[D], T (&g)[D]) const {
        const T c = t_exp<T>((T)4 * s0 * x[0]);
        const T den = c + (T)1;
        g[0] = ((T)2 * ((x[0] - s0) * c + s0 + x[0])) / den;
#pragma unroll
        for (int k = 1; k < D; ++k) g[k] = ((T)2 * (x[k] * c + x[k])) / den;
    }
    __device__ __forceinline__ T energy(const T (&x)[D]) const {
        T rest = (T)0;
#pragma unroll
        for (int k = 1; k < D; ++k) rest += x[k] * x[k];
        const T a = (x[0] + s0) * (x[0] + s0) + rest;
        const T b = (x[0] - s0) * (x[0] - s0) + rest;
        return -t_log<T>(t_exp<T>(-a) + t_exp<T>(-b));
    }
};

// Energies that share work with their gradient (Dist::kGradAux): grad_aux returns it, energy_with takes it.
template <class Dist, class = void> struct grad_has_aux { static constexpr bool value = false; };
template <class Dist> struct grad_has_aux<Dist, decltype((void)Dist::kGradAux)> { static constexpr bool value = Dist::kGradAux; };

template <class Dist, typename T, int D>
__device__ __forceinline__ T grad_aux(const Dist& dist, const T (&x)[D], T (&g)[D]) {
    if constexpr (grad_has_aux<Dist>::value) { return dist.grad_aux(x, g); }
    else { dist.grad(x, g); return (T)0; }
}
// energy at the point where the gradient that returned `aux` was evaluated
template <class Dist, typename T, int D>
__device__ __forceinline__ T energy_after(const Dist& dist, const T (&x)[D], T aux) {
    if constexpr (grad_has_aux<Dist>::value) { return dist.energy_with(x, aux); }
    else { return dist.energy(x); }
}

}  // namespace mjhmc
