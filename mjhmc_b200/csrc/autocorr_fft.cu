// K7b: circular autocorrelation by the cross-correlation theorem -- what the reference does (misc/autocor.py:37-49:
// FFT along time with no padding, |.|^2, inverse FFT, mean over dims and particles).  The direct-product kernel of
// analysis.cu costs O(T * n_lags) per series (48 ms for 256 steps x 128 lags of 10 M series: longer than the sampling it
// analysed); this one costs O(T log T) per series and ONE inverse transform per launch:
//
//     ac[tau] = sum_series sum_t x[t] x[(t + tau) mod T] = (1/T) sum_k P[k] cos(2 pi k tau / T),
//     P[k]    = sum_series |FFT(x)[k]|^2
//
// so the per-series work is a forward FFT and an accumulation of the power spectrum; the inverse transform runs once
// on the T accumulated values (autocorr_fft_finish).  Two real series ride in one complex transform (z = a + i b:
// consecutive particles ARE the (re, im) pair in memory); |Z[k]|^2 = |A|^2 + |B|^2 + 2 Im(A conj B), and the cross
// term is odd in k, so it cancels against the even cos(2 pi k tau / T) in the final sum.
//
// Transform: in-place decimation-in-frequency, radix-4 stages (one radix-2 stage first when log2 T is odd) on a tile of
// B series pairs in shared memory, fp64 throughout (fp32 samples are widened on load; the parity tests hold the curve
// to 1e-10).  The output stays in digit-reversed order: only sum |Z|^2 per POSITION is accumulated, and the finishing
// kernel maps position -> frequency (freq_of_pos).  T must be a power of two, 16 <= T <= 4096; the host falls back to
// the direct kernel otherwise.
#include "common.cuh"
#include "analysis.h"

namespace mjhmc {

constexpr int kFftThreads = 256;

struct cplx { double re, im; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)}; }
__device__ __forceinline__ cplx mul_neg_i(cplx a) { return {a.im, -a.re}; }       // a * (-i)

// frequency index held at position p after the in-place DIF stages (radix 2 first when log2 T is odd, then radix 4):
// a radix-R stage on a block of length N leaves the frequencies k = j (mod R) in sub-block j
__host__ __device__ inline int freq_of_pos(int p, int Tn, int log2T) {
    int rem = p, len = Tn, mult = 1, k = 0;
    int stage_bits = (log2T & 1) ? 1 : 2;
    int bits_left = log2T;
    while (bits_left > 0) {
        const int R = 1 << stage_bits;
        const int sub = len / R;
        const int j = rem / sub;
        rem -= j * sub;
        k += j * mult;
        mult *= R;
        len = sub;
        bits_left -= stage_bits;
        stage_bits = 2;
    }
    return k;
}

template <typename T>
__global__ void __launch_bounds__(kFftThreads)
autocorr_fft_kernel(const T* __restrict__ samples, long long stride_k, long long stride_it, long long n, int Tn, int log2T,
                    int B, double* __restrict__ Q) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    cplx* W = reinterpret_cast<cplx*>(fft_smem);              // W[j] = exp(-2 pi i j / T), j < T
    double* Qb = reinterpret_cast<double*>(W + Tn);            // per-position power sums of this block
    cplx* Z = reinterpret_cast<cplx*>(Qb + Tn);                // [B][T + 1] series pairs (row pad: the loads of one time step hit distinct banks)
    const int ZS = Tn + 1;
    const int tid = threadIdx.x;
    for (int j = tid; j < Tn; j += kFftThreads) {
        double s, c;
        sincospi(-2.0 * (double)j / (double)Tn, &s, &c);
        W[j] = {c, s};
        Qb[j] = 0.0;
    }
    const T* base = samples + (long long)blockIdx.y * stride_k;
    const long long n_pairs = (n + 1) / 2;
    const long long n_tiles = (n_pairs + B - 1) / B;
    for (long long tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        __syncthreads();                                       // the previous tile is no longer read
        // ---- load: 2B consecutive particles per time step are B (re, im) pairs
        const long long i0 = tl * 2 * B;
        for (int e = tid; e < Tn * 2 * B; e += kFftThreads) {
            const int t = e / (2 * B), c = e - t * 2 * B;
            const long long i = i0 + c;
            const double val = i < n ? (double)base[(long long)t * stride_it + i] : 0.0;
            double* z = reinterpret_cast<double*>(Z + (size_t)(c >> 1) * ZS + t);
            z[c & 1] = val;
        }
        __syncthreads();
        // ---- in-place DIF
        int len = Tn;
        if (log2T & 1) {                                       // one radix-2 stage
            const int half = len >> 1;
            for (int e = tid; e < B * half; e += kFftThreads) {
                const int b = e / half, pos = e - b * half;
                cplx* z = Z + (size_t)b * ZS;
                const cplx a = z[pos], c2 = z[pos + half];
                z[pos] = cadd(a, c2);
                z[pos + half] = cmul(csub(a, c2), W[pos]);
            }
            len = half;
            __syncthreads();
        }
        while (len >= 4) {
            const int quarter = len >> 2, tw = Tn / len;
            const int per_series = Tn >> 2;                    // butterflies of one series in this stage
            for (int e = tid; e < B * per_series; e += kFftThreads) {
                const int b = e / per_series, r = e - b * per_series;
                const int blk = r / quarter, pos = r - blk * quarter;
                cplx* z = Z + (size_t)b * ZS + (size_t)blk * len + pos;
                const cplx z0 = z[0], z1 = z[quarter], z2 = z[2 * quarter], z3 = z[3 * quarter];
                const cplx a = cadd(z0, z2), bb = csub(z0, z2), c2 = cadd(z1, z3), dd = mul_neg_i(csub(z1, z3));
                z[0] = cadd(a, c2);
                z[quarter] = cmul(cadd(bb, dd), W[pos * tw]);
                z[2 * quarter] = cmul(csub(a, c2), W[2 * pos * tw]);
                z[3 * quarter] = cmul(csub(bb, dd), W[3 * pos * tw]);
            }
            len = quarter;
            __syncthreads();
        }
        // ---- power per position, summed over the pairs of the tile (positions are owned by threads: no atomics)
        for (int pp = tid; pp < Tn; pp += kFftThreads) {
            double s = 0.0;
            for (int b = 0; b < B; ++b) {
                const cplx v = Z[(size_t)b * ZS + pp];
                s = fma(v.re, v.re, fma(v.im, v.im, s));
            }
            Qb[pp] += s;
        }
    }
    __syncthreads();
    for (int pp = tid; pp < Tn; pp += kFftThreads)
        if (Qb[pp] != 0.0) atomicAdd(Q + pp, Qb[pp]);
}

// ac[tau] += (1/T) sum_p Q[p] cos(2 pi freq(p) tau / T)
__global__ void __launch_bounds__(256) autocorr_fft_finish(const double* __restrict__ Q, int Tn, int log2T, int n_lags,
                                                           double* __restrict__ ac) {
    extern __shared__ double fin[];                            // P in frequency order
    for (int pp = threadIdx.x; pp < Tn; pp += blockDim.x) fin[freq_of_pos(pp, Tn, log2T)] = Q[pp];
    __syncthreads();
    const int tau = blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= n_lags) return;
    double s = 0.0;
    for (int k = 0; k < Tn; ++k) {
        const int ph = (int)(((long long)k * tau) & (Tn - 1));          // k tau mod T (T is a power of two)
        s = fma(fin[k], cospi(2.0 * (double)ph / (double)Tn), s);
    }
    ac[tau] += s / (double)Tn;
}

static int ilog2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

bool autocorr_fft_supported(int Tn, int circular) { return circular && Tn >= 16 && Tn <= 4096 && ilog2_exact(Tn) > 0; }

long long autocorr_fft_scratch_bytes(int Tn) { return (long long)sizeof(double) * Tn; }

cudaError_t launch_autocorr_fft(int dtype, int d, const void* samples, long long stride_k, long long stride_it,
                                long long n, int Tn, int n_lags, double* ac, double* scratch, cudaStream_t s) {
    if (n == 0 || Tn == 0 || n_lags == 0) return cudaSuccess;
    const int log2T = ilog2_exact(Tn);
    if (log2T < 0) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * Tn, s);
    if (e != cudaSuccess) return e;
    // B series pairs per tile: as many as fit beside the twiddle table and the power sums (cap 8: 16 particles = 128 B rows)
    const size_t fixed = (size_t)Tn * (sizeof(cplx) + sizeof(double));
    int B = (int)((200 * 1024 - fixed) / ((size_t)(Tn + 1) * sizeof(cplx)));
    if (B > 8) B = 8;
    if (B < 1) return cudaErrorInvalidValue;
    const size_t smem = fixed + (size_t)B * (Tn + 1) * sizeof(cplx);
    const long long n_tiles = ((n + 1) / 2 + B - 1) / B;
    long long gx = (148 * 2 + d - 1) / d;
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)d);
    if (dtype == MJHMC_F64) {
        e = cudaFuncSetAttribute(autocorr_fft_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        autocorr_fft_kernel<double><<<grid, kFftThreads, smem, s>>>((const double*)samples, stride_k, stride_it, n, Tn, log2T, B, scratch);
    } else {
        e = cudaFuncSetAttribute(autocorr_fft_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        autocorr_fft_kernel<float><<<grid, kFftThreads, smem, s>>>((const float*)samples, stride_k, stride_it, n, Tn, log2T, B, scratch);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    autocorr_fft_finish<<<(n_lags + 255) / 256, 256, sizeof(double) * Tn, s>>>(scratch, Tn, log2T, n_lags, ac);
    return cudaGetLastError();
}

}  // namespace mjhmc
