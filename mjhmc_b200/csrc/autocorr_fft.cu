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

constexpr int kFftThreads = 512;

struct __align__(16) cplx { double re, im; };

// XOR swizzle of element indices inside a series: 8 consecutive 16-byte elements cover the 32 banks; without it the late
// stages (butterflies 1, 4, 16 elements apart: lane stride 64 .. 1024 bytes) put a whole warp on 2 bank groups
// (1.2e9 bank conflicts, 72 % short-scoreboard stalls in profiles/r2k)
__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 7); }
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

// One tile = B series pairs = 2B consecutive particles x T steps.  A thread owns column c = tid mod 2B of the tile and
// the rows tid / 2B + k * (threads / 2B): its (at most kFftPre) samples of the NEXT tile are fetched into registers before
// the FFT stages of the current one start, so the global-memory latency hides behind the transform.
constexpr int kFftPre = 32;

template <typename T>
__global__ void __launch_bounds__(kFftThreads)
autocorr_fft_kernel(const T* __restrict__ samples, long long stride_k, long long stride_it, long long n, int Tn, int log2T,
                    int log2B, double* __restrict__ Q) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    // twiddles, one contiguous table per stage: radix-2 stage W2[pos] = w_T^pos (pos < T/2); a radix-4 stage on blocks of
    // length len holds {w_len^pos, w_len^2pos, w_len^3pos} for pos < len/4 (lane-consecutive reads, no strides)
    cplx* W = reinterpret_cast<cplx*>(fft_smem);              // 2T entries
    double* Qb = reinterpret_cast<double*>(W + 2 * Tn);        // per-position power sums of this block
    cplx* Z = reinterpret_cast<cplx*>(Qb + Tn);                // [B][T + 1] series pairs, elements swizzled inside a row
    const int ZS = Tn + 1;
    const int B = 1 << log2B, cols = 2 * B, log2cols = log2B + 1;
    const int tid = threadIdx.x;
    {
        int off = 0, len = Tn;
        if (log2T & 1) {
            for (int j = tid; j < Tn / 2; j += kFftThreads) {
                double sn, cs;
                sincospi(-2.0 * (double)j / (double)Tn, &sn, &cs);
                W[j] = {cs, sn};
            }
            off = Tn / 2;
            len = Tn / 2;
        }
        while (len >= 4) {
            const int quarter = len >> 2;
            for (int j = tid; j < 3 * quarter; j += kFftThreads) {
                const int pos = j / 3, mm = j - pos * 3 + 1;
                double sn, cs;
                sincospi(-2.0 * (double)(mm * pos) / (double)len, &sn, &cs);
                W[off + j] = {cs, sn};
            }
            off += 3 * quarter;
            len = quarter;
        }
        for (int j = tid; j < Tn; j += kFftThreads) Qb[j] = 0.0;
    }
    const T* base = samples + (long long)blockIdx.y * stride_k;
    const long long n_pairs = (n + 1) / 2;
    const long long n_tiles = (n_pairs + B - 1) >> log2B;
    const int col = tid & (cols - 1), row0 = tid >> log2cols, row_step = kFftThreads >> log2cols;
    const int nl = (Tn + row_step - 1) / row_step;            // loads per thread and tile (<= kFftPre)
    double* zcol = reinterpret_cast<double*>(Z + (size_t)(col >> 1) * ZS) + (col & 1);

    double pre[kFftPre];
    auto fetch = [&](long long tl) {
        const long long i = tl * cols + col;
        const T* src = base + i + (long long)row0 * stride_it;
#pragma unroll
        for (int k = 0; k < kFftPre; ++k) {
            const int t = row0 + k * row_step;
            pre[k] = (k < nl && t < Tn && i < n) ? (double)src[(long long)k * row_step * stride_it] : 0.0;
        }
    };
    long long tl = blockIdx.x;
    if (tl < n_tiles) fetch(tl);
    for (; tl < n_tiles; tl += gridDim.x) {
        __syncthreads();                                       // the previous tile is no longer read
#pragma unroll
        for (int k = 0; k < kFftPre; ++k) {
            const int t = row0 + k * row_step;
            if (k < nl && t < Tn) zcol[2 * swz(t)] = pre[k];
        }
        __syncthreads();
        if (tl + gridDim.x < n_tiles) fetch(tl + gridDim.x);   // in flight during the transform below
        // ---- in-place DIF
        int len = Tn, off = 0;
        if (log2T & 1) {                                       // one radix-2 stage
            const int half = len >> 1, log2h = log2T - 1;
            for (int e = tid; e < (B << log2h); e += kFftThreads) {
                const int b = e >> log2h, pos = e & (half - 1);
                cplx* z = Z + (size_t)b * ZS;
                const cplx a = z[swz(pos)], c2 = z[swz(pos + half)];
                z[swz(pos)] = cadd(a, c2);
                z[swz(pos + half)] = cmul(csub(a, c2), W[pos]);
            }
            len = half;
            off = half;
            __syncthreads();
        }
        const int log2ps = log2T - 2;                          // butterflies of one series in a radix-4 stage = T / 4
        int log2q = ((log2T & 1) ? log2T - 1 : log2T) - 2;     // log2(quarter) of the current stage
        while (len >= 4) {
            const int quarter = len >> 2;
            const cplx* tw = W + off;
            for (int e = tid; e < (B << log2ps); e += kFftThreads) {
                const int b = e >> log2ps, r = e & ((1 << log2ps) - 1);
                const int blk = r >> log2q, pos = r & (quarter - 1);
                cplx* z = Z + (size_t)b * ZS;
                const int i0 = blk * len + pos;
                const int p0 = swz(i0), p1 = swz(i0 + quarter), p2 = swz(i0 + 2 * quarter), p3 = swz(i0 + 3 * quarter);
                const cplx z0 = z[p0], z1 = z[p1], z2 = z[p2], z3 = z[p3];
                const cplx a = cadd(z0, z2), bb = csub(z0, z2), c2 = cadd(z1, z3), dd = mul_neg_i(csub(z1, z3));
                z[p0] = cadd(a, c2);
                z[p1] = cmul(cadd(bb, dd), tw[3 * pos]);
                z[p2] = cmul(csub(a, c2), tw[3 * pos + 1]);
                z[p3] = cmul(csub(bb, dd), tw[3 * pos + 2]);
            }
            off += 3 * quarter;
            len = quarter;
            log2q -= 2;
            __syncthreads();
        }
        // ---- power per position, summed over the pairs of the tile (positions are owned by threads: no atomics)
        for (int pp = tid; pp < Tn; pp += kFftThreads) {
            double s = 0.0;
            for (int b = 0; b < B; ++b) {
                const cplx v = Z[(size_t)b * ZS + pp];
                s = fma(v.re, v.re, fma(v.im, v.im, s));
            }
            Qb[swz(pp)] += s;                                  // physical slot pp holds logical position swz(pp)
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
    // B = 2^log2B series pairs per tile: as many as fit beside the twiddle table and the power sums, at most 8
    // (16 particles = one 128-byte row per time step) and at most kFftPre prefetched samples per thread
    const size_t fixed = (size_t)Tn * (2 * sizeof(cplx) + sizeof(double));
    int log2B = 3;
    while (log2B > 0 && (fixed + ((size_t)(Tn + 1) << log2B) * sizeof(cplx) > 200 * 1024 ||
                         ((long long)Tn << (log2B + 1)) > (long long)kFftPre * kFftThreads)) --log2B;
    const int B = 1 << log2B;
    const size_t smem = fixed + (size_t)B * (Tn + 1) * sizeof(cplx);
    if (smem > 227 * 1024 || ((long long)Tn << (log2B + 1)) > (long long)kFftPre * kFftThreads) return cudaErrorInvalidValue;
    const long long n_tiles = ((n + 1) / 2 + B - 1) / B;
    long long gx = 148;
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    // one launch per dim: a time row of one dim is n * sizeof(T) bytes and a tile reads 16 particles of EVERY time row, so a
    // block cycles through the T / (rows per 2 MB page) pages of its dim; with all dims in flight at once the page
    // working set (T * d rows) overran the TLB (0.3 TB/s); dim by dim it is T rows
    dim3 grid((unsigned)gx, 1u);
    e = dtype == MJHMC_F64
            ? cudaFuncSetAttribute(autocorr_fft_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(autocorr_fft_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (int k = 0; k < d; ++k) {
        if (dtype == MJHMC_F64)
            autocorr_fft_kernel<double><<<grid, kFftThreads, smem, s>>>((const double*)samples + (long long)k * stride_k, stride_k,
                                                                        stride_it, n, Tn, log2T, log2B, scratch);
        else
            autocorr_fft_kernel<float><<<grid, kFftThreads, smem, s>>>((const float*)samples + (long long)k * stride_k, stride_k,
                                                                       stride_it, n, Tn, log2T, log2B, scratch);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    autocorr_fft_finish<<<(n_lags + 255) / 256, 256, sizeof(double) * Tn, s>>>(scratch, Tn, log2T, n_lags, ac);
    return cudaGetLastError();
}

}  // namespace mjhmc
