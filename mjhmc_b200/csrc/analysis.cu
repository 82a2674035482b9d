// K6 dwell-time resampler and K7 autocorrelation.
//   resample : ContinuousTimeHMC.sample, samplers/markov_jump_hmc.py:321-328
//   autocorr : fft_autocor, misc/autocor.py:37-49 (as a direct circular product)
#include "common.cuh"
#include "analysis.h"

namespace mjhmc {

// ------------------------------------------------------------------ inclusive scan (double)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                     // per thread
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ double block_exclusive_scan(double v, double* sm, double& total) {
    // Hillis-Steele over warp totals; returns the exclusive prefix of v within the block.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = lane < (kScanThreads / 32) ? sm[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < (kScanThreads / 32)) sm[lane] = w;
    }
    __syncthreads();
    const double warp_off = warp ? sm[warp - 1] : 0.0;
    total = sm[kScanThreads / 32 - 1];
    __syncthreads();
    return warp_off + inc - v;
}

// phase 1: chunk sums
__global__ void __launch_bounds__(kScanThreads) scan_chunk_sums(const double* __restrict__ in, long long m,
                                                                double* __restrict__ sums) {
    __shared__ double sm[32];
    const long long base = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) if (base + j < m) s += in[base + j];
    double total;
    block_exclusive_scan(s, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// phase 2: in-place inclusive->exclusive scan of the chunk sums by one block
__global__ void __launch_bounds__(kScanThreads) scan_sums_inplace(double* __restrict__ sums, long long nchunks) {
    __shared__ double sm[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (long long base = 0; base < nchunks; base += kScanThreads) {
        const long long idx = base + threadIdx.x;
        const double v = idx < nchunks ? sums[idx] : 0.0;
        double total;
        const double ex = block_exclusive_scan(v, sm, total);
        const double carry = carry_s;
        if (idx < nchunks) sums[idx] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

// phase 3: scan inside each chunk + chunk offset
__global__ void __launch_bounds__(kScanThreads) scan_chunks(const double* __restrict__ in, long long m,
                                                            const double* __restrict__ offsets,
                                                            double* __restrict__ out) {
    __shared__ double sm[32];
    const long long base = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    double loc[kScanItems];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { loc[j] = base + j < m ? in[base + j] : 0.0; s += loc[j]; }
    double total;
    double run = offsets[blockIdx.x] + block_exclusive_scan(s, sm, total);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { run += loc[j]; if (base + j < m) out[base + j] = run; }
}

// idx[j] = first i with cumul[i] > r[j]  (np.searchsorted(cumul, r, 'right')), then gather columns
template <typename T>
__global__ void __launch_bounds__(256) search_gather_kernel(const double* __restrict__ cumul, long long m,
                                                            const double* __restrict__ r, long long m_out, int d,
                                                            const T* __restrict__ samples, long long ld_in,
                                                            T* __restrict__ out, long long ld_out,
                                                            long long* __restrict__ idx_out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m_out) return;
    const double rv = r[j];
    long long lo = 0, hi = m;                    // first index with cumul > rv
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (cumul[mid] > rv) hi = mid; else lo = mid + 1;
    }
    if (idx_out) idx_out[j] = lo;
    const long long src = lo < m ? lo : m - 1;   // the reference raises IndexError past the end
    for (int k = 0; k < d; ++k) out[k * ld_out + j] = samples[k * ld_in + src];
}

long long resample_scratch_bytes(long long m) {
    const long long nchunks = (m + kScanChunk - 1) / kScanChunk;
    return (long long)sizeof(double) * (m + nchunks + 8);
}

cudaError_t launch_resample(int dtype, int d, const double* dwell, long long m, const double* r, long long m_out,
                            const void* samples, long long ld_in, void* out, long long ld_out, long long* idx_out,
                            void* scratch, cudaStream_t s) {
    if (m == 0 || m_out == 0) return cudaSuccess;
    const long long nchunks = (m + kScanChunk - 1) / kScanChunk;
    double* cumul = (double*)scratch;
    double* sums = cumul + m;
    scan_chunk_sums<<<(unsigned)nchunks, kScanThreads, 0, s>>>(dwell, m, sums);
    scan_sums_inplace<<<1, kScanThreads, 0, s>>>(sums, nchunks);
    scan_chunks<<<(unsigned)nchunks, kScanThreads, 0, s>>>(dwell, m, sums, cumul);
    const unsigned grid = (unsigned)((m_out + 255) / 256);
    if (dtype == MJHMC_F64)
        search_gather_kernel<double><<<grid, 256, 0, s>>>(cumul, m, r, m_out, d, (const double*)samples, ld_in,
                                                          (double*)out, ld_out, idx_out);
    else
        search_gather_kernel<float><<<grid, 256, 0, s>>>(cumul, m, r, m_out, d, (const float*)samples, ld_in,
                                                         (float*)out, ld_out, idx_out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ autocorrelation
// Block = (one dim k, a tile of kAcP particles).  The T x kAcP series tile sits in shared
// memory; thread (p, g) accumulates lags tau = g, g+kAcG, ... for particle p; partial sums
// over the particles of the tile are combined by a shared-memory reduction and added to ac[].
constexpr int kAcP = 16;
constexpr int kAcG = 16;

template <typename T>
__global__ void __launch_bounds__(kAcP * kAcG) autocorr_kernel(const T* __restrict__ samples, long long stride_k,
                                                               long long stride_it, long long n, int Tn, int n_lags,
                                                               int circular, double* __restrict__ ac) {
    extern __shared__ double tile[];              // [Tn][kAcP]
    __shared__ double red[kAcG][kAcP + 1];
    const int p = threadIdx.x % kAcP, g = threadIdx.x / kAcP;
    const long long i0 = (long long)blockIdx.x * kAcP;
    const T* base = samples + (long long)blockIdx.y * stride_k;
    for (int t = g; t < Tn; t += kAcG) {
        const long long i = i0 + p;
        tile[t * kAcP + p] = i < n ? (double)base[(long long)t * stride_it + i] : 0.0;
    }
    __syncthreads();
    for (int tau0 = 0; tau0 < n_lags; tau0 += kAcG) {
        const int tau = tau0 + g;
        double acc = 0.0;
        if (tau < n_lags) {
            if (circular) {
                int t2 = tau % Tn;
                for (int t = 0; t < Tn; ++t) {
                    acc += tile[t * kAcP + p] * tile[t2 * kAcP + p];
                    t2 = t2 + 1 == Tn ? 0 : t2 + 1;
                }
            } else {
                // linear window: sum_{t < T - tau} x[t] x[t + tau]  (slow_autocorrelation, autocor.py:177-211)
                for (int t = 0; t + tau < Tn; ++t) acc += tile[t * kAcP + p] * tile[(t + tau) * kAcP + p];
            }
        }
        red[g][p] = acc;
        __syncthreads();
        if (p == 0 && tau < n_lags) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < kAcP; ++q) s += red[g][q];
            atomicAdd(ac + tau, s);
        }
        __syncthreads();
    }
}

// Register-blocked form: thread (p, g) owns particle p of the tile and kAbR consecutive lags, and keeps a
// sliding window of the series in registers, so 16 shared-memory loads feed 64 fp64 FMAs (the kernel above
// issues two loads per FMA and is bound by shared-memory bandwidth at 1/8 of the fp64 rate).  The series tile is
// extended past T by its own beginning (circular) or by zeros (linear) so the window needs no index arithmetic;
// a block walks several particle tiles and flushes its per-lag sums once (n_lags atomics per block, not per tile).
constexpr int kAbP = 16;                 // particles per tile
constexpr int kAbG = 16;                 // lag groups per block
constexpr int kAbR = 8;                  // lags per thread and pass
constexpr int kAbPass = kAbG * kAbR;     // lags per pass of a block

template <typename T>
__global__ void __launch_bounds__(kAbP * kAbG) autocorr_blocked_kernel(const T* __restrict__ samples, long long stride_k,
                                                                       long long stride_it, long long n, int Tn, int n_lags,
                                                                       int circular, double* __restrict__ ac, int ext) {
    extern __shared__ double smem_ac[];
    double* tile = smem_ac;                         // [ext][kAbP]
    double* lag_sum = smem_ac + (size_t)ext * kAbP; // [n_passes * kAbPass]
    const int p = threadIdx.x % kAbP, g = threadIdx.x / kAbP;
    const int n_pass = (n_lags + kAbPass - 1) / kAbPass;
    for (int q = threadIdx.x; q < n_pass * kAbPass; q += blockDim.x) lag_sum[q] = 0.0;
    const T* base = samples + (long long)blockIdx.y * stride_k;
    const long long n_tiles = (n + kAbP - 1) / kAbP;
    for (long long tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        __syncthreads();                                             // the previous tile is no longer read
        const long long i = tl * kAbP + p;
        for (int e = g; e < ext; e += kAbG) {
            double v = 0.0;
            if (i < n) {
                if (e < Tn) v = (double)base[(long long)e * stride_it + i];
                else if (circular) v = (double)base[(long long)(e % Tn) * stride_it + i];
            }
            tile[e * kAbP + p] = v;
        }
        __syncthreads();
        for (int pass = 0; pass < n_pass; ++pass) {
            const int tau0 = pass * kAbPass + g * kAbR;
            double acc[kAbR];
#pragma unroll
            for (int j = 0; j < kAbR; ++j) acc[j] = 0.0;
            double b[2 * kAbR];
#pragma unroll
            for (int j = 0; j < kAbR; ++j) b[j] = tile[(tau0 + j) * kAbP + p];
            for (int t = 0; t < Tn; t += kAbR) {
                double a[kAbR];
#pragma unroll
                for (int j = 0; j < kAbR; ++j) {
                    a[j] = t + j < Tn ? tile[(t + j) * kAbP + p] : 0.0;
                    b[kAbR + j] = tile[(t + tau0 + kAbR + j) * kAbP + p];
                }
#pragma unroll
                for (int ii = 0; ii < kAbR; ++ii)
#pragma unroll
                    for (int j = 0; j < kAbR; ++j) acc[j] = fma(a[ii], b[ii + j], acc[j]);
#pragma unroll
                for (int j = 0; j < kAbR; ++j) b[j] = b[kAbR + j];
            }
            // sum over the 16 particles of the tile (a half-warp), then one owner per lag
#pragma unroll
            for (int j = 0; j < kAbR; ++j) {
#pragma unroll
                for (int o = kAbP / 2; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            }
            if (p == 0) {
#pragma unroll
                for (int j = 0; j < kAbR; ++j) lag_sum[tau0 + j] += acc[j];
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n_lags; q += blockDim.x)
        if (lag_sum[q] != 0.0) atomicAdd(ac + q, lag_sum[q]);
}

cudaError_t launch_autocorr(int dtype, int d, const void* samples, long long stride_k, long long stride_it,
                            long long n, int Tn, int n_lags, int circular, double* ac, cudaStream_t s) {
    if (n == 0 || Tn == 0 || n_lags == 0) return cudaSuccess;
    cudaError_t e;
    // blocked kernel when the extended tile fits
    const int n_pass = (n_lags + kAbPass - 1) / kAbPass;
    const int ext = (Tn + kAbR - 1) / kAbR * kAbR + n_pass * kAbPass;
    const size_t smem_b = sizeof(double) * ((size_t)ext * kAbP + (size_t)n_pass * kAbPass);
    if (smem_b <= 200 * 1024) {
        const long long n_tiles = (n + kAbP - 1) / kAbP;
        long long gx = (148 * 4 + d - 1) / d;
        if (gx > n_tiles) gx = n_tiles;
        if (gx < 1) gx = 1;
        dim3 grid((unsigned)gx, (unsigned)d);
        if (dtype == MJHMC_F64) {
            e = cudaFuncSetAttribute(autocorr_blocked_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
            if (e != cudaSuccess) return e;
            autocorr_blocked_kernel<double><<<grid, kAbP * kAbG, smem_b, s>>>((const double*)samples, stride_k, stride_it, n,
                                                                               Tn, n_lags, circular, ac, ext);
        } else {
            e = cudaFuncSetAttribute(autocorr_blocked_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
            if (e != cudaSuccess) return e;
            autocorr_blocked_kernel<float><<<grid, kAbP * kAbG, smem_b, s>>>((const float*)samples, stride_k, stride_it, n,
                                                                              Tn, n_lags, circular, ac, ext);
        }
        return cudaGetLastError();
    }
    const size_t smem = sizeof(double) * (size_t)Tn * kAcP;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((n + kAcP - 1) / kAcP), (unsigned)d);
    if (dtype == MJHMC_F64) {
        e = cudaFuncSetAttribute(autocorr_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        autocorr_kernel<double><<<grid, kAcP * kAcG, smem, s>>>((const double*)samples, stride_k, stride_it, n, Tn, n_lags, circular, ac);
    } else {
        e = cudaFuncSetAttribute(autocorr_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        autocorr_kernel<float><<<grid, kAcP * kAcG, smem, s>>>((const float*)samples, stride_k, stride_it, n, Tn, n_lags, circular, ac);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ state-ladder walk
// experiments/spectral.py:107-131: one thread per particle replays its recorded operator choices on the dihedral state
// group of samplers/algebraic_hmc.py:485-519 and counts node visits in a (2, 2K+1) window kept in shared memory.
__global__ void __launch_bounds__(256) ladder_visits_kernel(const unsigned char* __restrict__ choice, long long n_iter,
                                                            long long n, int K, int* __restrict__ state,
                                                            unsigned long long* __restrict__ visits) {
    extern __shared__ unsigned int lv[];             // 2 (2K + 1) + 1 block-local counts
    const int W = 2 * K + 1, nb = 2 * W + 1;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) lv[j] = 0u;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int k1 = state[2 * i], k2 = state[2 * i + 1];
        for (long long it = 0; it < n_iter; ++it) {
            const unsigned c = choice[it * n + i];
            if (c == 2u) { k1 = 0; k2 = 0; }                          // R: a new ladder
            else if (c == 0u) k2 += k1 ? -1 : 1;                      // L = F (F L): moves along the ladder, keeps k1
            else if (c == 1u) k1 ^= 1;                                // F (any other code: no move on the ladder)
            const int slot = (k2 >= -K && k2 <= K) ? k1 * W + k2 + K : 2 * W;
            atomicAdd(&lv[slot], 1u);
        }
        state[2 * i] = k1; state[2 * i + 1] = k2;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nb; j += blockDim.x) if (lv[j]) atomicAdd(visits + j, (unsigned long long)lv[j]);
}

cudaError_t launch_ladder_visits(const unsigned char* choice, long long n_iter, long long n, int K, int* state,
                                 long long* visits, cudaStream_t s) {
    if (n == 0 || n_iter == 0) return cudaSuccess;
    const size_t smem = sizeof(unsigned int) * (2 * (2 * (size_t)K + 1) + 1);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    ladder_visits_kernel<<<(unsigned)((n + 255) / 256), 256, smem, s>>>(choice, n_iter, n, K, state, (unsigned long long*)visits);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ moments
// out[0] += sum x, out[1] += sum x^2 over `count` contiguous elements (online_variance of
// misc/gen_mj_init.py:76-98 in chunks: the caller merges the chunks with Chan's formula).
template <typename T>
__global__ void __launch_bounds__(256) moments_kernel(const T* __restrict__ x, long long count, double* __restrict__ out) {
    __shared__ double sm[2][8];
    double s = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        s += v;
        s2 += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[0][warp] = s; sm[1][warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sm[0][w]; b += sm[1][w]; }
        atomicAdd(out, a);
        atomicAdd(out + 1, b);
    }
}

cudaError_t launch_moments(int dtype, const void* x, long long count, double* out, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    long long blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == MJHMC_F64) moments_kernel<double><<<(unsigned)blocks, 256, 0, s>>>((const double*)x, count, out);
    else moments_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)x, count, out);
    return cudaGetLastError();
}

}  // namespace mjhmc
