// K4: fused sampler for the dense-contraction energies, fp64 tensor-core (DMMA) path.
//
//   full-covariance Gaussian  misc/distributions.py:268-273   dEdX = S X,  E = x.(S x)/2,  S = (J + J^T)/2
//   ProductOfT                misc/distributions.py:420-433   Y = W^T X + b, E = sum_j (nu_j+1)/2 log(1 + (y_j/nu_j)^2)
//                             (+ hand-derived autodiff of :431) G_j = (nu_j+1) y_j / (nu_j^2 + y_j^2), dEdX = W G
//
// Work decomposition: A WARP RUNS 8 TRAJECTORIES AT A TIME (the N of the MMA) out of the job list of its particle
// range (28-32 particles per range; the F.L.F trajectories of only the particles that need them, then the L
// trajectory of every particle -- see the main loop).  The gradient of the 8 columns is a (rows x K) . (K x 8)
// product issued as mma.sync.m8n8k4.f64 with the matrix (S, or W^T then W) as the A operand from shared memory
// (staged once per CTA, read by every warp) and the warp's own 8 positions as the B operand from a warp-private
// shared tile.  The accumulator layout of the MMA gives every lane a fixed set of (dimension, column) elements:
// the momentum of exactly those elements lives in that lane's registers across the L leapfrog steps, the gradient
// arrives in the same registers as the MMA result, and the position goes through the warp-private tile (written by
// its owner lane, read as the next B operand).  Warps never synchronise with each other inside the sampling loop
// (only __syncwarp).
//
// The particle state (X, V) stays in HBM between iterations of one launch and is re-read at the
// start of every trajectory: at L = 10..25 the kernel does 2 d^2 L flops per 5 d S bytes
// (50..125 flop/B in fp64) -- far above the fp64 machine balance, so the bound is the fp64
// tensor pipe, not HBM.
#include "dense.h"

namespace mjhmc {

constexpr int kDenseMaxWarps = 8;
constexpr int kDenseThreads = kDenseMaxWarps * 32;   // upper bound; the launcher picks the warp count that fits smem
constexpr int kCols = 8;                         // particles per warp (N of the MMA)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Row stride (doubles) of the staged matrices: 8*MT + 4 == 4 (mod 8), so the 8x4 A fragment of
// mma.m8n8k4 (rows at stride LD, 4 consecutive doubles each) and the transposed 4x8 read of the
// same matrix both hit every bank exactly twice -- the minimum for a 256-byte warp request.  A
// compile-time stride turns the per-tile offsets into LDS immediates.
template <int MT> struct DenseLd { static constexpr int value = 8 * MT + 4; };

// sum over the 8 lanes that share lane%4 (i.e. over the rows of an accumulator column pair)
__device__ __forceinline__ double col_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// value of particle column `c` (0..7) from the per-lane column-pair sums {s0, s1}
__device__ __forceinline__ double col_pick(double s0, double s1, int c) {
    const double a = __shfl_sync(0xffffffffu, s0, c >> 1);
    const double b = __shfl_sync(0xffffffffu, s1, c >> 1);
    return (c & 1) ? b : a;
}

struct DenseShared {
    double* A1;     // [8*MT][LD]  Gaussian: S.  ProductOfT: W (rows = dims, cols = experts), read both ways
    double* nu;     // ProductOfT [8*MT]
    double* bias;   // ProductOfT [8*MT]
    double* Xw;     // warp-private [kpad][8]
    double* Yw;     // ProductOfT warp-private [kpad][8]
};

// acc[mt] = M[rows 8mt.., :] . B  (TRANS = false)  or  M^T[rows 8mt.., :] . B  (TRANS = true)
// for this warp's 8 columns; M is [8*MT][LD] in shared memory, the B tile is [kpad][8].
template <int MT, bool TRANS>
__device__ __forceinline__ void warp_gemm(const double* __restrict__ M, const double* __restrict__ B,
                                          int ksteps, double (&acc)[MT][2]) {
    constexpr int LD = DenseLd<MT>::value;
    const int lane = threadIdx.x & 31;
    const int ar = lane >> 2, ac = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { acc[mt][0] = 0.0; acc[mt][1] = 0.0; }
    // A fragment element (row 8mt + ar, col 4kk + ac) of M, resp. of M^T = M[4kk + ac][8mt + ar]
    const double* a_ptr = TRANS ? M + ac * LD + ar : M + ar * LD + ac;
    const double* b_ptr = B + ac * kCols + ar;
#pragma unroll 1
    for (int kk = 0; kk < ksteps; ++kk) {
        const double b = *b_ptr;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(acc[mt], TRANS ? a_ptr[mt * 8] : a_ptr[mt * 8 * LD], b);
        a_ptr += TRANS ? 4 * LD : 4;
        b_ptr += 4 * kCols;
    }
}

template <int MT, bool POT>
__global__ void __launch_bounds__(kDenseThreads, 1)
dense_sample_kernel(const __grid_constant__ LaunchParams p) {
    extern __shared__ double smem[];
    const int d = p.d;
    const int rows = 8 * MT;                       // padded dims (== padded experts for ProductOfT)
    const int kpad = (d + 3) & ~3;
    const int ksteps = kpad >> 2;
    constexpr int ld = DenseLd<MT>::value;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ar = lane >> 2, q = lane & 3;        // accumulator row within a tile, column pair

    DenseShared sh;
    sh.A1 = smem;
    double* cur = sh.A1 + rows * ld;
    sh.nu = nullptr; sh.bias = nullptr; sh.Yw = nullptr;
    if (POT) {
        sh.nu = cur; cur += rows;
        sh.bias = cur; cur += rows;
    }
    sh.Xw = cur + warp * (POT ? 2 : 1) * rows * kCols;
    if (POT) sh.Yw = sh.Xw + rows * kCols;

    // ---- stage the matrix once per CTA (zero padded): one TMA bulk copy per matrix row (cp.async.bulk, 8 d bytes,
    // global row r -> shared row r at stride ld) on one mbarrier, issued by the lanes of warp 0; the threads zero the
    // padding the copies do not touch.  Rows that are not 16-byte multiples (odd ndims) are staged with plain loads.
    __shared__ __align__(8) uint64_t bar_stage;
    {
        const double* Mg = (const double*)p.a0;     // Gaussian: S [d][d].  ProductOfT: W [d][nb], nb == d
        const bool bulk = (d & 1) == 0 && ((uintptr_t)Mg & 15) == 0;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_stage);
        if (bulk) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(d * d * 8)) : "memory");
            }
            __syncthreads();
            if (warp == 0) {
                for (int r = lane; r < d; r += 32)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(sh.A1 + r * ld)), "l"(Mg + (size_t)r * d),
                                   "r"((uint32_t)(d * 8)), "r"(bar) : "memory");
            }
        }
        for (int idx = threadIdx.x; idx < rows * ld; idx += blockDim.x) {
            const int r = idx / ld, c = idx - r * ld;
            const bool data = r < d && c < d;
            if (!data) sh.A1[idx] = 0.0;
            else if (!bulk) sh.A1[idx] = Mg[r * d + c];
        }
        if (POT) {
            const double* nu = (const double*)p.a1;
            const double* b = (const double*)p.a2;
            for (int idx = threadIdx.x; idx < rows; idx += blockDim.x) {
                sh.nu[idx] = idx < d ? nu[idx] : 1.0;
                sh.bias[idx] = idx < d ? b[idx] : 0.0;
            }
        }
        if (bulk) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "DSTAGE_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                "@p bra DSTAGE_DONE;\n"
                "bra DSTAGE_WAIT;\n"
                "DSTAGE_DONE:\n"
                "}\n" ::"r"(bar) : "memory");
        }
    }
    for (int idx = lane; idx < (POT ? 2 : 1) * rows * kCols; idx += 32) sh.Xw[idx] = 0.0;
    __syncthreads();

    const double eps = p.eps, nhe = -p.eps / 2.0;
    const int L = p.L, sampler = p.sampler;
    unsigned int n_l = 0, n_f = 0, n_fl = 0, n_r = 0, n_E = 0, n_exec = 0;

    unsigned long long* work_head = p.counters + (size_t)MJHMC_COUNTER_STRIPES * MJHMC_N_COUNTERS;
    double v[MT][2], g[MT][2];

    // Gradient of the positions in Xw -> g, and (when want_e) their energy as per-lane column-pair
    // sums e0, e1.  One call site only: the unrolled MMA sequences are the bulk of the code and the
    // eight warps of a persistent CTA run out of phase, so every extra copy costs instruction-cache
    // misses (profiles/r1_dense_pot_v3: 46 % stall_no_inst with five inlined copies).
    //   Gaussian   : g = S x,  E = x.g / 2
    //   ProductOfT : y = W^T x + b;  E = sum_j (nu_j+1)/2 log(1 + (y_j/nu_j)^2);  g = W [(nu+1) y / (nu^2 + y^2)]
    auto gradient = [&](bool want_e, double& e0, double& e1) {
        e0 = 0.0; e1 = 0.0;
        if (POT) {
            double y[MT][2];
            warp_gemm<MT, true>(sh.A1, sh.Xw, ksteps, y);            // Y = W^T X
            if (want_e) {
                // energy through a rolled loop over the tile rows (the values take a detour through the
                // warp's Yw tile) so the two log() expansions exist once, not 2*MT times
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    *reinterpret_cast<double2*>(sh.Yw + (mt * 8 + ar) * kCols + 2 * q) = make_double2(y[mt][0], y[mt][1]);
#pragma unroll 1
                for (int mt = 0; mt < MT; ++mt) {
                    const int j = mt * 8 + ar;
                    if (j < d) {
                        const double2 yy = *reinterpret_cast<const double2*>(sh.Yw + j * kCols + 2 * q);
                        const double nu = sh.nu[j], b = sh.bias[j], alpha = (nu + 1.0) * 0.5;
                        const double r0 = (yy.x + b) / nu, r1 = (yy.y + b) / nu;
                        e0 += alpha * log(1.0 + r0 * r0);
                        e1 += alpha * log(1.0 + r1 * r1);
                    }
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int j = mt * 8 + ar;
                const double nu = sh.nu[j], b = sh.bias[j];
                const double y0 = y[mt][0] + b, y1 = y[mt][1] + b;
                // (nu+1) y / (nu^2 + y^2) with a correctly rounded reciprocal (MUFU.RCP64H + Newton) instead of
                // the full division sequence: <= 1 ulp apart, a third of the code
                double2 o;
                o.x = j < d ? (nu + 1.0) * y0 * __drcp_rn(fma(y0, y0, nu * nu)) : 0.0;
                o.y = j < d ? (nu + 1.0) * y1 * __drcp_rn(fma(y1, y1, nu * nu)) : 0.0;
                *reinterpret_cast<double2*>(sh.Yw + j * kCols + 2 * q) = o;
            }
            __syncwarp();
            warp_gemm<MT, false>(sh.A1, sh.Yw, ksteps, g);          // dEdX = W G
        } else {
            warp_gemm<MT, false>(sh.A1, sh.Xw, ksteps, g);
            if (want_e) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const double2 x = *reinterpret_cast<const double2*>(sh.Xw + (mt * 8 + ar) * kCols + 2 * q);
                    e0 += x.x * g[mt][0];
                    e1 += x.y * g[mt][1];
                }
                e0 *= 0.5; e1 *= 0.5;
            }
        }
        __syncwarp();
        if (want_e) { e0 = col_sum(e0); e1 = col_sum(e1); }
    };

    auto kinetic2 = [&](double& k0, double& k1) {
        k0 = 0.0; k1 = 0.0;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { k0 += v[mt][0] * v[mt][0]; k1 += v[mt][1] * v[mt][1]; }
        k0 = col_sum(k0) * 0.5; k1 = col_sum(k1) * 0.5;
    };

    // Persistent warps own RANGES of up to 32 consecutive particles (lane j <-> particle start + j), handed out
    // through an atomic work head (row MJHMC_COUNTER_STRIPES of the counter block, counted in particles), and the
    // matrix is staged once per SM for the whole launch.  COLUMNS ARE JOBS: per iteration the range is a list of
    // trajectories -- first the F.L.F trajectories of only those particles whose cached FLF energy is not valid
    // (hmc_state.py:109-119), then the L trajectory of every particle -- run eight at a time (the N of the MMA).
    // Until round 2 a warp owned 8 particles and ran a second pass over all 8 whenever one of them needed the FLF
    // energy: 1.34 passes per 8 particles for ProductOfT-100d at the searched hyper-parameters where 1.05 are needed.
    // Columns of an MMA are independent, so the packing cannot change a result.
    const bool mjs = sampler == MJHMC_SAMPLER_MARKOV_JUMP;
    const long long n_warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    unsigned int seen_parts = 0, seen_flf = 0;     // this warp's running FLF fraction picks the next range size

    for (;;) {
        // ---- range size: the expected number of passes per particle, E ceil((P + Binomial(P, f)) / 8) / P, is
        // smallest at these P (tabulated offline for f = FLF fraction); the last ranges of a launch are short so the
        // warps finish together
        int take = 32;
#ifndef MJ_DENSE_FIXED_RANGE
        if (mjs && seen_parts >= 64u) {
            const float f = (float)seen_flf / (float)seen_parts;
            take = f < 0.01f ? 32 : f < 0.035f ? 31 : f < 0.065f ? 30 : f < 0.10f ? 29 : f < 0.16f ? 28
                 : f < 0.25f ? 32 : f < 0.40f ? 30 : f < 0.90f ? 31 : 32;
        }
#else
        take = MJ_DENSE_FIXED_RANGE;
#endif
        long long start = 0;
        if (lane == 0) {
            const long long seen_head = (long long)*(volatile unsigned long long*)work_head;
            if (p.n - seen_head < n_warps_total * 24) take = 8;
            start = (long long)atomicAdd(work_head, (unsigned long long)take);
        }
        start = __shfl_sync(0xffffffffu, start, 0);
        take = __shfl_sync(0xffffffffu, take, 0);
        if (start >= p.n) break;
        const int np = (int)min((long long)take, p.n - start);
        // per-particle bookkeeping lives in lane j < np (particle start + j)
        const long long ip = start + lane;
        const bool plive = lane < np;
        unsigned int cflags = 0;
        double Hc = 0.0, dwell = 0.0;
        bool failed = false;
        if (plive && mjs) { cflags = p.ca_in[ip]; Hc = ((const double*)p.Hc_in)[ip]; }

        for (int it = 0; it < p.n_iter; ++it) {
            const unsigned long long attempt = p.attempt0 + (unsigned long long)it;
            const double* Xc = (const double*)(it == 0 ? p.Xin : p.Xout);
            const double* Vc = (const double*)(it == 0 ? p.Vin : p.Vout);
            double* Xo = (double*)p.Xout;
            double* Vo = (double*)p.Vout;
            const bool active = plive && !failed;

            // ---- the job list of this iteration: slots [0, nF) = FLF jobs in particle order, [nF, nF + np) = L jobs
            const bool need = mjs && active && !(cflags & 2u);
            const unsigned int needmask = __ballot_sync(0xffffffffu, need);
            const int nF = __popc(needmask);
            const int njobs = nF + np;
            const int sF = __popc(needmask & ((1u << lane) - 1u));          // my FLF slot (if need)
            const int sL = nF + lane;                                       // my L slot
            if (active) { n_E += 1; n_exec += 1; }
            if (mjs && active && !(cflags & 1u)) n_E += 1;                  // the reference evaluates the FLF state here
            if (need) n_exec += 1;
            seen_parts += (unsigned int)np; seen_flf += (unsigned int)nF;

            unsigned int coin = 0;
            if (sampler == MJHMC_SAMPLER_DISCRETE) {
                if (lane == 0) coin = draw_coin(p, attempt) < p.p_r;
                coin = __shfl_sync(0xffffffffu, coin, 0);
            }

            double Hflf = Hc, H = 0.0, Hl = 0.0;
            for (int j0 = 0; j0 < njobs; j0 += kCols) {
                // ---- this lane's column pair: job slot -> (particle, sign)
                long long pi[2];
                int pl[2];                                                  // lane of the particle; -1: empty column
                double sg[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int slot = j0 + 2 * q + e;
                    pl[e] = slot >= njobs ? -1 : (slot < nF ? (int)__fns(needmask, 0u, slot + 1) : slot - nF);
                    sg[e] = slot < nF ? -1.0 : 1.0;
                    pi[e] = start + (pl[e] < 0 ? 0 : pl[e]);
                }
                // ---- load (x, sign * v): every load is issued before the first value is used (clamped row, the
                // first particle of the range for an empty column) -- with a branch per row the loads of one row waited
                // for the shared-memory store of the row before: 13 DRAM round trips per pass
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const long long ro = (long long)min(mt * 8 + ar, d - 1) * p.ld;
                    g[mt][0] = Xc[ro + pi[0]]; g[mt][1] = Xc[ro + pi[1]];
                    v[mt][0] = Vc[ro + pi[0]]; v[mt][1] = Vc[ro + pi[1]];
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int r = mt * 8 + ar;
                    const bool in0 = r < d && pl[0] >= 0, in1 = r < d && pl[1] >= 0;
                    v[mt][0] = in0 ? sg[0] * v[mt][0] : 0.0;
                    v[mt][1] = in1 ? sg[1] * v[mt][1] : 0.0;
                    *reinterpret_cast<double2*>(sh.Xw + r * kCols + 2 * q) = make_double2(in0 ? g[mt][0] : 0.0, in1 ? g[mt][1] : 0.0);
                }
                __syncwarp();

                // ---- the trajectory of the 8 columns (hmc_state.py:93-100); st = 0: only dEdX (and E) of the start point
                double h_start = 0.0, h_end = 0.0;                          // every lane holds column lane & 7
                double k0, k1, e0, e1;
                for (int st = 0; st <= L; ++st) {
                    if (st > 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            v[mt][0] += nhe * g[mt][0];
                            v[mt][1] += nhe * g[mt][1];
                            double2* xp = reinterpret_cast<double2*>(sh.Xw + (mt * 8 + ar) * kCols + 2 * q);
                            double2 x = *xp;
                            x.x += eps * v[mt][0];
                            x.y += eps * v[mt][1];
                            *xp = x;
                        }
                        __syncwarp();
                    }
                    const bool want_e = st == 0 || st == L;
                    gradient(want_e, e0, e1);
                    if (st > 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) { v[mt][0] += nhe * g[mt][0]; v[mt][1] += nhe * g[mt][1]; }
                    }
                    if (want_e) {
                        kinetic2(k0, k1);
                        const double h = col_pick(e0 + k0, e1 + k1, lane & 7);      // EX + EV, hmc_state.py:80-84
                        if (st == 0) h_start = h;
                        if (st == L) h_end = h;
                    }
                }

                // ---- the energies go to the lanes of their particles
                {
                    const double hf = __shfl_sync(0xffffffffu, h_end, sF & 7);
                    if (need && sF >= j0 && sF < j0 + kCols) Hflf = hf;
                    const double hs = __shfl_sync(0xffffffffu, h_start, sL & 7);
                    const double hl = __shfl_sync(0xffffffffu, h_end, sL & 7);
                    if (sL >= j0 && sL < j0 + kCols) { H = hs; Hl = hl; }
                }
                const bool mine = plive && sL >= j0 && sL < j0 + kCols;     // my L job ran in this pass

                // ---- decision per particle whose L job just finished (its FLF job ran in this pass or an earlier one)
                // take: 0 keep, 1 proposal, 2 proposal with flipped momentum; flip / refresh of the resulting momentum
                unsigned int tk_ = 0, flip = 0, refresh = 0, choice = 0;
                if (mine && active) {
                    if (mjs) {
                        const Decision dc = decide_mj(p, ip, attempt, H - Hl, H - Hflf, p.dwell != nullptr || (it + 1 == p.n_iter && p.dwell_last != nullptr));
                        if (dc.fail) { report_failure(p, it); failed = true; }
                        else {
                            choice = dc.choice; dwell = dc.dwell;
                            if (choice == 0) { tk_ = 1; Hc = H; cflags = 3u; n_l += 1; }
                            else if (choice == 1) { flip = 1; Hc = Hl; cflags = 2u; n_f += 1; }
                            else { refresh = 1; cflags = 0u; n_r += 1; }
                        }
                    } else if (sampler == MJHMC_SAMPLER_CONTINUOUS_TIME) {
                        const Decision dc = decide_ct(p, ip, attempt, H - Hl, p.dwell != nullptr || (it + 1 == p.n_iter && p.dwell_last != nullptr));
                        if (dc.fail) { report_failure(p, it); failed = true; }
                        else {
                            choice = dc.choice; dwell = dc.dwell;
                            if (choice == 1) { tk_ = 2; n_fl += 1; }
                            else if (choice == 0) { flip = 1; n_f += 1; }
                            else { refresh = 1; n_r += 1; }
                        }
                    } else {
                        const Decision dc = decide_discrete(p, ip, attempt, H - Hl, coin != 0);
                        choice = dc.choice;
                        const bool acc = choice & 1u, fl = choice & 2u;
                        if (acc) tk_ = 2;
                        flip = fl; refresh = (choice & 4u) ? 1u : 0u;
                        n_l += (acc && fl); n_f += (fl && !acc); n_fl += (acc && !fl); n_r += refresh;
                    }
                }
                const unsigned int ok = (mine && active && !failed) ? 1u : 0u;
                const unsigned int code = tk_ | (flip << 2) | (refresh << 3) | (ok << 4);

                // ---- apply to the L-job columns of this lane's pair.  A particle that did not take the trajectory keeps
                // its state: where a column of the warp needs it, the old state of the column is fetched in one batch of
                // independent loads into the (dead) gradient registers (load -> store pairs row by row serialise on the
                // memory latency: Xout may alias Xin)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const unsigned int cd = __shfl_sync(0xffffffffu, code, pl[e] < 0 ? 0 : pl[e]);
                    const bool lcol = pl[e] >= 0 && sg[e] > 0.0;           // not an empty column / an FLF job (only its energy is read)
                    const long long i = pi[e];
                    const unsigned int tk = cd & 3u, fp = (cd >> 2) & 1u, rf = (cd >> 3) & 1u, okc = (cd >> 4) & 1u;
                    const bool took = okc && tk;
                    if (lcol && !took) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const long long o = (long long)min(mt * 8 + ar, d - 1) * p.ld + i;
                            g[mt][0] = Xc[o];
                            g[mt][1] = Vc[o];
                        }
                    }
                    if (!lcol) continue;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int r = mt * 8 + ar;
                        if (r >= d) continue;
                        const long long o = (long long)r * p.ld + i;
                        double xn, vn;
                        if (took) {
                            xn = sh.Xw[r * kCols + 2 * q + e];
                            vn = tk == 1 ? v[mt][e] : -v[mt][e];
                        } else {
                            xn = g[mt][0];
                            vn = g[mt][1];
                        }
                        if (okc && fp) vn = -vn;
                        if (okc && rf) {
                            double z0, z1;
                            normal_pair(p, i, attempt, r >> 1, d, z0, z1);
                            vn = vn * p.r_keep + ((r & 1) ? z1 : z0) * p.r_mix;        // hmc_state.py:126
                        }
                        Xo[o] = xn;
                        Vo[o] = vn;
                        if (okc && p.samples) ((double*)p.samples)[(long long)r * p.s_stride_k + (long long)it * p.s_stride_it + i] = xn;
                    }
                }
                if (ok) {
                    if (p.dwell) p.dwell[(long long)it * p.n + ip] = dwell;
                    if (p.choice) p.choice[(long long)it * p.n + ip] = (uint8_t)choice;
                }
                __syncwarp();
            }
        }

        if (plive) {
            if (mjs) { p.ca_out[ip] = (uint8_t)cflags; ((double*)p.Hc_out)[ip] = Hc; }
            if (p.dwell_last && sampler != MJHMC_SAMPLER_DISCRETE) p.dwell_last[ip] = dwell;
        }
        if (p.n_iter == 0 && plive) {
            // nothing ran: the state still has to reach the output buffers
            for (int r = 0; r < d; ++r) {
                ((double*)p.Xout)[(long long)r * p.ld + ip] = ((const double*)p.Xin)[(long long)r * p.ld + ip];
                ((double*)p.Vout)[(long long)r * p.ld + ip] = ((const double*)p.Vin)[(long long)r * p.ld + ip];
            }
        }
    }

    // per-warp flush (no CTA barrier: warps retire independently)
    const unsigned int loc[6] = {n_l, n_f, n_fl, n_r, n_E, n_exec};
    const int slot[6] = {MJHMC_CNT_L, MJHMC_CNT_F, MJHMC_CNT_FL, MJHMC_CNT_R, MJHMC_CNT_E, MJHMC_CNT_EXEC};
    unsigned long long* row = p.counters + (size_t)((blockIdx.x * 8 + warp) % MJHMC_COUNTER_STRIPES) * MJHMC_N_COUNTERS;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const unsigned long long sum = __reduce_add_sync(0xffffffffu, loc[c]);
        if (lane == 0 && sum) {
            atomicAdd(row + slot[c], c == 5 ? sum * (unsigned long long)L : sum);
            if (c == 4) atomicAdd(row + MJHMC_CNT_DEDX, sum * (unsigned long long)L);
        }
    }
}

static size_t dense_smem_bytes(int MT, int d, bool pot, int nwarps) {
    const int rows = 8 * MT, ld = 8 * MT + 4;
    (void)d;
    size_t doubles = (size_t)rows * ld + (pot ? 2 * rows : 0) +
                     (size_t)nwarps * (pot ? 2 : 1) * rows * kCols;
    return doubles * sizeof(double);
}

// most warps per CTA (one CTA per SM) whose tiles fit next to the staged matrices
static int dense_warps(int MT, int d, bool pot) {
    for (int nw = kDenseMaxWarps; nw >= 2; nw -= 2)
        if (dense_smem_bytes(MT, d, pot, nw) <= 227 * 1024) return nw;
    return 0;
}

static int dense_mt(int d) {
    if (d <= 32) return 4;
    if (d <= 56) return 7;
    if (d <= 104) return 13;
    return 0;
}

bool dense_supported(int dtype, int kind, int ndims, int nbasis) {
    if (dtype == MJHMC_F32) return dense_tc_supported(kind, ndims, nbasis);
    if (kind == MJHMC_DIST_PRODUCT_OF_T && nbasis != ndims) return false;
    const int mt = dense_mt(ndims);
    if (!mt) return false;
    return dense_warps(mt, ndims, kind == MJHMC_DIST_PRODUCT_OF_T) > 0;
}

template <int MT, bool POT>
static cudaError_t launch_dense_T(const LaunchParams& p, cudaStream_t stream) {
    const int nwarps = dense_warps(MT, p.d, POT);
    if (!nwarps) return cudaErrorNotSupported;
    const size_t smem = dense_smem_bytes(MT, p.d, POT, nwarps);
    cudaError_t e = cudaFuncSetAttribute(dense_sample_kernel<MT, POT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long groups = (p.n + kCols - 1) / kCols;
    long long blocks = (groups + nwarps - 1) / nwarps;
    if (blocks > sms) blocks = sms;                            // one persistent CTA per SM
    dense_sample_kernel<MT, POT><<<(unsigned)blocks, nwarps * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_dense(int dtype, int kind, const LaunchParams& p, cudaStream_t stream) {
    if (dtype == MJHMC_F32) return dense_tc_supported(kind, p.d, p.nbasis) ? launch_dense_tc(kind, p, stream) : cudaErrorNotSupported;
    const bool pot = kind == MJHMC_DIST_PRODUCT_OF_T;
    switch (dense_mt(p.d)) {
        case 4:  return pot ? launch_dense_T<4, true>(p, stream) : launch_dense_T<4, false>(p, stream);
        case 7:  return pot ? launch_dense_T<7, true>(p, stream) : launch_dense_T<7, false>(p, stream);
        case 13: return pot ? launch_dense_T<13, true>(p, stream) : launch_dense_T<13, false>(p, stream);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace mjhmc
