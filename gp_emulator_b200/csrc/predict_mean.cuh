// FP64 GP prediction without the variance: mean + input gradient (+ optional Hessian) in one pass.
//
//   mu_n    = sum_j k_nj alpha_j                                   reference GaussianProcess.py:237
//   deriv_n = w_d sum_j k_nj alpha_j (x_jd - t_nd)                 reference GaussianProcess.py:244-247
//   hess_n  = sum_j k_nj alpha_j [w_d (x_jd - t_nd) w_e (x_je - t_ne) - delta_de w_d]
//                                                                  reference GaussianProcess.py:345-366
// Used for predict(do_unc=False), for GaussianProcess.hessian and for models uploaded without invQ.
// K* lives only in registers.  Thread (n, g): 8 points x 4 training-point lanes per warp, 4 warps per CTA
// (32 points); the 4 lanes of a point are combined with two shuffle steps.  Outputs are staged through shared
// memory so the (N, D) gradient and (N, D, D) Hessian rows are written fully coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_math.cuh"

namespace gpe {

constexpr int kMeanThreads = 128;
constexpr int kMeanTN = 32;

// One entry per emulator of a bank (same training inputs, own hyper-parameters): blockIdx.y selects it.
struct MeanBankEntry {
    const double* xchunks;
    double sqrt_w[32];
};

struct MeanParams {
    const double* testing;  // (N, D)
    int64_t N;
    double* mu;
    double* deriv;
    double* hess;
    int64_t ld_mu, ld_deriv, ld_hess;  // element strides between consecutive points (1, D, D*D for a single GP)
    const double* xchunks;  // [nchunks][ JC*DP xs | JC b*alpha ]
    int M, D, JC, nchunks;
    uint32_t off_xc, off_ts, off_out;  // smem byte offsets
    double sqrt_w[32];
    const MeanBankEntry* bank;          // null: single GP.  Else gridDim.y emulators, outputs offset by e * eo_*
    int64_t eo_mu, eo_deriv, eo_hess;   // element offsets per emulator into the point-major bank outputs
    double* kstar;                      // k_predict_mean2<DP, true>: K* scratch [ceil(N/16)][kblk][16][4] (predict_var_large.cuh)
    int kblk;
    uint32_t smem_need;                 // extent of the carve-up above, recomputed by the host from the same offsets and
                                        // checked against the launch's dynamic shared memory in launch_mean (gpemu.cu).  Not
                                        // at kernel entry like the other kernels: k_predict_mean2 sits at its register cap
                                        // (3 CTAs/SM) and even a one-compare guard + trap made ptxas spill (-17 %)
};

template <int DP, bool HESS>
__global__ void __launch_bounds__(kMeanThreads) k_predict_mean(const MeanParams p) {
    constexpr int TN = kMeanTN;
    constexpr int NTRI = HESS ? DP * (DP + 1) / 2 : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    double* Xc = reinterpret_cast<double*>(smem + p.off_xc);
    double* ts_s = reinterpret_cast<double*>(smem + p.off_ts);   // [TN][D]; reused as [TN][D+1] outs
    double* out_s = reinterpret_cast<double*>(smem + p.off_out); // HESS: [TN][D*D]
    __shared__ double sqw_s[32];
    __shared__ double exp_tab[64];   // 2^(j/64) for exp_neg_tab; visible after the first CTA barrier

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    exp_tab_load(exp_tab, tid);
    const int g_low = lane & 3, n_loc = warp * 8 + (lane >> 2);
    const int D = p.D, M = p.M, DV = D + 1;
    const int em = blockIdx.y;
    const double* xchunks = p.bank ? p.bank[em].xchunks : p.xchunks;
    if (tid < 32) sqw_s[tid] = p.bank ? p.bank[em].sqrt_w[tid] : p.sqrt_w[tid];
    double* const o_mu = p.mu ? p.mu + em * p.eo_mu : nullptr;
    double* const o_deriv = p.deriv ? p.deriv + em * p.eo_deriv : nullptr;
    double* const o_hess = p.hess ? p.hess + em * p.eo_hess : nullptr;

    const int64_t ntiles = (p.N + TN - 1) / TN;
    bool x_resident = false;
    constexpr int XP = x_pitch(DP);   // row pitch of the training chunk (conflict-free LDS.128)
    const int chunk_doubles = p.JC * (XP + 1);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * TN;
        const int npts = (int)min((int64_t)TN, p.N - n0);
        __syncthreads();  // previous tile's staged outputs fully drained
        for (int e = tid; e < TN * D; e += kMeanThreads) {
            const int r = e / D;
            const int64_t src = (r < npts) ? (n0 * D + e) : ((p.N - 1) * D + (e - r * D));
            ts_s[e] = __ldg(p.testing + src);
        }
        __syncthreads();
        double ts[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) ts[d] = (d < D) ? ts_s[n_loc * D + d] * sqw_s[d] : 0.0;

        double mu = 0.0;
        double g[DP];
        double T[NTRI];
#pragma unroll
        for (int d = 0; d < DP; ++d) g[d] = 0.0;
#pragma unroll
        for (int i = 0; i < NTRI; ++i) T[i] = 0.0;

        for (int c = 0; c < p.nchunks; ++c) {
            if (!x_resident) {
                __syncthreads();
                const double2* src = reinterpret_cast<const double2*>(xchunks + (size_t)c * chunk_doubles);
                double2* dst = reinterpret_cast<double2*>(Xc);
                for (int e = tid; e < chunk_doubles / 2; e += kMeanThreads) dst[e] = __ldg(src + e);
                __syncthreads();
                if (p.nchunks == 1) x_resident = true;
            }
            const int jn = min(p.JC, M - c * p.JC);
            const double* al = Xc + p.JC * XP;
            // Software pipeline over this lane's training points: the distance + exp chain of point i + 1 (a long
            // dependent FP64 sequence) is issued in the same iteration as the independent accumulate FMAs of point i
            // (D for the gradient, D (D + 1) / 2 for the Hessian), so the FP64 pipe always has ready work.
            auto stage1 = [&](int jl, double (&u)[DP]) -> double {
                const double2* x1 = reinterpret_cast<const double2*>(Xc + jl * XP);
                double r2 = 0.0;
#pragma unroll
                for (int d = 0; d < DP; d += 2) {
                    const double2 a = x1[d >> 1];
                    u[d] = a.x - ts[d];
                    u[d + 1] = a.y - ts[d + 1];
                    r2 = fma(u[d], u[d], r2);
                    r2 = fma(u[d + 1], u[d + 1], r2);
                }
                return exp_neg_tab(-0.5 * r2, exp_tab) * al[jl];
            };
            if (!HESS) {
                // mean + gradient only: few registers -> many resident warps hide the chain; no pipelining needed
                for (int jl = g_low; jl < jn; jl += 4) {
                    double u[DP];
                    const double cj = stage1(jl, u);
                    mu += cj;
#pragma unroll
                    for (int d = 0; d < DP; ++d) g[d] = fma(cj, u[d], g[d]);
                }
            } else if (g_low < jn) {
                double uc[DP], un[DP];
                double cc = stage1(g_low, uc);
                for (int jl = g_low; jl < jn; jl += 4) {
                    const double cn = stage1(min(jl + 4, jn - 1), un);   // next point (clamped; unused after the last)
                    mu += cc;
                    {
                        int t = 0;
#pragma unroll
                        for (int d = 0; d < DP; ++d) {
                            const double cu = cc * uc[d];
                            g[d] += cu;
#pragma unroll
                            for (int e = d; e < DP; ++e) {
                                T[t] = fma(cu, uc[e], T[t]);
                                ++t;
                            }
                        }
                    }
                    cc = cn;
#pragma unroll
                    for (int d = 0; d < DP; ++d) uc[d] = un[d];
                }
            }
        }

        mu += __shfl_xor_sync(0xffffffffu, mu, 1);
        mu += __shfl_xor_sync(0xffffffffu, mu, 2);
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            g[d] += __shfl_xor_sync(0xffffffffu, g[d], 1);
            g[d] += __shfl_xor_sync(0xffffffffu, g[d], 2);
        }
        __syncthreads();  // everyone has read ts_s
        double* outs = ts_s;
        if (g_low == 0) {
            outs[n_loc * DV] = mu;
#pragma unroll
            for (int d = 0; d < DP; ++d)
                if (d < D) outs[n_loc * DV + 1 + d] = g[d];
        }
        if (HESS) {
            int t = 0;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
#pragma unroll
                for (int e = d; e < DP; ++e) {
                    double v = T[t];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    ++t;
                    if (g_low == 0 && e < D) {   // e < D implies d < D
                        double h = sqw_s[d] * sqw_s[e] * v;
                        if (d == e) h -= sqw_s[d] * sqw_s[d] * mu;
                        out_s[n_loc * D * D + d * D + e] = h;
                        out_s[n_loc * D * D + e * D + d] = h;
                    }
                }
            }
        }
        __syncthreads();
        if (o_mu != nullptr && tid < npts) o_mu[(n0 + tid) * p.ld_mu] = outs[tid * DV];
        if (o_deriv != nullptr) {
            for (int e = tid; e < npts * D; e += kMeanThreads) {
                const int r = e / D, d = e - r * D;
                o_deriv[(n0 + r) * p.ld_deriv + d] = sqw_s[d] * outs[r * DV + 1 + d];
            }
        }
        if (HESS && o_hess != nullptr) {
            const int DD = D * D;
            for (int e = tid; e < npts * DD; e += kMeanThreads) {
                const int r = e / DD;
                o_hess[(n0 + r) * p.ld_hess + (e - r * DD)] = out_s[e];
            }
        }
    }
}

// Mean + gradient only, with the phase-A design of the fused kernel: a warp owns 8 test rows, 8 lanes sweep the
// training points, every thread carries TWO test rows (each training row pulled from shared memory feeds two
// pairs), training-row loads are software-pipelined one step ahead, and the 8 lanes are combined by a shuffle
// reduce-scatter.  32 points per 128-thread CTA.
// KSTAR: also store K* (without alpha) into the scratch consumed by k_var_large (M > 1024 variance path).
template <int DP, bool KSTAR>
__global__ void __launch_bounds__(kMeanThreads, (DP <= 12 ? 3 : 1)) k_predict_mean2(const MeanParams p) {
    constexpr int TN = kMeanTN;
    constexpr int NV = DP + 1;
    extern __shared__ __align__(128) unsigned char smem_m2[];
    double* Xc = reinterpret_cast<double*>(smem_m2 + p.off_xc);
    double* ts_s = reinterpret_cast<double*>(smem_m2 + p.off_ts);   // [TN][D]; reused as [TN][D+1] outs
    __shared__ double sqw_s[32];
    __shared__ double exp_tab[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    exp_tab_load(exp_tab, tid);   // visible after the barrier at the first tile start
    const int g_low = lane & 7, n_a = warp * 8 + (lane >> 3), n_b = n_a + 4;
    const int D = p.D, M = p.M, DV = D + 1;
    const int em = blockIdx.y;
    const double* xchunks = p.bank ? p.bank[em].xchunks : p.xchunks;
    if (tid < 32) sqw_s[tid] = p.bank ? p.bank[em].sqrt_w[tid] : p.sqrt_w[tid];
    double* const o_mu = p.mu ? p.mu + em * p.eo_mu : nullptr;
    double* const o_deriv = p.deriv ? p.deriv + em * p.eo_deriv : nullptr;
    const int64_t ntiles = (p.N + TN - 1) / TN;
    bool x_resident = false;
    constexpr int XP = x_pitch(DP);   // row pitch of the training chunk (conflict-free LDS.128)
    const int chunk_doubles = p.JC * (XP + 1);

    // The test rows of the NEXT tile are fetched into registers while the current tile computes (the global-load
    // latency at every tile start showed up as 0.4 long-scoreboard + 0.2 barrier stall cycles per issue in ncu).
    constexpr int PF = (TN * DP + kMeanThreads - 1) / kMeanThreads;
    double pf[PF];
    auto fetch_rows = [&](int64_t t) {
        const int64_t m0 = t * TN;
        const int mpts = (int)min((int64_t)TN, p.N - m0);
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int e = tid + q * kMeanThreads;
            if (e < TN * D) {
                const int r = e / D;
                const int64_t src = (r < mpts) ? (m0 * D + e) : ((p.N - 1) * D + (e - r * D));
                pf[q] = __ldg(p.testing + src);
            }
        }
    };
    if ((int64_t)blockIdx.x < ntiles) fetch_rows(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * TN;
        const int npts = (int)min((int64_t)TN, p.N - n0);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int e = tid + q * kMeanThreads;
            if (e < TN * D) ts_s[e] = pf[q];
        }
        __syncthreads();
        double tsa[DP], tsb[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            tsa[d] = (d < D) ? ts_s[n_a * D + d] * sqw_s[d] : 0.0;
            tsb[d] = (d < D) ? ts_s[n_b * D + d] * sqw_s[d] : 0.0;
        }
        if (tile + gridDim.x < ntiles) fetch_rows(tile + gridDim.x);
        double va[NV], vb[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) va[i] = vb[i] = 0.0;
        for (int c = 0; c < p.nchunks; ++c) {
            if (!x_resident) {
                __syncthreads();
                const double2* src = reinterpret_cast<const double2*>(xchunks + (size_t)c * chunk_doubles);
                double2* dst = reinterpret_cast<double2*>(Xc);
                for (int e = tid; e < chunk_doubles / 2; e += kMeanThreads) dst[e] = __ldg(src + e);
                __syncthreads();
                if (p.nchunks == 1) x_resident = true;
            }
            const int jn = min(p.JC, M - c * p.JC);
            const double* al = Xc + p.JC * XP;
            int jl = g_low;
            if (jl < jn) {
                double2 xn[DP / 2];
                double aln;
                {
                    const double2* xr = reinterpret_cast<const double2*>(Xc + jl * XP);
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) xn[q] = xr[q];
                    aln = al[jl];
                }
                for (; jl < jn; jl += 8) {
                    double2 x[DP / 2];
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) x[q] = xn[q];
                    const double alj = aln;
                    {
                        const int jnx = min(jl + 8, jn - 1);
                        const double2* xr = reinterpret_cast<const double2*>(Xc + jnx * XP);
#pragma unroll
                        for (int q = 0; q < DP / 2; ++q) xn[q] = xr[q];
                        aln = al[jnx];
                    }
                    double ua[DP], ub[DP];
                    double ra = 0.0, rb = 0.0;
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) {
                        const int d = 2 * q;
                        ua[d] = x[q].x - tsa[d];
                        ub[d] = x[q].x - tsb[d];
                        ua[d + 1] = x[q].y - tsa[d + 1];
                        ub[d + 1] = x[q].y - tsb[d + 1];
                        ra = fma(ua[d], ua[d], ra);
                        rb = fma(ub[d], ub[d], rb);
                        ra = fma(ua[d + 1], ua[d + 1], ra);
                        rb = fma(ub[d + 1], ub[d + 1], rb);
                    }
                    const double ka = exp_neg_tab(-0.5 * ra, exp_tab), kb = exp_neg_tab(-0.5 * rb, exp_tab);
                    if (KSTAR) {
                        const int jg = c * p.JC + jl;
                        const size_t col = ((size_t)(jg >> 2) * 16) * 4 + (jg & 3);
                        const int64_t na = n0 + n_a, nb = n0 + n_b;
                        p.kstar[((size_t)(na >> 4) * p.kblk * 16 + (na & 15)) * 4 + col] = ka;
                        p.kstar[((size_t)(nb >> 4) * p.kblk * 16 + (nb & 15)) * 4 + col] = kb;
                    }
                    const double ca = ka * alj, cb = kb * alj;
                    va[0] += ca;
                    vb[0] += cb;
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        va[1 + d] = fma(ca, ua[d], va[1 + d]);
                        vb[1 + d] = fma(cb, ub[d], vb[1 + d]);
                    }
                }
            }
        }
        __syncthreads();   // everyone has read ts_s
        double* outs = ts_s;
        {
            RS<NV> rs;
            rs.run(va, vb, lane);
            double* dst = outs + ((lane & 4) ? n_b : n_a) * DV;
            const int half_base = (lane & 2) ? RS<NV>::H2 : 0;
            const int base3 = (lane & 1) ? RS<NV>::H3 : 0;
#pragma unroll
            for (int i = 0; i < RS<NV>::H3; ++i) {
                const int i2 = base3 + i, idx = half_base + i2;
                if (i2 < RS<NV>::H2 && idx < DV) dst[idx] = rs.r3[i];
            }
        }
        __syncthreads();
        if (o_mu != nullptr && tid < npts) o_mu[(n0 + tid) * p.ld_mu] = outs[tid * DV];
        if (o_deriv != nullptr) {
            for (int e = tid; e < npts * D; e += kMeanThreads) {
                const int r = e / D, d = e - r * D;
                o_deriv[(n0 + r) * p.ld_deriv + d] = sqw_s[d] * outs[r * DV + 1 + d];
            }
        }
    }
}

// Hessian for input dimensions beyond the register budget of the triangular kernel (D > 12): HR rows of the
// (symmetric) Hessian at a time, all DP columns, K* recomputed for every row block.  Same thread mapping and
// formulas as k_predict_mean; costs ~(DP / HR) x the K* work, which is acceptable for this completeness path.
template <int DP, int HR>
__global__ void __launch_bounds__(kMeanThreads) k_hessian_rows(const MeanParams p) {
    constexpr int TN = kMeanTN;
    extern __shared__ __align__(128) unsigned char smem_h[];
    double* Xc = reinterpret_cast<double*>(smem_h + p.off_xc);
    double* ts_s = reinterpret_cast<double*>(smem_h + p.off_ts);
    double* out_s = reinterpret_cast<double*>(smem_h + p.off_out);   // [TN][HR][D]
    __shared__ double sqw_s[32];
    __shared__ double exp_tab[64];   // 2^(j/64) for exp_neg_tab; visible after the first CTA barrier
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    exp_tab_load(exp_tab, tid);
    const int g_low = lane & 3, n_loc = warp * 8 + (lane >> 2);
    const int D = p.D, M = p.M;
    const int em = blockIdx.y;
    const double* xchunks = p.bank ? p.bank[em].xchunks : p.xchunks;
    if (tid < 32) sqw_s[tid] = p.bank ? p.bank[em].sqrt_w[tid] : p.sqrt_w[tid];
    double* const o_hess = p.hess + em * p.eo_hess;
    const int64_t ntiles = (p.N + TN - 1) / TN;
    constexpr int XP = x_pitch(DP);   // row pitch of the training chunk (conflict-free LDS.128)
    const int chunk_doubles = p.JC * (XP + 1);
    bool x_resident = false;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * TN;
        const int npts = (int)min((int64_t)TN, p.N - n0);
        __syncthreads();
        for (int e = tid; e < TN * D; e += kMeanThreads) {
            const int r = e / D;
            const int64_t src = (r < npts) ? (n0 * D + e) : ((p.N - 1) * D + (e - r * D));
            ts_s[e] = __ldg(p.testing + src);
        }
        __syncthreads();
        double ts[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) ts[d] = (d < D) ? ts_s[n_loc * D + d] * sqw_s[d] : 0.0;
        for (int r0 = 0; r0 < D; r0 += HR) {
            double T[HR][DP];
            double mu = 0.0;
#pragma unroll
            for (int i = 0; i < HR; ++i)
#pragma unroll
                for (int e = 0; e < DP; ++e) T[i][e] = 0.0;
            for (int c = 0; c < p.nchunks; ++c) {
                if (!x_resident) {
                    __syncthreads();
                    const double2* src = reinterpret_cast<const double2*>(xchunks + (size_t)c * chunk_doubles);
                    double2* dst = reinterpret_cast<double2*>(Xc);
                    for (int e = tid; e < chunk_doubles / 2; e += kMeanThreads) dst[e] = __ldg(src + e);
                    __syncthreads();
                    if (p.nchunks == 1) x_resident = true;
                }
                const int jn = min(p.JC, M - c * p.JC);
                const double* al = Xc + p.JC * XP;
                for (int jl = g_low; jl < jn; jl += 4) {
                    const double2* x1 = reinterpret_cast<const double2*>(Xc + jl * XP);
                    double u[DP];
                    double r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < DP; d += 2) {
                        const double2 a = x1[d >> 1];
                        u[d] = a.x - ts[d];
                        u[d + 1] = a.y - ts[d + 1];
                        r2 = fma(u[d], u[d], r2);
                        r2 = fma(u[d + 1], u[d + 1], r2);
                    }
                    const double cj = exp_neg_tab(-0.5 * r2, exp_tab) * al[jl];
                    mu += cj;
#pragma unroll
                    for (int i = 0; i < HR; ++i) {
                        // u[r0 + i] with a runtime r0: select from registers without dynamic indexing
                        double ur = 0.0;
#pragma unroll
                        for (int d = 0; d < DP; ++d) ur = (d == r0 + i) ? u[d] : ur;
                        const double cu = cj * ur;
#pragma unroll
                        for (int e = 0; e < DP; ++e) T[i][e] = fma(cu, u[e], T[i][e]);
                    }
                }
            }
            mu += __shfl_xor_sync(0xffffffffu, mu, 1);
            mu += __shfl_xor_sync(0xffffffffu, mu, 2);
            __syncthreads();   // out_s of the previous row block has been written out
#pragma unroll
            for (int i = 0; i < HR; ++i) {
#pragma unroll
                for (int e = 0; e < DP; ++e) {
                    double v = T[i][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int d = r0 + i;
                    if (g_low == 0 && d < D && e < D) {
                        double h = sqw_s[d] * sqw_s[e] * v;
                        if (d == e) h -= sqw_s[d] * sqw_s[d] * mu;
                        out_s[(n_loc * HR + i) * D + e] = h;
                    }
                }
            }
            __syncthreads();
            const int nr = min(HR, D - r0);
            for (int e = tid; e < npts * nr * D; e += kMeanThreads) {
                const int r = e / (nr * D), q = e - r * (nr * D);
                const int i = q / D, col = q - i * D;
                o_hess[(n0 + r) * p.ld_hess + (int64_t)(r0 + i) * D + col] = out_s[(r * HR + i) * D + col];
            }
        }
    }
}

}  // namespace gpe
