// Bank mean + input gradient with the input differences SHARED between emulators (BASELINE north_star item 5:
// "a batched multi-emulator path that shares test inputs across many hyperparameter sets").
//
// The emulators of a bank have the same training inputs and differ in their hyper-parameters (MultivariateEmulator:
// one GP per principal component, multivariate_gp.py:150-188 / 195-222; per-band banks, tests/test_perband_emulator.py),
// so for a (test point n, training point j) pair the differences u_d = x_jd - t_nd and their squares s_d = u_d^2 do not
// depend on the emulator.  A thread keeps them in registers and evaluates G emulators on them:
//     r_e   = sum_d (-w_ed / 2) s_d                  D FMA      (reference GaussianProcess.py:228-234)
//     k_e   = exp(r_e)                               10         (gpe_math.cuh::exp_neg_tab)
//     c_e   = k_e (b_e alpha_ej)                     1          (GaussianProcess.py:237)
//     mu_e += c_e,  g_ed += c_e u_d                  1 + D      (GaussianProcess.py:244-247; scaled by w_ed at the end)
// i.e. 2D + 12 + 2D / G FP64-pipe operations per (pair, emulator) instead of the 3D + 14 of the one-emulator kernels
// (predict_mean.cuh): 36 instead of 44 at D = 10, G = 5.  The per-emulator weights are read from shared memory as
// broadcast LDS.128 (the FP64 pipe, not the issue slots, is the limiter: an FP64 instruction occupies it for two cycles).
//
// Thread (n, g): 8 test rows x 4 training-point lanes per warp, 4 warps per CTA (32 points per tile), blockIdx.y = group
// of G emulators.  The four lanes of a point are combined by a two-level shuffle reduce-scatter, results are staged in
// shared memory and written point-major: mu (N, E), deriv (N, E, D) -- the layout of gpe_bank_predict.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_math.cuh"
#include "gpe_ptx.cuh"

namespace gpe {

// LDS.128 the compiler may not hoist: the G x DP weights are loop-invariant, and held in registers they would cost more
// than the accumulators
__device__ __forceinline__ double2 lds_pinned_v2(uint32_t saddr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}

constexpr int kBankThreads = 128;
constexpr int kBankTN = 32;

struct BankMeanParams {
    const double* testing;   // (N, D)
    int64_t N;
    double* mu;              // (N, E) or null
    double* deriv;           // (N, E, D) or null
    const double* xraw;      // [nchunks][JC][x_pitch(DP)]   raw training inputs, pad rows / dims zero
    const double* galpha;    // [ngroups][nchunks][JC][GP]  b_e alpha_ej, pad emulators / rows zero (GP = G rounded up to 2)
    const double* gw;        // [ngroups][2][G][DP]  first -w_ed / 2, then w_ed; pads zero
    int M, D, E, JC, nchunks;
    uint32_t off_x, off_a, off_w, off_ts, smem_need;   // shared-memory byte offsets / extent (host: plan_bank_mean)
};

// GRAD = false: means only (MultivariateEmulator.predict(do_deriv=False), forward modelling): the accumulators shrink to one
// per emulator, so groups of up to 10 fit and D + 11 + 2D / G = 23 operations per (pair, emulator) remain at D = 10.
template <int DP, int G, bool GRAD>
__global__ void __launch_bounds__(kBankThreads, GRAD ? 2 : 3) k_bank_mean(const BankMeanParams p) {
    constexpr int TN = kBankTN;
    constexpr int GP = (G + 1) & ~1;
    constexpr int XP = x_pitch(DP);           // row pitch of the training chunk (conflict-free LDS.128)
    constexpr int AV = GRAD ? DP + 1 : 1;     // accumulators per emulator: mu [, g_0 .. g_{DP-1}]
    constexpr int NV = G * AV;                // values per point
    constexpr int H1 = (NV + 1) / 2, H2 = (H1 + 1) / 2;
    extern __shared__ __align__(128) unsigned char smem_bm[];
    double* Xs = reinterpret_cast<double*>(smem_bm + p.off_x);     // [JC][XP]
    double* As = reinterpret_cast<double*>(smem_bm + p.off_a);     // [JC][GP]
    double* Ws = reinterpret_cast<double*>(smem_bm + p.off_w);     // [2][G][DP]
    double* ts_s = reinterpret_cast<double*>(smem_bm + p.off_ts);  // [TN][D] test rows, then [TN][NV] results
    __shared__ double exp_tab[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    smem_guard(p.smem_need);
    exp_tab_load(exp_tab, tid);
    const int g_low = lane & 3, n_loc = warp * 8 + (lane >> 2);
    const int D = p.D, M = p.M;
    const int grp = blockIdx.y, e0 = grp * G;
    const int gact = min(G, p.E - e0);
    const uint32_t ws_addr = (uint32_t)__cvta_generic_to_shared(Ws);
    const double* ga = p.galpha + (size_t)grp * p.nchunks * p.JC * GP;
    for (int i = tid; i < 2 * G * DP; i += kBankThreads) Ws[i] = __ldg(p.gw + (size_t)grp * 2 * G * DP + i);
    const int64_t ntiles = (p.N + TN - 1) / TN;
    bool resident = false;

    // test rows of the next tile travel through registers while the current tile computes
    constexpr int PF = (TN * DP + kBankThreads - 1) / kBankThreads;
    double pf[PF];
    auto fetch_rows = [&](int64_t t) {
        const int64_t m0 = t * TN;
        const int mpts = (int)min((int64_t)TN, p.N - m0);
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int e = tid + q * kBankThreads;
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
        __syncthreads();   // previous tile's staged results drained
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int e = tid + q * kBankThreads;
            if (e < TN * D) ts_s[e] = pf[q];
        }
        __syncthreads();
        double t[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) t[d] = (d < D) ? ts_s[n_loc * D + d] : 0.0;
        if (tile + gridDim.x < ntiles) fetch_rows(tile + gridDim.x);

        double acc[G][AV];
#pragma unroll
        for (int e = 0; e < G; ++e)
#pragma unroll
            for (int i = 0; i < AV; ++i) acc[e][i] = 0.0;

        for (int c = 0; c < p.nchunks; ++c) {
            if (!resident) {
                __syncthreads();
                const double2* sx = reinterpret_cast<const double2*>(p.xraw + (size_t)c * p.JC * XP);
                double2* dx = reinterpret_cast<double2*>(Xs);
                for (int e = tid; e < p.JC * XP / 2; e += kBankThreads) dx[e] = __ldg(sx + e);
                const double2* sa = reinterpret_cast<const double2*>(ga + (size_t)c * p.JC * GP);
                double2* da = reinterpret_cast<double2*>(As);
                for (int e = tid; e < p.JC * GP / 2; e += kBankThreads) da[e] = __ldg(sa + e);
                __syncthreads();
                if (p.nchunks == 1) resident = true;
            }
            const int jn = min(p.JC, M - c * p.JC);
            int jl = g_low;
            if (jl < jn) {
                double2 xn[DP / 2];
                {
                    const double2* xr = reinterpret_cast<const double2*>(Xs + jl * XP);
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) xn[q] = xr[q];
                }
                for (; jl < jn; jl += 4) {
                    double u[DP], s[DP];
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) {
                        u[2 * q] = xn[q].x - t[2 * q];
                        u[2 * q + 1] = xn[q].y - t[2 * q + 1];
                        s[2 * q] = u[2 * q] * u[2 * q];
                        s[2 * q + 1] = u[2 * q + 1] * u[2 * q + 1];
                    }
                    {
                        const double2* xr = reinterpret_cast<const double2*>(Xs + min(jl + 4, jn - 1) * XP);
#pragma unroll
                        for (int q = 0; q < DP / 2; ++q) xn[q] = xr[q];
                    }
                    const double2* ar = reinterpret_cast<const double2*>(As + jl * GP);
#pragma unroll
                    for (int e2 = 0; e2 < GP / 2; ++e2) {
                        const double2 a2 = ar[e2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int e = 2 * e2 + h;
                            if (e < G) {
                                double r = 0.0;
#pragma unroll
                                for (int q = 0; q < DP / 2; ++q) {
                                    const double2 w2 = lds_pinned_v2(ws_addr + (uint32_t)(e * DP + 2 * q) * 8u);
                                    r = fma(w2.x, s[2 * q], r);
                                    r = fma(w2.y, s[2 * q + 1], r);
                                }
                                const double cj = exp_neg_tab(r, exp_tab) * (h ? a2.y : a2.x);
                                acc[e][0] += cj;
                                if constexpr (GRAD) {
#pragma unroll
                                    for (int d = 0; d < DP; ++d) acc[e][1 + d] = fma(cj, u[d], acc[e][1 + d]);
                                }
                            }
                        }
                    }
                }
            }
        }

        // reduce-scatter over the 4 training-point lanes of a point: NV -> H1 -> H2 values per lane
        double r2[H2];
        {
            const bool b1 = lane & 2, b0 = lane & 1;
            double r1[2 * H2];
#pragma unroll
            for (int i = 0; i < H1; ++i) {
                const int ia = i, ib = H1 + i;
                const int ea = ia / AV, va = ia % AV;
                const int eb = ib / AV, vb = ib % AV;
                const double lo = acc[ea][va];
                const double hi = (ib < NV) ? acc[eb < G ? eb : 0][vb] : 0.0;
                const double send = b1 ? lo : hi, keep = b1 ? hi : lo;
                r1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
            for (int i = H1; i < 2 * H2; ++i) r1[i] = 0.0;
#pragma unroll
            for (int i = 0; i < H2; ++i) {
                const double send = b0 ? r1[i] : r1[H2 + i], keep = b0 ? r1[H2 + i] : r1[i];
                r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        }
        __syncthreads();   // everyone has read its test row
        double* outs = ts_s;   // [TN][NV]
        {
            const int base1 = (lane & 2) ? H1 : 0, base2 = (lane & 1) ? H2 : 0;
#pragma unroll
            for (int i = 0; i < H2; ++i) {
                const int i1 = base2 + i, idx = base1 + i1;
                if (i1 < H1 && idx < NV) outs[n_loc * NV + idx] = r2[i];
            }
        }
        __syncthreads();
        if (p.mu != nullptr) {
            for (int e = tid; e < npts * gact; e += kBankThreads) {
                const int r = e / gact, em = e - r * gact;
                p.mu[(n0 + r) * (int64_t)p.E + e0 + em] = outs[r * NV + em * AV];
            }
        }
        if (GRAD && p.deriv != nullptr) {
            const int gd = gact * D;
            const double* wout = Ws + G * DP;
            for (int e = tid; e < npts * gd; e += kBankThreads) {
                const int r = e / gd, q = e - r * gd, em = q / D, d = q - em * D;
                p.deriv[((n0 + r) * (int64_t)p.E + e0 + em) * D + d] = wout[em * DP + d] * outs[r * NV + em * AV + 1 + d];
            }
        }
    }
}

}  // namespace gpe
