// Fused FP64 GP prediction for one GP: mean + variance + input gradient in one pass (sm_100a).
//
// For a tile of TN test points a persistent CTA
//   phase A  forms K*[n][j] = exp(-1/2 sum_d (sqrt(w_d) x_jd - sqrt(w_d) t_nd)^2) in shared memory
//            (never written to HBM) and, in the same sweep, the mean  sum_j K* (b alpha_j)  and the gradient
//            sums  sum_j K* (b alpha_j) (xs_jd - ts_nd)  with warp-shuffle reductions
//            -- reference GaussianProcess.py:232-237 and the D-loop :244-247;
//   phase B  contracts  G = K* (TN x M) . invQ^T (M x M)  on the FP64 tensor path (DMMA.8x8x4) with the whole
//            TN x Mp accumulator tile resident in registers, invQ streamed L2 -> smem by TMA bulk copies through
//            an mbarrier ring, then var_n = b - b^2 sum_j G_nj K*_nj  -- reference GaussianProcess.py:240.
//
// Layouts chosen for the hardware (built once at model upload, see gpemu.cu):
//   xchunks : per chunk of JC training points  [JC][DP] sqrt(w)-scaled inputs | [JC] b*alpha  (one bulk copy)
//   s_tiled : [ceil(M/4)][Mp][4]  s_tiled[kb][j][c] = invQ[j][4 kb + c], zero padded: a k-block of the B operand
//             is one contiguous 32*Mp-byte run whose smem image is bank-conflict-free for DMMA B fragments.
//   K* smem : [TN][Mp + 4] doubles; pitch = 4 (mod 16) doubles makes both the phase-A stores (8 rows x 4 columns
//             per warp) and the DMMA A-fragment loads (same shape) conflict-free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_math.cuh"
#include "gpe_ptx.cuh"

namespace gpe {

constexpr int kFullThreads = 256;
constexpr int kMaxD = 32;

struct FullParams {
    const double* testing;  // (N, D) row-major
    int64_t N;
    double* mu;     // (N) or null
    double* var;    // (N) or null
    double* deriv;  // (N, D) or null
    int64_t ld_mu, ld_var, ld_deriv;  // element strides between consecutive points (1, 1, D for a single GP)
    const double* xchunks;
    const double* s_tiled;
    int M, D;
    int Mp;         // padded output width = WC * nt_act * 8
    int nt_act;     // active 8-column DMMA tiles per warp (<= NT)
    int kblk;       // ceil(M / 4) k-blocks
    int kbps;       // k-blocks per pipeline stage
    int nit;        // pipeline iterations per tile = ceil(kblk / kbps)
    int nstage;     // ring depth (power of two)
    int JC;         // training points per phase-A chunk (multiple of 4)
    int nchunks;
    double b;       // signal variance exp(theta[D])
    // shared-memory carve-up (byte offsets)
    uint32_t off_bar, off_sqw, off_ks, off_bst, off_xc, off_ts, off_pa, off_vred;
    uint32_t stage_bytes;
    double sqrt_w[kMaxD];
};

template <int MT, int NT, int WR, int WC, int DP>
__global__ void __launch_bounds__(kFullThreads, 1) k_predict_full(const FullParams p) {
    constexpr int TN = WR * MT * 8;       // test points per tile
    constexpr int NHI = TN / 8;           // warps along n in phase A
    constexpr int GH = 8 / NHI;           // warps along j in phase A
    static_assert(WR * WC == 8, "8 warps");
    static_assert(NHI * GH == 8 && NHI >= 1, "TN in {8,16,32,64}");
    static_assert(DP % 2 == 0, "DP even (16-byte rows)");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* bar_empty = bar_full + 8;
    uint64_t* bar_x = bar_full + 16;
    double* sqw_s = reinterpret_cast<double*>(smem + p.off_sqw);
    double* Ks = reinterpret_cast<double*>(smem + p.off_ks);
    unsigned char* Bst = smem + p.off_bst;
    double* Xc = reinterpret_cast<double*>(smem + p.off_xc);
    double* ts_s = reinterpret_cast<double*>(smem + p.off_ts);   // [TN][D] raw test rows, later outs [TN][D+1]
    double* pa_s = reinterpret_cast<double*>(smem + p.off_pa);   // [GH][TN][D+1] (GH > 1 only)
    double* vred = reinterpret_cast<double*>(smem + p.off_vred); // [WC][TN]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.D, M = p.M, Mp = p.Mp;
    const int pitch = Mp + 4;
    const int nstage = p.nstage;

    const int64_t ntiles = (p.N + TN - 1) / TN;
    const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool want_var = (p.var != nullptr);
    const int64_t total_it = want_var ? my_tiles * p.nit : 0;

    // ---- one-time setup -------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 8);
        }
        mbar_init(bar_x, 1);
        fence_mbar_init();
    }
    if (tid < kMaxD) sqw_s[tid] = p.sqrt_w[tid];
    // zero the K* columns [M, Mp + 4): phase A never writes them, phase B multiplies them by zero rows of invQ
    for (int r = warp; r < TN; r += 8)
        for (int c = M + lane; c < pitch; c += 32) Ks[r * pitch + c] = 0.0;
    __syncthreads();

    // producer state (thread 0 only): next pipeline iteration to issue
    int64_t issue_gi = 0;
    int issue_it = 0, issue_s = 0;
    auto issue_stage = [&]() {
        const int kb0 = issue_it * p.kbps;
        const int nkb = min(p.kbps, p.kblk - kb0);
        const uint32_t bytes = (uint32_t)nkb * (uint32_t)Mp * 32u;
        mbar_arrive_expect_tx(&bar_full[issue_s], bytes);
        tma_bulk_g2s(Bst + (size_t)issue_s * p.stage_bytes, p.s_tiled + (size_t)kb0 * Mp * 4, bytes,
                     &bar_full[issue_s]);
        ++issue_gi;
        if (++issue_it == p.nit) issue_it = 0;
        if (++issue_s == nstage) issue_s = 0;
    };
    if (tid == 0) {
        for (int s = 0; s < nstage && issue_gi < total_it; ++s) issue_stage();
    }

    // consumer ring state (all threads)
    int cs = 0;             // stage consumed next
    uint32_t cpar = 0;      // its parity
    int64_t cgi = 0;        // global pipeline iteration
    int es = 0;             // producer: stage whose release is awaited next (lags the consumer by one)
    uint32_t epar = 0;
    uint32_t xpar = 0;
    bool x_resident = false;

    // phase-A thread coordinates
    const int g_low = lane & 3, n_low = lane >> 2;
    const int n_hi = warp % NHI, g_hi = warp / NHI;
    const int n_loc = n_hi * 8 + n_low;
    // phase-B warp coordinates
    const int wrow = warp / WC, wcol = warp % WC;
    const int nt_act = p.nt_act;
    const int DV = D + 1;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * TN;
        const int npts = (int)min((int64_t)TN, p.N - n0);

        // ---- stage the tile's test rows (coalesced), pull mine into registers pre-scaled --------------
        for (int e = tid; e < TN * D; e += kFullThreads) {
            const int r = e / D;
            const int64_t src = (r < npts) ? (n0 * D + e) : ((p.N - 1) * D + (e - r * D));
            ts_s[e] = __ldg(p.testing + src);
        }
        __syncthreads();
        double ts[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) ts[d] = (d < D) ? ts_s[n_loc * D + d] * sqw_s[d] : 0.0;
        __syncthreads();  // ts_s is reused as the output staging area below

        // ---- phase A: K* tile + mean + gradient sums ----------------------------------------------------
        double mu = 0.0;
        double g[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) g[d] = 0.0;

        for (int c = 0; c < p.nchunks; ++c) {
            if (!x_resident) {
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)p.JC * (DP + 1) * 8u;
                    mbar_arrive_expect_tx(bar_x, bytes);
                    tma_bulk_g2s(Xc, p.xchunks + (size_t)c * p.JC * (DP + 1), bytes, bar_x);
                }
                mbar_wait(bar_x, xpar);
                xpar ^= 1;
                if (p.nchunks == 1) x_resident = true;
            }
            const int jn = min(p.JC, M - c * p.JC);
            const double* al = Xc + p.JC * DP;
            double* krow = Ks + n_loc * pitch + c * p.JC;
            int jl = 4 * g_hi + g_low;
            // two training points per trip: two independent exp chains in flight
            for (; jl + 4 * GH < jn; jl += 8 * GH) {
                const int jl2 = jl + 4 * GH;
                const double2* x1 = reinterpret_cast<const double2*>(Xc + jl * DP);
                const double2* x2 = reinterpret_cast<const double2*>(Xc + jl2 * DP);
                double u1[DP], u2[DP];
                double r1 = 0.0, r2 = 0.0;
#pragma unroll
                for (int d = 0; d < DP; d += 2) {
                    const double2 a1 = x1[d >> 1], a2 = x2[d >> 1];
                    u1[d] = a1.x - ts[d];
                    u1[d + 1] = a1.y - ts[d + 1];
                    u2[d] = a2.x - ts[d];
                    u2[d + 1] = a2.y - ts[d + 1];
                    r1 = fma(u1[d], u1[d], r1);
                    r2 = fma(u2[d], u2[d], r2);
                    r1 = fma(u1[d + 1], u1[d + 1], r1);
                    r2 = fma(u2[d + 1], u2[d + 1], r2);
                }
                const double k1 = exp_neg(-0.5 * r1);
                const double k2 = exp_neg(-0.5 * r2);
                krow[jl] = k1;
                krow[jl2] = k2;
                const double c1 = k1 * al[jl], c2 = k2 * al[jl2];
                mu += c1;
                mu += c2;
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    g[d] = fma(c1, u1[d], g[d]);
                    g[d] = fma(c2, u2[d], g[d]);
                }
            }
            for (; jl < jn; jl += 4 * GH) {
                const double2* x1 = reinterpret_cast<const double2*>(Xc + jl * DP);
                double u1[DP];
                double r1 = 0.0;
#pragma unroll
                for (int d = 0; d < DP; d += 2) {
                    const double2 a1 = x1[d >> 1];
                    u1[d] = a1.x - ts[d];
                    u1[d + 1] = a1.y - ts[d + 1];
                    r1 = fma(u1[d], u1[d], r1);
                    r1 = fma(u1[d + 1], u1[d + 1], r1);
                }
                const double k1 = exp_neg(-0.5 * r1);
                krow[jl] = k1;
                const double c1 = k1 * al[jl];
                mu += c1;
#pragma unroll
                for (int d = 0; d < DP; ++d) g[d] = fma(c1, u1[d], g[d]);
            }
            if (!x_resident) __syncthreads();  // all reads of Xc done before the next chunk lands
        }

        // reduce the 4 j-lanes of each point, then (GH > 1) the j-warps through smem
        mu += __shfl_xor_sync(0xffffffffu, mu, 1);
        mu += __shfl_xor_sync(0xffffffffu, mu, 2);
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            if (d < D) {
                g[d] += __shfl_xor_sync(0xffffffffu, g[d], 1);
                g[d] += __shfl_xor_sync(0xffffffffu, g[d], 2);
            }
        }
        double* outs = ts_s;  // [TN][D+1]: mean, then unscaled gradient sums
        if (GH == 1) {
            if (g_low == 0) {
                outs[n_loc * DV] = mu;
#pragma unroll
                for (int d = 0; d < DP; ++d)
                    if (d < D) outs[n_loc * DV + 1 + d] = g[d];
            }
            __syncthreads();
        } else {
            if (g_low == 0) {
                double* dst = pa_s + (g_hi * TN + n_loc) * DV;
                dst[0] = mu;
#pragma unroll
                for (int d = 0; d < DP; ++d)
                    if (d < D) dst[1 + d] = g[d];
            }
            __syncthreads();
            for (int e = tid; e < TN * DV; e += kFullThreads) {
                double s = 0.0;
                for (int gh = 0; gh < GH; ++gh) s += pa_s[gh * TN * DV + e];
                outs[e] = s;
            }
            __syncthreads();
        }
        // K* tile and outs are now visible to every warp
        if (p.mu != nullptr && tid < npts) p.mu[(n0 + tid) * p.ld_mu] = outs[tid * DV];
        if (p.deriv != nullptr) {
            for (int e = tid; e < npts * D; e += kFullThreads) {
                const int r = e / D, d = e - r * D;
                p.deriv[(n0 + r) * p.ld_deriv + d] = sqw_s[d] * outs[r * DV + 1 + d];
            }
        }

        // ---- phase B: variance contraction on the FP64 tensor path --------------------------------------
        if (want_var) {
            double acc[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

            const double* a_base = Ks + (wrow * MT * 8 + (lane >> 2)) * pitch + (lane & 3);
            const int b_off = (wcol * nt_act * 8 + (lane >> 2)) * 4 + (lane & 3);

            for (int it = 0; it < p.nit; ++it) {
                // producer: refill the stage released one iteration ago (its consumers are almost surely done),
                // so the copy for iteration cgi - 1 + nstage overlaps the DMMAs of iterations cgi .. cgi + nstage - 2
                if (tid == 0 && cgi >= 1 && issue_gi < total_it) {
                    mbar_wait(&bar_empty[es], epar);
                    if (++es == nstage) { es = 0; epar ^= 1; }
                    issue_stage();
                }
                mbar_wait(&bar_full[cs], cpar);
                const double* bs = reinterpret_cast<const double*>(Bst + (size_t)cs * p.stage_bytes) + b_off;
                const int kb0 = it * p.kbps;
                const int nkb = min(p.kbps, p.kblk - kb0);
                for (int kk = 0; kk < nkb; ++kk) {
                    double a[MT];
#pragma unroll
                    for (int i = 0; i < MT; ++i) a[i] = a_base[i * 8 * pitch + (kb0 + kk) * 4];
                    const double* bk = bs + kk * Mp * 4;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        if (j < nt_act) {
                            const double bf = bk[j * 32];
#pragma unroll
                            for (int i = 0; i < MT; ++i) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], bf);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[cs]);
                ++cgi;
                if (++cs == nstage) { cs = 0; cpar ^= 1; }
            }

            // epilogue: var_n = b - b^2 sum_j G_nj K*_nj
            double vs[MT];
#pragma unroll
            for (int i = 0; i < MT; ++i) vs[i] = 0.0;
            const double* k_base = Ks + (wrow * MT * 8 + (lane >> 2)) * pitch + wcol * nt_act * 8 + 2 * (lane & 3);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (j < nt_act) {
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const double2 kk = *reinterpret_cast<const double2*>(k_base + i * 8 * pitch + j * 8);
                        vs[i] = fma(acc[i][j][0], kk.x, vs[i]);
                        vs[i] = fma(acc[i][j][1], kk.y, vs[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                vs[i] += __shfl_xor_sync(0xffffffffu, vs[i], 1);
                vs[i] += __shfl_xor_sync(0xffffffffu, vs[i], 2);
                if ((lane & 3) == 0) vred[wcol * TN + wrow * MT * 8 + i * 8 + (lane >> 2)] = vs[i];
            }
            __syncthreads();
            if (tid < npts) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < WC; ++w) v += vred[w * TN + tid];
                p.var[(n0 + tid) * p.ld_var] = p.b - p.b * p.b * v;
            }
        }
        __syncthreads();  // K*, outs, vred are free for the next tile
    }
}

}  // namespace gpe
