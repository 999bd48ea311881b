// Fused FP64 GP prediction for one GP: mean + variance + input gradient in one pass (sm_100a).
//
// For a tile of TN test points a persistent CTA
//   phase A  forms K*[n][j] = exp(-1/2 sum_d (sqrt(w_d) x_jd - sqrt(w_d) t_nd)^2) in shared memory
//            (never written to HBM) and, in the same sweep, the mean  sum_j K* (b alpha_j)  and the gradient
//            sums  sum_j K* (b alpha_j) (xs_jd - ts_nd), combined across lanes by a shuffle reduce-scatter
//            -- reference GaussianProcess.py:232-237 and the D-loop :244-247;
//   phase B  contracts  G = K* (TN x M) . invQ^T (M x M)  on the FP64 tensor path (DMMA.8x8x4) with the whole
//            TN x Mp accumulator tile resident in registers, invQ streamed L2 -> smem by TMA bulk copies through
//            an mbarrier ring, then var_n = b - b^2 sum_j G_nj K*_nj  -- reference GaussianProcess.py:240.
//   phase C  (HESS variants) Hessian of the mean on the same tensor path.  With centred scaled coordinates
//            x' = sqrt(w) x - c, t' = sqrt(w) t - c:
//              sum_j k_j a_j (x'_jd - t'_d)(x'_je - t'_e) = S2_de - t'_d g_e - t'_e g_d - t'_d t'_e S0,
//            where S0 / g are the mean / gradient sums phase A already has and
//              S2 = K* (TN x M) . P (M x NC),  P[j][(d,e)] = b alpha_j x'_jd x'_je  (d <= e, NC = 8 ceil(D(D+1)/16))
//            is one more GEMM against the K* tile that is already in shared memory; P streams through the same ring
//            after invQ.  Reference GaussianProcess.py:345-366.  The expansion cancels digits when the training
//            inputs span many length scales, so the host only selects these variants when max |x'| is small
//            (gpemu.cu: hess_fused_ok); otherwise the direct kernel in predict_mean.cuh runs.
// The next tile's test rows are prefetched by a TMA bulk copy while the current tile computes.
//
// Layouts chosen for the hardware (built once at model upload, see gpemu.cu):
//   xchunks : per chunk of JC training points  [JC][DP] sqrt(w)-scaled inputs | [JC] b*alpha  (one bulk copy)
//   s_tiled : [ceil(M/4)][Mp][4]  s_tiled[kb][j][c] = invQ[j][4 kb + c], zero padded: a k-block of the B operand
//             is one contiguous 32*Mp-byte run whose smem image is bank-conflict-free for DMMA B fragments.
//   p_tiled : [ceil(M/4)][NC][4]  p_tiled[kb][col][c] = P[4 kb + c][col], same idea for the Hessian operand.
//   K* smem : [TN][Mp + 4] doubles; pitch = 4 (mod 16) doubles keeps the DMMA A-fragment loads conflict-free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "gpe_math.cuh"
#include "gpe_ptx.cuh"

namespace gpe {

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
    int nit;        // pipeline iterations per tile = ceil(kblk / KB), KB k-blocks per ring stage (template)
    int nstage;     // ring depth (<= 8)
    int lag;        // refill distance behind the consumer (1 or 2, < nstage)
    int JC;         // training points per phase-A chunk (multiple of 4)
    int symmetric;  // 1: s_tiled holds the upper-triangular fold of invQ (opt-in, half the DMMAs)
    int skew;       // cycles the second warp of every SM sub-partition waits before its first ring step (0: none)
    int nchunks;
    double b;       // signal variance exp(theta[D])
    // shared-memory carve-up (byte offsets)
    uint32_t off_bar, off_sqw, off_ks, off_bst, off_xc, off_ts, off_pa, off_vred;
    uint32_t ts_bytes;     // size of ONE of the two test-row / output staging buffers at off_ts
    uint32_t stage_bytes;
    // fused Hessian (HESS variants; hess == null skips phase C)
    double* hess;          // (N, D, D) or null
    int64_t ld_hess;       // element stride between consecutive points (D * D for a single GP)
    const double* p_tiled;
    int kbh;               // Hessian k-blocks per ring stage
    int nit_h;             // ceil(kblk / kbh)
    uint32_t off_hts;      // [TN][D] centred scaled test rows
    long long* trace;      // dev aid (normally null): CTA 0 stores clock64() at phase boundaries of its first 64 tiles
    double sqrt_w[kMaxD];
    double centre[kMaxD];  // c_d of the centred coordinates (HESS variants)
};

template <int I, int N, typename F>
__device__ __forceinline__ void static_for_impl(F& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for_impl<I + 1, N>(f);
    }
}
// f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N - 1>), in order
template <int N, typename F>
__device__ __forceinline__ void static_for(F f) {
    static_for_impl<0, N>(f);
}

template <int MT, int NT, int WR, int WC, int DP, int MINB, int KB, bool FULLNT, bool SYM, bool HESS>
__global__ void __launch_bounds__(WR * WC * 32, MINB) k_predict_full(const FullParams p) {
    constexpr int NW = WR * WC;           // warps per CTA (power of two)
    constexpr int NTHR = NW * 32;
    constexpr int TN = WR * MT * 8;       // test points per tile
    constexpr int NHI = TN / 8;           // warps along n in phase A
    constexpr int GH = NW / NHI;          // warps along j in phase A
    constexpr int TPR = NTHR / TN;        // threads per test row when writing outputs
    constexpr int NV = DP + 1;            // values reduced per point: mean + DP gradient sums
    constexpr int XP = x_pitch(DP);       // row pitch of the training chunk (conflict-free LDS.128)
    static_assert((NW & (NW - 1)) == 0, "warp count must be a power of two");
    static_assert(NHI * GH == NW && NHI >= 1 && GH >= 1, "TN must be 8 * (a divisor of the warp count)");
    static_assert(TPR >= 1 && TPR * TN == NTHR, "TN must divide the thread count");
    static_assert(DP % 2 == 0, "DP even (16-byte rows)");
    // phase C: the TN / 8 row tiles go to warp % NMT, the NCT column tiles are dealt cyclically to the CG = NW / NMT
    // warps that share a row tile (8 warps, TN = 64: one row tile and all column tiles per warp -- balanced across
    // the four SM sub-partitions, which warp % 4 maps onto)
    // cfg 0 / 1 (MT = 4): tile-start chores after the row barrier and the conflict-free epilogue loads (+0.9 % at
    // M = 250).  cfg 2 (MT = 2, 16-point tiles) measured 2 % slower with either change -- same source in the hot loop,
    // different ptxas schedule -- so it keeps the first arrangement.
    constexpr bool kLateChores = (MT >= 4);
    constexpr int NMT = TN / 8;
    constexpr int CG = (NW >= NMT) ? NW / NMT : 1;
    constexpr int NCT = (DP * (DP + 1) / 2 + 7) / 8;
    constexpr int NC = NCT * 8;
    constexpr int NTH = (NCT + CG - 1) / CG;
    static_assert(!HESS || (NW % NMT == 0 && NW >= NMT), "phase C needs the row tiles to divide the warps");
    static_assert(!(HESS && SYM), "the fused Hessian is built for the plain variance operand only");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* bar_empty = bar_full + 8;
    uint64_t* bar_x = bar_full + 16;
    uint64_t* bar_t = bar_full + 17;      // [2] test-row prefetch
    double* sqw_s = reinterpret_cast<double*>(smem + p.off_sqw);
    double* Ks = reinterpret_cast<double*>(smem + p.off_ks);
    unsigned char* Bst = smem + p.off_bst;
    double* Xc = reinterpret_cast<double*>(smem + p.off_xc);
    double* pa_s = reinterpret_cast<double*>(smem + p.off_pa);   // [GH][TN][D+1] (GH > 1 only)
    double* vred = reinterpret_cast<double*>(smem + p.off_vred); // [WC][TN]
    double* hts = reinterpret_cast<double*>(smem + p.off_hts);   // HESS: [TN][D] centred scaled test rows
    __shared__ int htab[HESS ? 256 : 1];                         // HESS: (d, e) -> d | e << 8 | column << 16
    __shared__ double cen_s[HESS ? kMaxD : 1];
    __shared__ double exp_tab[64];                               // 2^(j/64) for exp_neg_tab

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.D, M = p.M, Mp = p.Mp;
    const int pitch = Mp + 4;
    const int nstage = p.nstage;
    const int lag = p.lag;
    const int DV = D + 1;

    const int64_t ntiles = (p.N + TN - 1) / TN;
    const bool want_var = (p.var != nullptr);
    const bool want_hess = HESS && (p.hess != nullptr);
    // ring iterations per tile: [0, nit_b) stream invQ for phase B, [nit_b, nit_tot) stream P for phase C
    const int nit_b = want_var ? p.nit : 0;
    const int nit_tot = nit_b + (want_hess ? p.nit_h : 0);
    // bulk copies need 16-byte aligned sources; row blocks of full tiles are multiples of 64 bytes
    const bool ts_tma_ok = (reinterpret_cast<uintptr_t>(p.testing) & 15) == 0;
    const uint32_t ts_tile_bytes = (uint32_t)TN * D * 8u;

    {   // every buffer of the carve-up must lie inside the dynamic shared memory of this launch
        uint32_t ext = umax2(p.off_bar + 192u, p.off_sqw + 256u);
        ext = umax2(ext, p.off_ks + (uint32_t)TN * (uint32_t)pitch * 8u);
        ext = umax2(ext, p.off_ts + 2u * p.ts_bytes);
        ext = umax2(ext, umax2((uint32_t)TN * (uint32_t)DV * 8u, ts_tile_bytes) + p.off_ts + p.ts_bytes);
        if (GH > 1) ext = umax2(ext, p.off_pa + (uint32_t)GH * TN * (uint32_t)DV * 8u);
        ext = umax2(ext, p.off_vred + (uint32_t)WC * TN * 8u);
        if (want_hess) ext = umax2(ext, p.off_hts + (uint32_t)TN * (uint32_t)D * 8u);
        if (nit_tot > 0) ext = umax2(ext, p.off_bst + (uint32_t)nstage * p.stage_bytes);
        ext = umax2(ext, p.off_xc + (uint32_t)p.JC * (XP + 1) * 8u);
        smem_guard(ext);
    }
    // ---- one-time setup -------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], NW);
        }
        mbar_init(bar_x, 1);
        mbar_init(&bar_t[0], 1);
        mbar_init(&bar_t[1], 1);
        fence_mbar_init();
    }
    if (tid < kMaxD) sqw_s[tid] = p.sqrt_w[tid];
    exp_tab_load(exp_tab, tid);
    if (HESS) {
        if (tid < kMaxD) cen_s[tid] = p.centre[tid];
        for (int e = tid; e < p.D * p.D; e += NTHR) {
            const int d1 = e / p.D, d2 = e - d1 * p.D;
            const int lo = min(d1, d2), hi = max(d1, d2);
            htab[e] = d1 | (d2 << 8) | ((lo * p.D - lo * (lo - 1) / 2 + hi - lo) << 16);
        }
    }
    // zero the K* columns [M, Mp + 4): phase A never writes them, phase B multiplies them by zero rows of invQ
    for (int r = warp; r < TN; r += NW)
        for (int c = M + lane; c < pitch; c += 32) Ks[r * pitch + c] = 0.0;
    __syncthreads();

    // B-operand ring.  Global iteration g uses stage g % nstage; (cs, cpar) track the stage / parity of the next
    // iteration to consume.  Loads are issued by lane 0 of a rotating warp so no single warp carries the cost:
    // a burst of min(nstage, nit) at the start of a tile (every earlier use was released before the end-of-tile
    // barrier, so no wait is needed), then during iteration `it` the stage released in iteration it - lag is
    // refilled with the k-blocks of iteration it - lag + nstage, i.e. the copy runs nstage - lag iterations ahead.
    // lag = 2 when the ring is deep enough: with lag = 1 the issuing warp had to wait for the slowest warp of the
    // previous iteration almost every time (ncu: 43 try_wait spins per refill), which re-synchronised the warps.
    int cs = 0;
    uint32_t cpar = 0;
    auto issue = [&](int stage, int it_local) {  // k-blocks [KB * it_local, +KB) -> ring stage
        if (HESS && it_local >= nit_b) {             // phase C operand: k-blocks [kbh * (it_local - nit_b), +kbh) of P
            const int kb0 = (it_local - nit_b) * p.kbh;
            const uint32_t bytes = (uint32_t)min(p.kbh, p.kblk - kb0) * (uint32_t)NC * 32u;
            mbar_arrive_expect_tx(&bar_full[stage], bytes);
            tma_bulk_g2s(Bst + (size_t)stage * p.stage_bytes, p.p_tiled + (size_t)kb0 * NC * 4, bytes, &bar_full[stage]);
            return;
        }
        const int kb0 = it_local * KB;
        const int nkb = min(KB, p.kblk - kb0);
        unsigned char* dst = Bst + (size_t)stage * p.stage_bytes;
        if (!SYM) {
            const uint32_t bytes = (uint32_t)nkb * (uint32_t)Mp * 32u;
            mbar_arrive_expect_tx(&bar_full[stage], bytes);
            tma_bulk_g2s(dst, p.s_tiled + (size_t)kb0 * Mp * 4, bytes, &bar_full[stage]);
        } else {
            // only columns >= 8 (kb / 2) of k-block kb are non-zero: copy that tail into place
            uint32_t total = 0;
            for (int kk = 0; kk < nkb; ++kk) total += (uint32_t)(Mp - 8 * ((kb0 + kk) >> 1)) * 32u;
            mbar_arrive_expect_tx(&bar_full[stage], total);
            for (int kk = 0; kk < nkb; ++kk) {
                const int c0 = 8 * ((kb0 + kk) >> 1);
                tma_bulk_g2s(dst + ((size_t)kk * Mp + c0) * 32, p.s_tiled + ((size_t)(kb0 + kk) * Mp + c0) * 4,
                             (uint32_t)(Mp - c0) * 32u, &bar_full[stage]);
            }
        }
    };
    // start-of-tile burst: iteration i of the tile goes to stage (cs + i) % nstage, issued by lane 0 of warp i % NW (one
    // thread doing all of them kept its warp ~100 cycles per stage behind the others: 0.8 k cycles per tile with 8 stages)
    auto issue_burst = [&]() {   // called by every thread
        const int burst = min(nstage, nit_tot);
        if (lane == 0) {
            for (int i = warp; i < burst; i += NW) {
                int s = cs + i;
                if (s >= nstage) s -= nstage;
                issue(s, i);
            }
        }
    };
    auto refill = [&](int it) {   // called by every thread at the top of ring iteration `it` of the tile
        if (it >= lag && lane == 0 && warp == (it & (NW - 1)) && it - lag + nstage < nit_tot) {
            const int ps = (cs >= lag) ? cs - lag : cs - lag + nstage;   // stage of iteration it - lag
            const uint32_t ppar = (cs >= lag) ? cpar : (cpar ^ 1);      // parity of that use
            mbar_wait(&bar_empty[ps], ppar);
            issue(ps, it - lag + nstage);
        }
    };
    uint32_t xpar = 0;
    bool x_resident = false;

    // test-row prefetch state
    uint32_t tpar = 0;  // bit `buf` = parity of the next completion of bar_t[buf]
    auto prefetch_rows = [&](int64_t tile, int buf) -> bool {  // one thread; false if this tile must be loaded in-line
        const int64_t n0 = tile * TN;
        if (!ts_tma_ok || n0 + TN > p.N) return false;
        fence_proxy_async();  // the buffer was last written through the generic proxy (output staging)
        mbar_arrive_expect_tx(&bar_t[buf], ts_tile_bytes);
        tma_bulk_g2s(smem + p.off_ts + (size_t)buf * p.ts_bytes, p.testing + n0 * D, ts_tile_bytes, &bar_t[buf]);
        return true;
    };
    // whether tile `t` can be / was prefetched is a pure function of t, so every thread can evaluate it
    auto rows_prefetched = [&](int64_t tile) { return ts_tma_ok && (tile * TN + TN <= p.N); };

    // phase-A thread coordinates.  A warp owns 8 consecutive test rows; lane = 8 * n_low + g_low, g_low = training-
    // point lane (0..7), n_low = 0..3; each thread carries TWO test rows (n_a and n_a + 4) so that every training
    // row it pulls from shared memory feeds two (test, train) pairs.
    const int g_low = lane & 7, n_low = lane >> 3;
    const int n_hi = warp % NHI, g_hi = warp / NHI;
    const int n_a = n_hi * 8 + n_low, n_b = n_a + 4;
    // phase-B warp coordinates
    // SYM: the fold leaves column warp c with fewer active tiles than column warp c + 1, and warp w runs on SM
    // sub-partition w % 4 -- so every second group of four warps takes the column-warp indices in mirrored order, which
    // gives each sub-partition (each DMMA pipe) the same number of active tiles in every ring step
    const int wrow = warp / WC;
    const int wcol = (SYM && ((warp >> 2) & 1)) ? (((warp % WC) & ~3) | (3 - ((warp % WC) & 3))) : warp % WC;
    const int nt_act = FULLNT ? NT : p.nt_act;  // FULLNT: every column tile active, no predicates around the DMMAs

    int buf = 0;
    if (tid == 0 && blockIdx.x < ntiles) prefetch_rows(blockIdx.x, 0);
    int trace_tile = 0;
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
#define GPE_TRACE(k) do { if (tracing && trace_tile < 64) p.trace[trace_tile * 8 + (k)] = clock64(); } while (0)

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int64_t n0 = tile * TN;
        const int npts = (int)min((int64_t)TN, p.N - n0);
        double* ts_s = reinterpret_cast<double*>(smem + p.off_ts + (size_t)buf * p.ts_bytes);
        GPE_TRACE(0);
        // (16-point tiles keep the B-operand burst here, before the row wait: see kLateChores below)
        if (!kLateChores && nit_tot > 0) issue_burst();

        // ---- this tile's test rows: prefetched by TMA during the previous tile, or loaded in-line ------------
        if (rows_prefetched(tile)) {
            mbar_wait(&bar_t[buf], (tpar >> buf) & 1u);
            tpar ^= 1u << buf;
        } else {
            for (int e = tid; e < TN * D; e += NTHR) {
                const int r = e / D;
                const int64_t src = (r < npts) ? (n0 * D + e) : ((p.N - 1) * D + (e - r * D));
                ts_s[e] = __ldg(p.testing + src);
            }
            __syncthreads();
        }
        // DP > 16: two rows of test coordinates, differences and gradient sums (6 DP + 2 doubles) do not fit the register
        // file -- phase A then sweeps the training points twice, once per row (kTwoPass), and the first row's sums wait in
        // registers; the training rows are read from shared memory twice, which the FP64-bound loop does not notice
        // (measured at M = 250: D = 24 1.17e8 -> see DESIGN 4.1; the two-row loop spilled 1-3 KB per thread there)
        constexpr bool kTwoPass = DP > 16;
        double tsa[kTwoPass ? 1 : DP], tsb[kTwoPass ? 1 : DP];
        if constexpr (!kTwoPass) {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                tsa[d] = (d < D) ? ts_s[n_a * D + d] * sqw_s[d] : 0.0;
                tsb[d] = (d < D) ? ts_s[n_b * D + d] * sqw_s[d] : 0.0;
            }
        }
        if constexpr (HESS) {
            if (want_hess && g_hi == 0) {   // the 8 training-point lanes of a row share its D stores
#pragma unroll
                for (int d = 0; d < DP; ++d)
                    if ((d & 7) == g_low && d < D) {
                        hts[n_a * D + d] = tsa[d] - cen_s[d];
                        hts[n_b * D + d] = tsb[d] - cen_s[d];
                    }
            }
        }
        // phase-A chunk 0 (unless resident) is requested before the barrier so its latency overlaps the barrier
        if (!x_resident && tid == 0) {
            const uint32_t bytes = (uint32_t)p.JC * (XP + 1) * 8u;
            mbar_arrive_expect_tx(bar_x, bytes);
            tma_bulk_g2s(Xc, p.xchunks, bytes, bar_x);
        }
        __syncthreads();  // ts_s is reused as the output staging area below; the other buffer is free for prefetch
        // Two serial chores, ~0.5 k and ~0.3 k cycles each, given to lane 0 of two different warps AFTER the barrier so
        // that nobody waits for them (they overlap the other warps' phase A): the B-operand burst for this tile's
        // contraction (lands while phase A runs), and the TMA prefetch of the next tile's test rows.
        if (kLateChores && nit_tot > 0) issue_burst();
        {
            const int64_t next = tile + gridDim.x;
            if (tid == ((kLateChores && NW > 4) ? 32 * (NW - 1) : 0) && next < ntiles) prefetch_rows(next, buf ^ 1);
        }

        GPE_TRACE(1);
        // ---- phase A: K* tile + mean + gradient sums ----------------------------------------------------
        double va[NV], vb[NV];  // [0] mean, [1 + d] gradient sums of the two rows
#pragma unroll
        for (int i = 0; i < NV; ++i) va[i] = vb[i] = 0.0;

        for (int c = 0; c < p.nchunks; ++c) {
            if (!x_resident) {
                if (c > 0 && tid == 0) {
                    const uint32_t bytes = (uint32_t)p.JC * (XP + 1) * 8u;
                    mbar_arrive_expect_tx(bar_x, bytes);
                    tma_bulk_g2s(Xc, p.xchunks + (size_t)c * p.JC * (XP + 1), bytes, bar_x);
                }
                mbar_wait(bar_x, xpar);
                xpar ^= 1;
                if (p.nchunks == 1) x_resident = true;
            }
            const int jn = min(p.JC, M - c * p.JC);
            const double* al = Xc + p.JC * XP;
            double* krow_a = Ks + n_a * pitch + c * p.JC;
            double* krow_b = krow_a + 4 * pitch;
            if constexpr (kTwoPass) {
                auto sweep = [&](int n_r, double (&v1)[NV]) {
                    double* krow = Ks + n_r * pitch + c * p.JC;
                    double ts1[DP];
#pragma unroll
                    for (int d = 0; d < DP; ++d) ts1[d] = (d < D) ? ts_s[n_r * D + d] * sqw_s[d] : 0.0;
                    for (int jl = 8 * g_hi + g_low; jl < jn; jl += 8 * GH) {
                        const double2* xr = reinterpret_cast<const double2*>(Xc + jl * XP);
                        double u[DP];
                        double r = 0.0;
#pragma unroll
                        for (int q = 0; q < DP / 2; ++q) {
                            const double2 x2 = xr[q];
                            u[2 * q] = x2.x - ts1[2 * q];
                            u[2 * q + 1] = x2.y - ts1[2 * q + 1];
                            r = fma(u[2 * q], u[2 * q], r);
                            r = fma(u[2 * q + 1], u[2 * q + 1], r);
                        }
                        const double k = exp_neg_tab(-0.5 * r, exp_tab);
                        krow[jl] = k;
                        const double cj = k * al[jl];
                        v1[0] += cj;
#pragma unroll
                        for (int d = 0; d < DP; ++d) v1[1 + d] = fma(cj, u[d], v1[1 + d]);
                    }
                };
                sweep(n_a, va);
                sweep(n_b, vb);
            } else {
            int jl = 8 * g_hi + g_low;
                if (jl < jn) {
                    // software pipeline: the next training row and alpha are in flight while this one is consumed
                    double2 xn[DP / 2];
                    double aln;
                    {
                        const double2* xr = reinterpret_cast<const double2*>(Xc + jl * XP);
    #pragma unroll
                        for (int q = 0; q < DP / 2; ++q) xn[q] = xr[q];
                        aln = al[jl];
                    }
                    for (; jl < jn; jl += 8 * GH) {
                        double2 x[DP / 2];
    #pragma unroll
                        for (int q = 0; q < DP / 2; ++q) x[q] = xn[q];
                        const double alj = aln;
                        {
                            const int jnx = min(jl + 8 * GH, jn - 1);
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
                        const double ka = exp_neg_tab(-0.5 * ra, exp_tab);
                        const double kb = exp_neg_tab(-0.5 * rb, exp_tab);
                        krow_a[jl] = ka;
                        krow_b[jl] = kb;
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
            if (!x_resident) __syncthreads();  // all reads of Xc done before the next chunk lands
        }

        GPE_TRACE(2);
        // combine the 8 j-lanes of each point (reduce-scatter), then (GH > 1) the j-warps through smem
        double* outs = ts_s;  // [TN][D+1]: mean, then unscaled gradient sums
        if constexpr (kTwoPass) __syncthreads();   // the second sweep read its test row from ts_s: everyone is done with it
        {
            RS<NV> rs;
            rs.run(va, vb, lane);
            double* dst = ((GH == 1) ? outs : pa_s + g_hi * TN * DV) + ((lane & 4) ? n_b : n_a) * DV;
            const int half_base = (lane & 2) ? RS<NV>::H2 : 0;
            const int base3 = (lane & 1) ? RS<NV>::H3 : 0;
#pragma unroll
            for (int i = 0; i < RS<NV>::H3; ++i) {
                const int i2 = base3 + i;            // index inside this lane's half
                const int idx = half_base + i2;      // value index: 0 mean, 1 + d gradient
                if (i2 < RS<NV>::H2 && idx < DV) dst[idx] = rs.r3[i];
            }
            __syncthreads();
            if (GH > 1) {
                for (int e = tid; e < TN * DV; e += NTHR) {
                    double sum = 0.0;
                    for (int gh = 0; gh < GH; ++gh) sum += pa_s[gh * TN * DV + e];
                    outs[e] = sum;
                }
                __syncthreads();
            }
        }
        // K* tile and outs are now visible to every warp
        {
            const int r = tid / TPR, q = tid % TPR;
            if (r < npts) {
                if (p.mu != nullptr && q == 0) p.mu[(n0 + r) * p.ld_mu] = outs[r * DV];
                if (p.deriv != nullptr) {
                    for (int d = q; d < D; d += TPR) p.deriv[(n0 + r) * p.ld_deriv + d] = sqw_s[d] * outs[r * DV + 1 + d];
                }
            }
        }

        GPE_TRACE(3);
        // ---- phase B: variance contraction on the FP64 tensor path --------------------------------------
        if (want_var) {
            double acc[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

            const double* a_base = Ks + (wrow * MT * 8 + (lane >> 2)) * pitch + (lane & 3);
            // column tiles are dealt to the WC column-warps cyclically: warp wcol owns tiles wcol, wcol + WC, ...
            // (keeps the warps balanced when SYM skips the tiles below the diagonal)
            const int b_off = (wcol * 8 + (lane >> 2)) * 4 + (lane & 3);

            // KB k-blocks (4 values of the contraction index each) per ring stage, fully unrolled.  The A fragments
            // come from the K* tile, not from the ring, so they are fetched before waiting on the stage barrier.
            // One ring step with this warp's column tiles [JM, NT) (JM = 0: all of them), JM a compile-time constant so
            // that the tile loop carries no predicate and ptxas hoists the B-fragment loads.
            auto ring_step = [&](auto jm_c, int it) {
                constexpr int JM = decltype(jm_c)::value;
                refill(it);
                const int kb0 = it * KB;
                double a[KB][MT];
#pragma unroll
                for (int kk = 0; kk < KB; ++kk)
#pragma unroll
                    for (int i = 0; i < MT; ++i) a[kk][i] = a_base[i * 8 * pitch + (kb0 + kk) * 4];  // pad columns are 0
                mbar_wait(&bar_full[cs], cpar);
                const double* bs = reinterpret_cast<const double*>(Bst + (size_t)cs * p.stage_bytes) + b_off;
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) {
                    if (KB == 1 || kb0 + kk < p.kblk) {
                        const double* bk = bs + kk * Mp * 4;
#pragma unroll
                        for (int j = JM; j < NT; ++j) {
                            if (!FULLNT && j >= nt_act) break;   // real exit: predicated-off DMMAs are not free
                            const double bf = bk[j * (WC * 32)];
#pragma unroll
                            for (int i = 0; i < MT; ++i) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[kk][i], bf);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[cs]);
                if (++cs == nstage) { cs = 0; cpar ^= 1; }
            };
            // The two warps of an SM sub-partition (w and w + 4) otherwise run their ring steps in lockstep: both fetch A
            // fragments and wait on the stage barrier while the DMMA pipe idles, then both queue for it.  Holding one of
            // them back by a fraction of a step lets each warp's step prologue run under the other's DMMAs.
            if (p.skew > 0 && ((warp >> 2) & 1)) {
                const long long t0 = clock64();
                while (clock64() - t0 < p.skew) {}
            }
            if constexpr (!SYM) {
                for (int it = 0; it < nit_b; ++it) ring_step(std::integral_constant<int, 0>{}, it);
            } else {
                // SYM: the B operand is the upper-triangular fold of invQ: column tile t is all zero for k-block kb unless
                // t >= kb / 2, and this warp's tile j is t = wcol + WC j.  So its first active tile JM(it) =
                // ceil((t0 - wcol) / WC), t0 = (it KB) / 2, grows by one every 2 WC / KB ring steps: the loop over the ring
                // is cut into NT runs, each compiled for its own fixed tile range (the earlier version tested
                // `j < jmin` inside one unrolled loop and ran at 70 % of the halved work).
                int it = 0;
                static_for<NT>([&](auto jm_c) {
                    constexpr int JM = decltype(jm_c)::value;
                    const int it_end = min(nit_b, (WC * JM + wcol + 1) * (2 / KB));
                    for (; it < it_end; ++it) ring_step(jm_c, it);
                });
                // past this warp's last tile it only keeps the ring moving (barrier waits, arrivals, its refill turns)
                for (; it < nit_b; ++it) ring_step(std::integral_constant<int, NT>{}, it);
            }

            GPE_TRACE(4);
            // epilogue: var_n = b - b^2 sum_j G_nj K*_nj
            double vs[MT];
#pragma unroll
            for (int i = 0; i < MT; ++i) vs[i] = 0.0;
            // K* in C-fragment layout: row lane / 4, columns 2 (lane % 4) + {0, 1}.  One 16-byte load per lane would
            // put rows r and r + 1 of a quarter-warp on overlapping banks (pitch = 4 mod 16 doubles, chosen for the
            // A fragments): 2-way conflicts on all 131 KB of the tile, 2.1 k cycles per tile.  Two 8-byte loads with the
            // column order swapped on odd rows are conflict-free (each half-warp covers all 32 banks exactly once).
            const int odd = (lane >> 2) & 1;
            const double* k_base = Ks + (wrow * MT * 8 + (lane >> 2)) * pitch + wcol * 8 + 2 * (lane & 3);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (FULLNT || j < nt_act) {
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const double* kp = k_base + i * 8 * pitch + j * (WC * 8);
                        if (kLateChores) {
                            const double k_first = kp[odd], k_second = kp[odd ^ 1];
                            vs[i] = fma(acc[i][j][0], odd ? k_second : k_first, vs[i]);
                            vs[i] = fma(acc[i][j][1], odd ? k_first : k_second, vs[i]);
                        } else {
                            const double2 kk = *reinterpret_cast<const double2*>(kp);
                            vs[i] = fma(acc[i][j][0], kk.x, vs[i]);
                            vs[i] = fma(acc[i][j][1], kk.y, vs[i]);
                        }
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
        // ---- phase C: Hessian of the mean, S2 = K* . P on the tensor path -------------------------------
        if (HESS && want_hess) {
            GPE_TRACE(5);
            const int mt = warp % NMT, cg = warp / NMT;
            double hacc[NTH][2];
#pragma unroll
            for (int j = 0; j < NTH; ++j) hacc[j][0] = hacc[j][1] = 0.0;
            const double* a_base = Ks + (mt * 8 + (lane >> 2)) * pitch + (lane & 3);
            const int b_off = (cg * 8 + (lane >> 2)) * 4 + (lane & 3);
            for (int it = nit_b; it < nit_tot; ++it) {
                refill(it);
                const int kb0 = (it - nit_b) * p.kbh;
                const int nkb = min(p.kbh, p.kblk - kb0);
                mbar_wait(&bar_full[cs], cpar);
                const double* bs = reinterpret_cast<const double*>(Bst + (size_t)cs * p.stage_bytes) + b_off;
                const double* ap = a_base + kb0 * 4;
#pragma unroll 3
                for (int kk = 0; kk < nkb; ++kk) {
                    const double a = ap[kk * 4];
                    const double* bk = bs + kk * (NC * 4);
#pragma unroll
                    for (int j = 0; j < NTH; ++j) {
                        if (NCT % CG != 0 && cg + CG * j >= NCT) break;
                        dmma_m8n8k4(hacc[j][0], hacc[j][1], a, bk[j * (CG * 32)]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[cs]);
                if (++cs == nstage) { cs = 0; cpar ^= 1; }
            }
            GPE_TRACE(6);
            __syncthreads();  // every warp is done with the K* tile: its rows now stage S2 ([TN][NC], pitch as K*)
#pragma unroll
            for (int j = 0; j < NTH; ++j) {
                if (cg + CG * j < NCT)
                    *reinterpret_cast<double2*>(Ks + (mt * 8 + (lane >> 2)) * pitch + (cg + CG * j) * 8 + 2 * (lane & 3)) =
                        make_double2(hacc[j][0], hacc[j][1]);
            }
            __syncthreads();
            // H_de = sqrt(w_d w_e) (S2_de - t'_d g_e - t'_e g_d - t'_d t'_e S0) - [d == e] w_d S0; a warp writes whole
            // (D x D) blocks, consecutive lanes consecutive addresses
            // A lane owns the same (d, e) elements in every row, so their indices and weights are fetched once per
            // tile; the row loop then has NE independent load / FMA chains in flight.
            constexpr int NE = (DP * DP + 31) / 32;
            const int DD = D * D;
            int o1[NE], o2[NE], oc[NE];
            double ow[NE], od[NE];
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                const int e = min(lane + 32 * i, DD - 1);
                const int t = htab[e];
                o1[i] = t & 255; o2[i] = (t >> 8) & 255; oc[i] = t >> 16;
                ow[i] = sqw_s[o1[i]] * sqw_s[o2[i]];
                od[i] = (o1[i] == o2[i]) ? ow[i] : 0.0;
            }
#pragma unroll 2
            for (int r = warp; r < npts; r += NW) {
                const double* s2 = Ks + r * pitch;
                const double* tp = hts + r * D;
                const double* g = outs + r * DV + 1;
                const double s0 = outs[r * DV];
                double* dst = p.hess + (n0 + r) * p.ld_hess;
#pragma unroll
                for (int i = 0; i < NE; ++i) {
                    const double t1 = tp[o1[i]], t2 = tp[o2[i]];
                    double v = s2[oc[i]];
                    v = fma(-t1, g[o2[i]], v);
                    v = fma(-t2, fma(t1, s0, g[o1[i]]), v);
                    v = fma(v, ow[i], -od[i] * s0);
                    if (lane + 32 * i < DD) dst[lane + 32 * i] = v;
                }
            }
            GPE_TRACE(7);
        }
        __syncthreads();  // K*, outs, vred are free for the next tile
        if (!(HESS && want_hess)) GPE_TRACE(5);
        ++trace_tile;
    }
}

#undef GPE_TRACE

}  // namespace gpe
