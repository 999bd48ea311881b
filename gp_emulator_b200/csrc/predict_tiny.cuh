// Mean + variance + gradient for a HANDFUL of test points (FP64): one thread-block CLUSTER of 8 CTAs per 16-point tile.
//
// The reference is mostly called with one point, or a few, at a time (MultivariateEmulator.predict inside an inversion,
// GaussianProcess.predict per pixel: GaussianProcess.py:327-341), and then a call is pure latency.  The fused kernel
// (predict_full.cuh) gives a tile to ONE CTA, which walks the whole M x M contraction as 63 dependent ring steps: ~25 us
// at M = 250 on an otherwise empty GPU.  Here the column tiles of invQ are dealt to the 8 CTAs of a cluster and, inside a
// CTA, the contraction index to its 8 warps, so a warp issues 1/64 of the DMMAs (8 k-blocks x <= 4 tiles x 2 row tiles);
// B fragments come straight from L2 (no ring to fill), the per-warp accumulators are summed through shared memory in a
// fixed order, and the 8 partial quadratic forms of a point meet in the shared memory of cluster rank 0 (distributed shared
// memory, one cluster barrier).  Every CTA forms the K* tile itself (16 x M exponentials: cheaper than passing it around);
// rank 0 also writes mean and gradient.  Same formulas as predict_full.cuh (reference GaussianProcess.py:228-247), same
// operands (xchunks of the fused plan, s_tiled), summation order of its own: results agree with the other plans to
// rounding, as the 16- and 64-point plans do with each other.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_math.cuh"
#include "gpe_ptx.cuh"

namespace gpe {

constexpr int kTinyTN = 16;        // points per cluster
constexpr int kTinyThreads = 256;
constexpr int kTinyCluster = 8;

struct TinyParams {
    const double* testing;   // (N, D)
    int64_t N;
    double* mu;              // strided outputs; mu / deriv may be null, var is always written
    double* var;
    double* deriv;
    int64_t ld_mu, ld_var, ld_deriv;
    const double* xchunks;   // [M4 * x_pitch(DP) scaled inputs | M4 b*alpha]  (the fused plan's single chunk)
    const double* s_tiled;   // [kblk][Mp][4] invQ (or its fold)
    int M, D, JC, Mp, kblk;
    double b;
    double sqrt_w[32];
};

// dynamic shared memory: Ks [16][Mp + 4] | work: max( Xc [JC][XP + 1] + ts [16][D],  red [16][16][DP + 1],  accs [8][16][32] )
template <int DP>
__global__ void __cluster_dims__(kTinyCluster, 1, 1) __launch_bounds__(kTinyThreads) k_predict_tiny(const TinyParams p) {
    namespace cg = cooperative_groups;
    constexpr int XP = x_pitch(DP), NV = DP + 1;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) double smem_t[];
    const int M = p.M, D = p.D, Mp = p.Mp, pitch = Mp + 4;
    double* Ks = smem_t;                          // [16][pitch]
    double* work = Ks + kTinyTN * pitch;
    double* Xc = work;                            // phase A
    double* ts = Xc + p.JC * (XP + 1);            // [16][D]
    __shared__ double exp_tab[64];
    __shared__ double sqw_s[32];
    __shared__ double part[kTinyCluster][kTinyTN];   // rank 0: the partial quadratic forms of the cluster
    __shared__ double rowsum[kTinyTN][33];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // the carve-up must lie inside the dynamic shared memory of this launch (same check as every other kernel)
        const uint32_t w0 = (uint32_t)p.JC * (XP + 1) + (uint32_t)kTinyTN * (uint32_t)D;
        const uint32_t wk = umax2(umax2(w0, 16u * 16u * NV), 8u * 16u * 32u);
        smem_guard(((uint32_t)kTinyTN * (uint32_t)pitch + wk) * 8u);
    }
    const int64_t tile = blockIdx.x / kTinyCluster;
    const int64_t n0 = tile * kTinyTN;
    exp_tab_load(exp_tab, tid);
    if (tid < 32) sqw_s[tid] = p.sqrt_w[tid];
    {
        const double2* src = reinterpret_cast<const double2*>(p.xchunks);
        double2* dst = reinterpret_cast<double2*>(Xc);
        for (int e = tid; e < p.JC * (XP + 1) / 2; e += kTinyThreads) dst[e] = __ldg(src + e);
        for (int e = tid; e < kTinyTN * D; e += kTinyThreads) {
            const int r = e / D, d = e - r * D;
            ts[e] = __ldg(p.testing + min(n0 + r, p.N - 1) * D + d);   // ragged tile: repeat the last row
        }
        for (int e = tid; e < kTinyTN * (pitch - M); e += kTinyThreads) {   // K* pad columns multiply zero rows of invQ
            const int r = e / (pitch - M);
            Ks[r * pitch + M + (e - r * (pitch - M))] = 0.0;
        }
    }
    __syncthreads();
    // ---- phase A: K* tile, mean and gradient sums.  thread = (point n, training-point lane g of 16) -------------------
    const int n = tid & 15, g = tid >> 4;
    double v[NV];
    {
        double t1[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) t1[d] = (d < D) ? ts[n * D + d] * sqw_s[d] : 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = 0.0;
        const double* al = Xc + p.JC * XP;
        for (int j = g; j < M; j += 16) {
            const double2* xr = reinterpret_cast<const double2*>(Xc + j * XP);
            double u[DP];
            double r = 0.0;
#pragma unroll
            for (int q = 0; q < DP / 2; ++q) {
                const double2 x2 = xr[q];
                u[2 * q] = x2.x - t1[2 * q];
                u[2 * q + 1] = x2.y - t1[2 * q + 1];
                r = fma(u[2 * q], u[2 * q], r);
                r = fma(u[2 * q + 1], u[2 * q + 1], r);
            }
            const double k = exp_neg_tab(-0.5 * r, exp_tab);
            Ks[n * pitch + j] = k;
            const double c = k * al[j];
            v[0] += c;
#pragma unroll
            for (int d = 0; d < DP; ++d) v[1 + d] = fma(c, u[d], v[1 + d]);
        }
    }
    __syncthreads();   // K* complete; Xc / ts no longer needed: `work` is reused below
    if (rank == 0 && (p.mu != nullptr || p.deriv != nullptr)) {
        double* red = work;   // [16 g][16 n][NV]
#pragma unroll
        for (int i = 0; i < NV; ++i) red[(g * 16 + n) * NV + i] = v[i];
        __syncthreads();
        for (int e = tid; e < kTinyTN * (D + 1); e += kTinyThreads) {
            const int r = e / (D + 1), i = e - r * (D + 1);
            double s = 0.0;
#pragma unroll
            for (int gg = 0; gg < 16; ++gg) s += red[(gg * 16 + r) * NV + i];
            if (n0 + r < p.N) {
                if (i == 0) { if (p.mu != nullptr) p.mu[(n0 + r) * p.ld_mu] = s; }
                else if (p.deriv != nullptr) p.deriv[(n0 + r) * p.ld_deriv + i - 1] = sqw_s[i - 1] * s;
            }
        }
        __syncthreads();
    }
    // ---- phase B: this CTA's column tiles t = rank, rank + 8, ...; warp w takes the k-blocks kb = w, w + 8, ... ---------
    const int T = Mp >> 3;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    {
        const double* a_base = Ks + (lane >> 2) * pitch + (lane & 3);
        const double* b_base = p.s_tiled + ((size_t)rank * 8 + (lane >> 2)) * 4 + (lane & 3);
        for (int kb = warp; kb < p.kblk; kb += 8) {
            const double a0 = a_base[kb * 4], a1 = a_base[8 * pitch + kb * 4];
            const double* bk = b_base + (size_t)kb * Mp * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (rank + kTinyCluster * j < T) {
                    const double bf = __ldg(bk + j * (kTinyCluster * 32));
                    dmma_m8n8k4(acc[0][j][0], acc[0][j][1], a0, bf);
                    dmma_m8n8k4(acc[1][j][0], acc[1][j][1], a1, bf);
                }
            }
        }
    }
    // sum the 8 warps' accumulators in a fixed order, multiply with the K* columns, reduce to one value per point
    {
        double* accs = work;   // [8 warps][16 values][32 lanes]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) accs[(warp * 16 + (i * 4 + j) * 2 + c) * 32 + lane] = acc[i][j][c];
        __syncthreads();
        // element e = (value index, lane): G[row = i * 8 + lane / 4][col = (rank + 8 j) * 8 + 2 (lane % 4) + c]
        for (int e = tid; e < 16 * 32; e += kTinyThreads) {
            const int vi = e >> 5, l = e & 31;
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += accs[(w * 16 + vi) * 32 + l];
            const int i = vi >> 3, j = (vi >> 1) & 3, c = vi & 1;
            const int row = i * 8 + (l >> 2), t = rank + kTinyCluster * j;
            const double kv = (t < T) ? Ks[row * pitch + t * 8 + 2 * (l & 3) + c] : 0.0;
            // 32 products per row: (j, c, l % 4) -> slot
            rowsum[row][(j * 2 + c) * 4 + (l & 3)] = s * kv;
        }
        __syncthreads();
        if (tid < kTinyTN) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 32; ++q) s += rowsum[tid][q];
            double* dst = cluster.map_shared_rank(&part[0][0], 0);   // rank 0's copy of `part`
            dst[rank * kTinyTN + tid] = s;
        }
    }
    cluster.sync();
    if (rank == 0 && tid < kTinyTN && n0 + tid < p.N) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < kTinyCluster; ++r) s += part[r][tid];
        p.var[(n0 + tid) * p.ld_var] = p.b - p.b * p.b * s;
    }
}

}  // namespace gpe
