// GP prediction for ANY number of inputs D (FP64): the path for D > 32, where the per-D compiled kernels
// (predict_full.cuh, predict_mean.cuh: test rows and gradient sums in registers) do not exist.
//
// The reference has no limit on D (GaussianProcess.py:228-247 loops `for d in range(self.D)`); emulators with more than
// 32 inputs are rare, so this is a completeness path: correct and parity-tested, not tuned.  It reuses the K* scratch and
// the column-pass variance kernel of the large-M path (predict_var_large.cuh), so the variance still runs on the FP64
// tensor cores:
//   k_generic_kstar   K*_nj = exp(-1/2 sum_d w_d (x_jd - t_nd)^2) for a tile of 16 points against all M training points,
//                     into the scratch [tile][k-block][16][4]; mu_n = sum_j K*_nj b alpha_j         (GaussianProcess.py:234-237)
//   k_generic_grad    deriv_nd = w_d sum_j K*_nj b alpha_j (x_jd - t_nd), K* read back from the scratch        (:244-247)
//   k_generic_hess    hess_nde = sum_j K*_nj b alpha_j [w_d (x_jd - t_nd) w_e (x_je - t_ne) - delta_de w_d]    (:345-366)
//   k_var_large       var_n = b - b^2 sum_ij K*_ni invQ_ij K*_nj                                                (:240)
// Coordinates are pre-scaled by sqrt(w_d) (x' on the host, t' when a tile's rows are staged).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_math.cuh"

namespace gpe {

constexpr int kGenThreads = 256;
constexpr int kGenTN = 16;     // points per tile = tile of the K* scratch
constexpr int kGenJC = 32;     // training rows staged per step

struct GenericParams {
    const double* testing;   // (N, D)
    int64_t N;
    double* mu;              // strided outputs (ld_* = element stride between points); any may be null
    double* deriv;
    double* hess;
    int64_t ld_mu, ld_deriv, ld_hess;
    const double* xs;        // (M, D) training inputs scaled by sqrt(w_d)
    const double* alpha;     // (M) b * invQt
    const double* sqw;       // (D) sqrt(w_d)
    double* kstar;           // scratch [ceil(N/16)][kblk][16][4]; pads (j >= M) stay zero
    double* mu_tmp;          // (N) means of this sub-batch (the Hessian's diagonal term needs them)
    int M, D, kblk;
};

// shared memory: ts [16][D + 1] | xc [kGenJC][D + 1] | red [8][16]
__global__ void __launch_bounds__(kGenThreads) k_generic_kstar(const GenericParams p) {
    extern __shared__ __align__(16) double smem_g[];
    const int D = p.D, Dp = D + 1, M = p.M;   // odd-ish pitch: the 16 rows of a half-warp land on distinct banks
    double* ts = smem_g;
    double* xc = ts + kGenTN * Dp;
    double* red = xc + kGenJC * Dp;
    __shared__ double exp_tab[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    exp_tab_load(exp_tab, tid);
    const int n = tid & 15, jj = tid >> 4;   // 16 points x 16 training-point lanes
    const int64_t ntiles = (p.N + kGenTN - 1) / kGenTN;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * kGenTN;
        __syncthreads();
        for (int e = tid; e < kGenTN * D; e += kGenThreads) {
            const int r = e / D, d = e - r * D;
            const int64_t row = min(n0 + r, p.N - 1);          // ragged last tile: repeat the last row
            ts[r * Dp + d] = __ldg(p.testing + row * D + d) * __ldg(p.sqw + d);
        }
        double* ktile = p.kstar + (size_t)tile * p.kblk * 64;
        double mu = 0.0;
        for (int j0 = 0; j0 < M; j0 += kGenJC) {
            const int jn = min(kGenJC, M - j0);
            __syncthreads();
            for (int e = tid; e < jn * D; e += kGenThreads) {
                const int r = e / D, d = e - r * D;
                xc[r * Dp + d] = __ldg(p.xs + (size_t)(j0 + r) * D + d);
            }
            __syncthreads();
            for (int jl = jj; jl < jn; jl += 16) {
                const double* xr = xc + jl * Dp;
                const double* tr = ts + n * Dp;
                double r2 = 0.0;
                for (int d = 0; d < D; ++d) {
                    const double u = xr[d] - tr[d];
                    r2 = fma(u, u, r2);
                }
                const double k = exp_neg_tab(-0.5 * r2, exp_tab);
                const int j = j0 + jl;
                ktile[(size_t)(j >> 2) * 64 + n * 4 + (j & 3)] = k;
                mu = fma(k, __ldg(p.alpha + j), mu);
            }
        }
        // mean: the 16 lanes of a point sit in lanes n and n + 16 of the 8 warps
        mu += __shfl_xor_sync(0xffffffffu, mu, 16);
        if (lane < 16) red[warp * 16 + lane] = mu;
        __syncthreads();
        if (tid < 16 && n0 + tid < p.N) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kGenThreads / 32; ++w) s += red[w * 16 + tid];
            p.mu_tmp[n0 + tid] = s;
            if (p.mu != nullptr) p.mu[(n0 + tid) * p.ld_mu] = s;
        }
    }
}

// deriv: thread (n, d) of a 16-point tile sums over the training points; grid.y strides the dimensions in blocks of 16
__global__ void __launch_bounds__(kGenThreads) k_generic_grad(const GenericParams p) {
    const int tid = threadIdx.x, D = p.D, M = p.M;
    const int n = tid >> 4, dl = tid & 15;   // consecutive threads: consecutive d (coalesced x rows and outputs)
    const int64_t ntiles = (p.N + kGenTN - 1) / kGenTN;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = tile * kGenTN + n;
        const double* ktile = p.kstar + (size_t)tile * p.kblk * 64 + n * 4;
        for (int d = blockIdx.y * 16 + dl; d < D; d += gridDim.y * 16) {
            const double sw = __ldg(p.sqw + d);
            const double t = __ldg(p.testing + min(row, p.N - 1) * D + d) * sw;
            double g = 0.0;
            for (int j = 0; j < M; ++j) {
                const double c = ktile[(size_t)(j >> 2) * 64 + (j & 3)] * __ldg(p.alpha + j);
                g = fma(c, __ldg(p.xs + (size_t)j * D + d) - t, g);
            }
            if (row < p.N) p.deriv[row * p.ld_deriv + d] = sw * g;
        }
    }
}

// hess: one CTA per point and pass over the (d, e) pairs, d <= e; both triangles written
__global__ void __launch_bounds__(kGenThreads) k_generic_hess(const GenericParams p) {
    const int tid = threadIdx.x, D = p.D, M = p.M;
    const int64_t npairs = (int64_t)D * (D + 1) / 2;
    for (int64_t row = blockIdx.x; row < p.N; row += gridDim.x) {
        const int64_t tile = row >> 4;
        const int n = (int)(row & 15);
        const double* ktile = p.kstar + (size_t)tile * p.kblk * 64 + n * 4;
        const double mu = p.mu_tmp[row];
        for (int64_t q = tid; q < npairs; q += kGenThreads) {
            // q -> (d, e), d <= e, row-major over the upper triangle
            int d = (int)((2.0 * D + 1.0 - sqrt((2.0 * D + 1.0) * (2.0 * D + 1.0) - 8.0 * (double)q)) * 0.5);
            while ((int64_t)d * D - (int64_t)d * (d - 1) / 2 > q) --d;
            while ((int64_t)(d + 1) * D - (int64_t)(d + 1) * d / 2 <= q) ++d;
            const int e = d + (int)(q - ((int64_t)d * D - (int64_t)d * (d - 1) / 2));
            const double swd = __ldg(p.sqw + d), swe = __ldg(p.sqw + e);
            const double td = __ldg(p.testing + row * D + d) * swd, te = __ldg(p.testing + row * D + e) * swe;
            double h = 0.0;
            for (int j = 0; j < M; ++j) {
                const double c = ktile[(size_t)(j >> 2) * 64 + (j & 3)] * __ldg(p.alpha + j);
                const double ud = __ldg(p.xs + (size_t)j * D + d) - td, ue = __ldg(p.xs + (size_t)j * D + e) - te;
                h = fma(c * ud, ue, h);
            }
            h *= swd * swe;
            if (d == e) h -= swd * swd * mu;
            p.hess[row * p.ld_hess + (int64_t)d * D + e] = h;
            p.hess[row * p.ld_hess + (int64_t)e * D + d] = h;
        }
    }
}

}  // namespace gpe
