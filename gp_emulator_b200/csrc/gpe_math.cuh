// Device math helpers shared by the predict kernels and the peak micro-benchmarks.
#pragma once
#include <cuda_runtime.h>

namespace gpe {

// exp(x) for x <= 0 in FP64, branch-free.
//
// The squared-exponential covariance only ever needs exp of a non-positive argument
// (k = b*exp(-r2/2), reference GaussianProcess.py:234), so the overflow / large-positive paths of a
// general exp() are dropped.  Range reduction x = n*ln2 + f with |f| <= ln2/2 (magic-number rint,
// two-term Cody-Waite ln2), degree-11 polynomial from Chebyshev interpolation on [-ln2/2, ln2/2]
// (max relative error of the polynomial 1.7e-17, derived with mpmath at 60 digits), then 2^n applied
// by an integer add on the exponent field.  Results below the normal range (x < -708) flush to 0.
// Measured against mpmath: <= 1 ulp on [-708, 0].
__device__ __forceinline__ double exp_neg(double x) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(x, 1.4426950408889634, kMagic);
    const int n = __double2loint(t);
    const double tn = t - kMagic;
    double f = fma(tn, -0.6931471805599453, x);
    f = fma(tn, -2.3190468138462996e-17, f);
    double p = 2.5110037605963777e-08;
    p = fma(p, f, 2.763263963904103e-07);
    p = fma(p, f, 2.755724091857897e-06);
    p = fma(p, f, 2.4801485482328494e-05);
    p = fma(p, f, 0.00019841269890047113);
    p = fma(p, f, 0.0013888888952314775);
    p = fma(p, f, 0.008333333333319601);
    p = fma(p, f, 0.0416666666664881);
    p = fma(p, f, 0.1666666666666668);
    p = fma(p, f, 0.5000000000000019);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double r = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return (x < -708.0) ? 0.0 : r;
}

// One FP64 tensor-core op: D(8x8) += A(8x4) * B(4x8), warp-wide (SASS: DMMA.8x8x4).
// Lane l holds A[l/4][l%4], B[k=l%4][n=l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Sum `NV` per-point values held by the 8 training-point lanes (lane bits 0..2) of two points a / b with a
// shuffle reduce-scatter: 8-lane butterflies would need 3 * 2 * NV shuffles, this needs NV + NV/2 + NV/4.
// On return lane (b2 b1 b0) holds the complete sums of point (b2 ? b : a) for value indices
//   base + i,  i < H3,  base = (b1 ? H2 : 0) + (b0 ? H3 : 0)   (indices >= the half / NV are padding).
template <int NV>
struct RS {
    static constexpr int H2 = (NV + 1) / 2;
    static constexpr int H3 = (H2 + 1) / 2;
    double r3[H3];
    __device__ __forceinline__ void run(const double (&va)[NV], const double (&vb)[NV], int lane) {
        const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
        double r1[2 * H2];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double send = b2 ? va[i] : vb[i];
            const double keep = b2 ? vb[i] : va[i];
            r1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int i = NV; i < 2 * H2; ++i) r1[i] = 0.0;
        double r2[2 * H3];
#pragma unroll
        for (int i = 0; i < H2; ++i) {
            const double send = b1 ? r1[i] : r1[H2 + i];
            const double keep = b1 ? r1[H2 + i] : r1[i];
            r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
        for (int i = H2; i < 2 * H3; ++i) r2[i] = 0.0;
#pragma unroll
        for (int i = 0; i < H3; ++i) {
            const double send = b0 ? r2[i] : r2[H3 + i];
            const double keep = b0 ? r2[H3 + i] : r2[i];
            r3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
};

}  // namespace gpe
