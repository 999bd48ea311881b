// Device math helpers shared by the predict kernels and the peak micro-benchmarks.
#pragma once
#include <cuda_runtime.h>

namespace gpe {

// Row pitch (doubles) of the training-input chunks the FP64 predict kernels keep in shared memory.  The 8 (or 4)
// training-point lanes of a quarter-warp read consecutive rows with LDS.128, so the pitch counted in 16-byte units has to
// be odd, or rows share banks: at pitch = DP the loads were 4-way conflicted for DP = 8, 2-way for 12, 8-way for 16
// (round 2, tools/d_sweep_probe.py: D = 8 ran SLOWER than D = 10).  DP = 2, 6, 10 are conflict-free as they are.
__host__ __device__ constexpr int x_pitch(int DP) { return ((DP / 2) & 1) ? DP : DP + 2; }

// exp(x) for x <= 0 in FP64, branch-free.
//
// The squared-exponential covariance only ever needs exp of a non-positive argument
// (k = b*exp(-r2/2), reference GaussianProcess.py:234), so the overflow / large-positive paths of a
// general exp() are dropped.  Range reduction x = n*ln2 + f with |f| <= ln2/2 (magic-number rint,
// two-term Cody-Waite ln2), degree-11 polynomial from Chebyshev interpolation on [-ln2/2, ln2/2]
// (max relative error of the polynomial 1.7e-17, derived with mpmath at 60 digits), then 2^n applied
// by an integer add on the exponent field.  Results below the normal range (x < -708) flush to 0.
// Measured against mpmath: <= 1 ulp on [-708, 0].  16 FP64-pipe operations; used by the training kernel -- the predict
// kernels use the 10-operation table-driven exp_neg_tab below.
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

// Table-driven variant, what every FP64 predict kernel uses: 10 FP64-pipe operations instead of 16 (+1.8 % on the fused
// kernel, +5.8 % on the mean + gradient kernel; the same routine everywhere keeps mean and gradient independent of whether
// the variance is requested).  x = (64 n + j) ln2/64 + f, |f| <= ln2/128 (two-term Cody-Waite ln2/64 whose high part has 32
// significant bits), exp(x) = 2^n T[j] (1 + f q(f)) with T[j] = 2^(j/64) correctly rounded (mpmath) read from a 64-entry
// table in shared memory (the load overlaps the polynomial) and q = expm1(f)/f to degree 4 (truncation f^6/720 <=
// 3.5e-17).  Measured against mpmath on the GPU: <= 1.3 ulp on [-708, 0] (table rounding + final FMA).
static __device__ const double kExp2Tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951,
};

__device__ __forceinline__ void exp_tab_load(double* tab_s, int tid) {   // followed by a CTA barrier at the call site
    if (tid < 64) tab_s[tid] = kExp2Tab[tid];
}

// (Sixteen interleaved table copies -- every lane on its own bank pair, no look-up conflicts -- were measured in round 2:
// 1 % SLOWER on k_bank_mean and k_predict_mean2; the conflicts cost LSU wavefronts, which these FP64-bound loops have to spare.)
__device__ __forceinline__ double exp_neg_tab(double x, const double* tab_s) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(x, 92.33248261689366, kMagic);            // 64 / ln2
    const int n = __double2loint(t);
    const double tn = t - kMagic;
    double f = fma(tn, -0.010830424693267559, x);                  // ln2_hi / 64
    f = fma(tn, -2.9815858269852933e-12, f);                       // ln2_lo / 64
    const double T = tab_s[n & 63];
    double q = fma(0.008333333333333333, f, 0.041666666666666664);
    q = fma(q, f, 0.16666666666666666);
    q = fma(q, f, 0.5);
    q = fma(q, f, 1.0);
    const double r0 = fma(T, q * f, T);
    // flush below the normal range (x < -708, -inf included; NaN propagates) with an INTEGER range test on the high word of
    // x: a DSETP would be one more instruction on the FP64 pipe, which is what bounds every kernel that calls this
    const bool flush = ((unsigned)__double2hiint(x) - 0xC0862001u) <= (0xFFF00000u - 0xC0862001u);
    const int hi = __double2hiint(r0) + ((n >> 6) << 20);
    return __hiloint2double(flush ? 0 : hi, flush ? 0 : __double2loint(r0));
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
