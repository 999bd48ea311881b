// FP64 pipe micro-benchmarks for sm_100a (B200).
//
// MEASURED_PEAKS.json (driver-written) holds HBM GB/s and bf16 TFLOP/s but no FP64 number, and the
// GP predict path is FP64-arithmetic bound.  These kernels measure, on the GPU the
// bench runs on, the sustained throughput of
//   * the FP64 FMA pipe                 (DFMA,  register operands only)
//   * the FP64 tensor path              (DMMA.8x8x4 = mma.sync.m8n8k4.f64, register operands only)
//   * both issued from the same SM      (do they add up, i.e. are they separate pipes?)
//   * the FP64 exp used for K*          (our gpe_exp_neg and CUDA's exp())
// The roofline denominator reported by bench.py is max(DFMA, DMMA) from this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "gpe_math.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_dfma(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chain alive
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kThreads) k_dmma(double* out, int iters, double a, double b) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// Half of the warps of every CTA run DFMA chains, the other half DMMA chains.
// fma_iters / mma_iters let the caller balance the two so both finish together.
__global__ void __launch_bounds__(kThreads) k_mixed(double* out, int fma_iters, int mma_iters, double a, double b) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < fma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[i];
    } else {
        double c[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
        for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    }
    if (s == 12345.678) out[0] = s;
}

template <int kWhich>
__global__ void __launch_bounds__(kThreads) k_exp(double* out, int iters, double x0, double dx) {
    double acc[4] = {0, 0, 0, 0};
    double x = x0 - threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double xi = x - i * 0.25;
            acc[i] += (kWhich == 0) ? gpe::exp_neg(xi) : exp(xi);
        }
        x -= dx;
        if (x < -40.0) x = x0;
    }
    double s = acc[0] + acc[1] + acc[2] + acc[3];
    if (s == 12345.678) out[0] = s;
}

__global__ void k_clock(long long* out) {
    // SM clock: cycles elapsed per nanosecond of globaltimer over a ~2 ms spin under load.
    unsigned long long t0, t1;
    long long c0 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    double x = threadIdx.x;
    do {
        for (int i = 0; i < 1024; ++i) x = fma(x, 1.0000001, 1e-9);
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    } while (t1 - t0 < 2000000ull);
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(t1 - t0); }
    if (x == 12345.678) out[2] = (long long)x;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace

// out[0]=DFMA TFLOP/s, out[1]=DMMA TFLOP/s, out[2]=mixed total TFLOP/s, out[3]=mixed DFMA part,
// out[4]=mixed DMMA part, out[5]=gpe exp Gexp/s, out[6]=CUDA exp Gexp/s, out[7]=SM MHz under FP64 load,
// out[8]=#SMs.  Returns 0 on success, a cudaError_t otherwise.
// Developer aid (not in the public header): y[i] = exp(x[i]) on the device with the kernels' own routines, host
// pointers; which = 0: exp_neg_tab (table-driven, what the FP64 predict kernels use), 1: exp_neg (polynomial only, used
// by the training kernel).  tests/test_gpu_parity.py::test_device_exp_accuracy holds both against mpmath.
__global__ void k_exp_eval(const double* x, double* y, long long n, int which) {
    __shared__ double tab[64];
    gpe::exp_tab_load(tab, threadIdx.x);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = which ? gpe::exp_neg(x[i]) : gpe::exp_neg_tab(x[i], tab);
}

extern "C" int gpe_debug_exp(const double* x, double* y, long long n, int which) {
    if (n <= 0) return 0;
    double *dx = nullptr, *dy = nullptr;
    cudaError_t e = cudaMalloc((void**)&dx, n * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dy, n * 8);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_exp_eval<<<(unsigned)((n + 255) / 256), 256>>>(dx, dy, n, which);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(y, dy, n * 8, cudaMemcpyDeviceToHost);
    if (dx) cudaFree(dx);
    if (dy) cudaFree(dy);
    return e == cudaSuccess ? 0 : -2;
}

extern "C" int gpe_measure_fp64_peaks(int device, double* out9) {
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess) return (int)err;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int sms = prop.multiProcessorCount;
    double* d_out;
    cudaMalloc(&d_out, 64);
    long long* d_clk;
    cudaMalloc(&d_clk, 64);
    const int ctas = sms * 8;  // 8 x 256 threads = 64 warps/SM resident
    const int iters = 4096;
    const double threads = (double)ctas * kThreads;

    float ms = time_ms([&] { k_dfma<<<ctas, kThreads>>>(d_out, iters, 1.0000001, 1e-9); }, 5);
    out9[0] = threads * iters * 16 * 2 / (ms * 1e-3) / 1e12;

    const double warps = threads / 32;
    ms = time_ms([&] { k_dmma<<<ctas, kThreads>>>(d_out, iters, 1.0000001, 1e-9); }, 5);
    out9[1] = warps * iters * 16 * (8 * 8 * 4 * 2) / (ms * 1e-3) / 1e12;

    // mixed: same number of FP64 flops on each side per warp-iteration?  DFMA warp-instr = 64 flop,
    // DMMA warp-instr = 512 flop.  Give each side the time it needs alone: fma_iters*16 DFMA vs
    // mma_iters*16 DMMA; choose iters so that alone they take the same time.
    {
        const double t_fma_per_iter = 1.0 / (out9[0] / 64.0);   // relative time per warp-instr
        const double t_mma_per_iter = 1.0 / (out9[1] / 512.0);
        int mma_iters = iters;
        int fma_iters = (int)(iters * t_mma_per_iter / t_fma_per_iter);
        if (fma_iters < 1) fma_iters = 1;
        ms = time_ms([&] { k_mixed<<<ctas, kThreads>>>(d_out, fma_iters, mma_iters, 1.0000001, 1e-9); }, 5);
        const double f_fma = (warps / 2) * fma_iters * 16.0 * 64.0;
        const double f_mma = (warps / 2) * mma_iters * 16.0 * 512.0;
        out9[2] = (f_fma + f_mma) / (ms * 1e-3) / 1e12;
        out9[3] = f_fma / (ms * 1e-3) / 1e12;
        out9[4] = f_mma / (ms * 1e-3) / 1e12;
    }

    const int eiters = 2048;
    ms = time_ms([&] { k_exp<0><<<ctas, kThreads>>>(d_out, eiters, -0.01, 0.37); }, 5);
    out9[5] = threads * eiters * 4 / (ms * 1e-3) / 1e9;
    ms = time_ms([&] { k_exp<1><<<ctas, kThreads>>>(d_out, eiters, -0.01, 0.37); }, 5);
    out9[6] = threads * eiters * 4 / (ms * 1e-3) / 1e9;

    k_clock<<<sms, 1024>>>(d_clk);
    long long h[2] = {0, 1};
    cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
    out9[7] = (double)h[0] / (double)h[1] * 1e3;
    out9[8] = sms;
    err = cudaDeviceSynchronize();
    cudaFree(d_out);
    cudaFree(d_clk);
    return (int)err;
}

#ifdef GPE_PEAKS_MAIN
int main() {
    double o[9];
    int rc = gpe_measure_fp64_peaks(0, o);
    printf("{\"rc\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"mixed_tflops\": %.3f, "
           "\"mixed_dfma\": %.3f, \"mixed_dmma\": %.3f, \"gpe_exp_gps\": %.2f, \"cuda_exp_gps\": %.2f, "
           "\"sm_mhz_fp64_load\": %.1f, \"sms\": %.0f}\n",
           rc, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8]);
    return rc;
}
#endif
