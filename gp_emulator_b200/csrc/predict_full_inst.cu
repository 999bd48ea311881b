// Explicit instantiations of the fused predict kernels for one padded input dimension GPE_DP.
// Compiled once per DP (-DGPE_DP=..) so the variants build in parallel; see Makefile.
#include "predict_full.cuh"
#include "predict_mean.cuh"
#include "launch.h"
#include <stdlib.h>

#ifndef GPE_DP
#error "compile with -DGPE_DP=<padded input dimension>"
#endif

#define GPE_CAT2(a, b) a##b
#define GPE_CAT(a, b) GPE_CAT2(a, b)

namespace gpe {

template <int MT, int NT, int WR, int WC, int MINB, int KB>
static cudaError_t launch_cfg(const FullParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = p.symmetric ? ((p.nt_act == NT) ? k_predict_full<MT, NT, WR, WC, GPE_DP, MINB, KB, true, true, false>
                                                : k_predict_full<MT, NT, WR, WC, GPE_DP, MINB, KB, false, true, false>)
                            : ((p.nt_act == NT) ? k_predict_full<MT, NT, WR, WC, GPE_DP, MINB, KB, true, false, false>
                                                : k_predict_full<MT, NT, WR, WC, GPE_DP, MINB, KB, false, false, false>);
#if GPE_DP <= 16
    // fused Hessian variants (gpemu.cu only asks for them when cfg <= 2 and the model is not symmetric-folded)
    if (p.hess != nullptr) {
        if (p.symmetric || MINB != 1 || WR * WC != 8) return cudaErrorInvalidValue;
        if constexpr (MINB == 1 && WR * WC == 8)
            kern = (p.nt_act == NT) ? k_predict_full<MT, NT, WR, WC, GPE_DP, MINB, KB, true, false, true>
                                    : k_predict_full<MT, NT, WR, WC, GPE_DP, MINB, KB, false, false, true>;
    }
#else
    if (p.hess != nullptr) return cudaErrorInvalidValue;
#endif
    // per function AND per device: set on every launch (microseconds) so multi-device processes stay correct
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, WR * WC * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t GPE_CAT(launch_full_dp, GPE_DP)(int cfg, const FullParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_cfg<4, 8, 2, 4, 1, 2>(p, grid, smem, st);   // TN = 64, Mp <= 256, 8 warps, 1 CTA/SM
        case 1: return launch_cfg<4, 8, 1, 8, 1, 1>(p, grid, smem, st);   // TN = 32, Mp <= 512, 8 warps, 1 CTA/SM
        case 2: return launch_cfg<2, 16, 1, 8, 1, 1>(p, grid, smem, st);  // TN = 16, Mp <= 1024, 8 warps, 1 CTA/SM
        case 3: return launch_cfg<4, 8, 1, 4, 2, 1>(p, grid, smem, st);   // TN = 32, Mp <= 256, 4 warps, 2 CTAs/SM
        case 4: return launch_cfg<4, 4, 2, 8, 1, 1>(p, grid, smem, st);   // TN = 64, Mp <= 256, 16 warps, 1 CTA/SM
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t GPE_CAT(launch_mean_dp, GPE_DP)(bool hess, const MeanParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    if (hess) {
#if GPE_DP <= 12
        auto kern = k_predict_mean<GPE_DP, true>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kMeanThreads, smem, st>>>(p);
        return cudaGetLastError();
#else
        // D > 12: row-block Hessian kernel (predict_mean.cuh::k_hessian_rows); mean / gradient, if also requested,
        // come from the plain mean kernel first
        if (p.mu != nullptr || p.deriv != nullptr) {
            MeanParams q = p;
            q.hess = nullptr;
            auto k0 = k_predict_mean<GPE_DP, false>;
            cudaError_t e0 = cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e0 != cudaSuccess) return e0;
            k0<<<grid, kMeanThreads, smem, st>>>(q);
        }
        constexpr int HR = (GPE_DP <= 16) ? 4 : 2;
        auto kern = k_hessian_rows<GPE_DP, HR>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kMeanThreads, smem, st>>>(p);
        return cudaGetLastError();
#endif
    }
    // mean + gradient: two-rows-per-thread kernel (k_predict_mean2); GPE_MEAN_V1=1 selects the first version
    static const bool v1 = getenv("GPE_MEAN_V1") != nullptr;
    auto kern = p.kstar != nullptr ? k_predict_mean2<GPE_DP, true>
                                   : (v1 ? k_predict_mean<GPE_DP, false> : k_predict_mean2<GPE_DP, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kMeanThreads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace gpe
