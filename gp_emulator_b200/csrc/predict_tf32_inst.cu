// Instantiations of the tcgen05 / TMEM single-precision predict kernel (one per float4-padded input dimension).
#include "predict_tf32.cuh"
#include "predict_tf32_big.cuh"
#include "launch.h"

namespace gpe {

template <int DP>
static cudaError_t launch_dp(const Tf32Params& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_predict_tf32<DP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kTfThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_tf32(int DP, const Tf32Params& p, int grid, size_t smem, cudaStream_t st) {
    switch (DP) {
        case 4: return launch_dp<4>(p, grid, smem, st);
        case 8: return launch_dp<8>(p, grid, smem, st);
        case 12: return launch_dp<12>(p, grid, smem, st);
        case 16: return launch_dp<16>(p, grid, smem, st);
        case 32: return launch_dp<32>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int DP>
static cudaError_t launch_big_dp(const Tf32BigParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_predict_tf32_big<DP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kTfThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_tf32_big(int DP, const Tf32BigParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (DP) {
        case 4: return launch_big_dp<4>(p, grid, smem, st);
        case 8: return launch_big_dp<8>(p, grid, smem, st);
        case 12: return launch_big_dp<12>(p, grid, smem, st);
        case 16: return launch_big_dp<16>(p, grid, smem, st);
        case 32: return launch_big_dp<32>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gpe
