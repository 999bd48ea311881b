// Host-side declarations of the per-DP kernel launchers (defined in predict_full_inst.cu, one object per DP).
#pragma once
#include <cuda_runtime.h>
#include "predict_full.cuh"
#include "predict_mean.cuh"
#include "predict_bank_mean.cuh"
#include "predict_tiny.cuh"
#include "predict_tf32.cuh"
#include "predict_tf32_big.cuh"
#include "predict_var_large.cuh"

namespace gpe {

#define GPE_DECL_DP(DPV)                                                                                          \
    cudaError_t launch_full_dp##DPV(int cfg, const FullParams& p, int grid, size_t smem, cudaStream_t st);         \
    cudaError_t launch_mean_dp##DPV(bool hess, const MeanParams& p, dim3 grid, size_t smem, cudaStream_t st);   \
    cudaError_t launch_bank_mean_dp##DPV(int G, bool grad, const BankMeanParams& p, dim3 grid, size_t smem, cudaStream_t st); \
    cudaError_t launch_tiny_dp##DPV(const TinyParams& p, int grid, size_t smem, cudaStream_t st);

GPE_DECL_DP(2)
GPE_DECL_DP(4)
GPE_DECL_DP(6)
GPE_DECL_DP(8)
GPE_DECL_DP(10)
GPE_DECL_DP(12)
GPE_DECL_DP(16)
GPE_DECL_DP(24)
GPE_DECL_DP(32)
#undef GPE_DECL_DP

cudaError_t launch_tf32(int DP, bool x3, const Tf32Params& p, int grid, size_t smem, cudaStream_t st);
cudaError_t launch_tf32_big(int DP, bool x3, const Tf32BigParams& p, int grid, size_t smem, cudaStream_t st);
cudaError_t launch_var_large(const VarLargeParams& p, int grid, size_t smem, cudaStream_t st);
// out (R, W) [+]= A (R, E <= 32) . basis, basis pre-tiled as [ks][Wp][4]; row r of A at (r / RD) ldn + (r % RD) ldd, element e
// at + e lde (project.cu)
cudaError_t project_rows(const double* A, int64_t R, int RD, int64_t ldn, int64_t lde, int64_t ldd, const double* b_tiled, int E,
                         int W, int Wp, double* out, int accumulate, cudaStream_t st);
// group sizes the shared-difference bank kernel (predict_bank_mean.cuh) is compiled for: the G x (DP + 1) accumulators
// of a thread (G without the gradient) have to stay in registers
inline bool bank_group_ok(int DP, int G, bool grad) {
    if (!grad) return DP <= 16 && (G == 4 || G == 5 || G == 8 || G == 10);
    if (DP <= 10) return G >= 3 && G <= 5;
    if (DP == 12) return G == 3 || G == 4;
    if (DP == 16) return G == 2 || G == 3;
    return false;
}
static const int kTfDpList[] = {4, 8, 12, 16, 32};

// padded input dimensions that have compiled kernels, ascending
static const int kDpList[] = {2, 4, 6, 8, 10, 12, 16, 24, 32};
static const int kNumDp = sizeof(kDpList) / sizeof(kDpList[0]);

}  // namespace gpe
