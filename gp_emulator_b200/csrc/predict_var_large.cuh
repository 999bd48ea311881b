// Variance contraction for 1024 < M <= GPE_MAX_TRAIN (16384) training points (FP64, sm_100a).
//
// The fused kernel (predict_full.cuh) keeps the whole K* tile in shared memory and the whole TN x Mp accumulator
// tile in registers; neither fits beyond M = 1024.  Here the work is split in two launches per batch of points:
//   k_predict_mean2<DP, true>  (predict_mean.cuh) forms K*, the mean and the gradient and writes K* to a scratch
//                              buffer in HBM / L2, tiled as [tile of 16 points][k-block][16 rows][4];
//   k_var_large                (this file) computes var_n = b - b^2 sum_j (sum_i K*_ni invQ_ji) K*_nj for 16 points
//                              per persistent CTA in column passes of 1024: per pass the 16 x 1024 accumulator tile
//                              lives in registers (8 warps x 16 x 128), and for every k-block one TMA stage brings
//                              both operands -- the 512-byte K* slab (A) and the 32 KB invQ slab (B, s_tiled layout
//                              of predict_full.cuh) -- through an mbarrier ring; the pass epilogue multiplies by the
//                              K* columns of the pass (read back from the scratch, C-fragment layout).
// Reference: GaussianProcess.py:240 (the same formula; the reference has no limit on M).
// Cost model: invQ is re-streamed once per 16-point tile (8 M^2 bytes, L2-resident: ncu shows a 95 % L2 hit rate and
// 22 % of the L2 throughput at M = 2048), the FP64 tensor sub-pipe is 80 % active -- the kernel is tensor-bound
// like the fused one; the K* scratch adds 8 M bytes per point of HBM write + read (profiles/r01_ncu_var_large_summary.md).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_math.cuh"
#include "gpe_ptx.cuh"

namespace gpe {

constexpr int kVlTN = 16;        // points per tile
constexpr int kVlWarps = 8;
constexpr int kVlPass = 1024;    // columns per pass
constexpr uint32_t kVlStageBytes = 512u + (uint32_t)kVlPass * 32u;

struct VarLargeParams {
    const double* kstar;    // [ceil(N/16)][kblk][16][4]
    const double* s_tiled;  // [kblk][Mp][4]
    double* var;
    int64_t ld_var;
    int64_t N;
    int Mp;                 // padded M (multiple of 64)
    int kblk;               // ceil(M / 4)
    int npass;              // ceil(Mp / 1024)
    int nstage;             // ring depth (<= 8)
    double b;
};

#ifdef GPE_VAR_LARGE_IMPL   // the kernel is not a template: only predict_var_large.cu compiles its body
__global__ void __launch_bounds__(kVlWarps * 32, 1) k_var_large(const VarLargeParams p) {
    constexpr int NW = kVlWarps, NT = 16, MT = 2;
    extern __shared__ __align__(128) unsigned char smem_vl[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_vl);
    uint64_t* bar_empty = bar_full + 8;
    double* vred = reinterpret_cast<double*>(smem_vl + 128);       // [NW][16]
    unsigned char* ring = smem_vl + 128 + NW * kVlTN * 8;            // nstage x kVlStageBytes (128-byte aligned)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nstage = p.nstage;
    const int lag = (nstage >= 3) ? 2 : 1;
    smem_guard(128u + NW * kVlTN * 8u + (uint32_t)nstage * kVlStageBytes);
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], NW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int64_t ntiles = (p.N + kVlTN - 1) / kVlTN;
    const int nit = p.npass * p.kblk;   // ring iterations per tile: pass-major, k-block minor
    int cs = 0;
    uint32_t cpar = 0;
    const double* ktile = nullptr;
    auto issue = [&](int stage, int g) {
        const int pass = g / p.kblk, kb = g - pass * p.kblk;
        const int pcw = min(kVlPass, p.Mp - pass * kVlPass);
        unsigned char* dst = ring + (size_t)stage * kVlStageBytes;
        mbar_arrive_expect_tx(&bar_full[stage], 512u + (uint32_t)pcw * 32u);
        tma_bulk_g2s(dst, ktile + (size_t)kb * 64, 512u, &bar_full[stage]);
        tma_bulk_g2s(dst + 512, p.s_tiled + ((size_t)kb * p.Mp + (size_t)pass * kVlPass) * 4, (uint32_t)pcw * 32u,
                     &bar_full[stage]);
    };
    const int a_off = (lane >> 2) * 4 + (lane & 3);
    const int b_off = (warp * 8 + (lane >> 2)) * 4 + (lane & 3);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        ktile = p.kstar + (size_t)tile * p.kblk * 64;
        if (tid == 0) {   // every stage was released before the end-of-tile barrier
            int s = cs;
            const int burst = min(nstage, nit);
            for (int i = 0; i < burst; ++i) {
                issue(s, i);
                if (++s == nstage) s = 0;
            }
        }
        double vs[MT] = {0.0, 0.0};
        int it = 0;
        for (int pass = 0; pass < p.npass; ++pass) {
            const int pcw = min(kVlPass, p.Mp - pass * kVlPass);
            double acc[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int kb = 0; kb < p.kblk; ++kb, ++it) {
                if (it >= lag && lane == 0 && warp == (it & (NW - 1)) && it - lag + nstage < nit) {
                    const int ps = (cs >= lag) ? cs - lag : cs - lag + nstage;
                    const uint32_t ppar = (cs >= lag) ? cpar : (cpar ^ 1);
                    mbar_wait(&bar_empty[ps], ppar);
                    issue(ps, it - lag + nstage);
                }
                mbar_wait(&bar_full[cs], cpar);
                const double* as = reinterpret_cast<const double*>(ring + (size_t)cs * kVlStageBytes);
                const double* bs = as + 64 + b_off;
                const double a0 = as[a_off], a1 = as[32 + a_off];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if ((warp + NW * j) * 8 >= pcw) break;   // ragged last pass (real exit, warp-uniform)
                    const double bf = bs[j * (NW * 32)];
                    dmma_m8n8k4(acc[0][j][0], acc[0][j][1], a0, bf);
                    dmma_m8n8k4(acc[1][j][0], acc[1][j][1], a1, bf);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[cs]);
                if (++cs == nstage) { cs = 0; cpar ^= 1; }
            }
            // pass epilogue: sum_j G_nj K*_nj over this warp's columns of the pass
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if ((warp + NW * j) * 8 >= pcw) break;
                const int jc = pass * kVlPass + (warp + NW * j) * 8 + 2 * (lane & 3);
                const double* kp = ktile + (size_t)(jc >> 2) * 64 + (lane >> 2) * 4 + (jc & 3);
                const double2 k0 = *reinterpret_cast<const double2*>(kp);
                const double2 k1 = *reinterpret_cast<const double2*>(kp + 32);
                vs[0] = fma(acc[0][j][0], k0.x, vs[0]);
                vs[0] = fma(acc[0][j][1], k0.y, vs[0]);
                vs[1] = fma(acc[1][j][0], k1.x, vs[1]);
                vs[1] = fma(acc[1][j][1], k1.y, vs[1]);
            }
        }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            vs[i] += __shfl_xor_sync(0xffffffffu, vs[i], 1);
            vs[i] += __shfl_xor_sync(0xffffffffu, vs[i], 2);
            if ((lane & 3) == 0) vred[warp * kVlTN + i * 8 + (lane >> 2)] = vs[i];
        }
        __syncthreads();
        if (tid < kVlTN && tile * kVlTN + tid < p.N) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += vred[w * kVlTN + tid];
            p.var[(tile * kVlTN + tid) * p.ld_var] = p.b - p.b * p.b * v;
        }
        __syncthreads();
    }
}

#endif  // GPE_VAR_LARGE_IMPL

}  // namespace gpe
