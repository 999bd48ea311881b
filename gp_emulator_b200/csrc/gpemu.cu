// libgpemu.so -- C ABI implementation (see include/gpemu.h for the contract and reference citations).
//
// Host responsibilities: validate, lay the trained model out for the kernels (done once per model, the
// reference re-uploads and re-transposes it for every 2e5-point block: gp_emulator/gpu/predict.cu:11-34,
// _gpu_predict.cpp:129-132), pick the kernel configuration, launch, and -- for host-resident callers --
// stream test points through a two-slot pinned pipeline (replaces the Python chunk loop of
// GaussianProcess.gpu_predict / get_gpu_block, gp_emulator/GaussianProcess.py:253-323).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpemu.h"
#include "launch.h"
#include "gpe_math.cuh"
#include "host_common.h"

using namespace gpe;

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
long long* g_trace = nullptr;  // dev aid: device buffer for per-tile phase timestamps (gpe_debug_trace)

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(GPE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

constexpr uint32_t kSmemMax = 232448;  // 227 KB opt-in limit per CTA on sm_100
constexpr int64_t kPipeChunk = 1 << 18;  // points per host-streaming chunk
constexpr int64_t kZeroCopyMax = 16384;  // host calls of up to this many points run on mapped page-locked buffers
                                         // (tools/zero_copy_probe.py: 1000 points 71 -> 58 us, 16000 points 470 -> 300 us)

inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

constexpr int kHessFusedMaxDp = 16;   // predict_full_inst.cu instantiates the HESS variants up to this DP

struct FullPlan {
    bool valid = false;
    int cfg = 0, TN = 0, WC = 0, GH = 1, alias_x = 0, ctas_per_sm = 1;
    int Mp = 0, nt_act = 0, kblk = 0, kbps = 0, nit = 0, nstage = 0, JC = 0, nchunks = 0;
    uint32_t off_bar = 0, off_sqw = 0, off_ks = 0, off_bst = 0, off_xc = 0, off_ts = 0, off_pa = 0, off_vred = 0;
    uint32_t stage_bytes = 0, smem = 0, ts_bytes = 0;
    // fused Hessian (phase C of the fused kernel): P operand width, k-blocks per ring stage, iterations
    uint32_t off_hts = 0;
    int NC = 0, kbh = 0, nit_h = 0;
};

struct Tf32Plan {
    bool valid = false;
    int DP = 0, Mp = 0, nslab = 0;
    bool big = false;      // 256 < M <= 1024: column passes + A ring (predict_tf32_big.cuh)
    int pass_cols = 0;
    uint32_t off_bar = 0, off_tmem = 0, off_a = 0, off_b = 0, off_x = 0, off_out = 0, off_vred = 0, bstage_bytes = 0, smem = 0;
};

struct MeanPlan {
    int JC = 0, nchunks = 0;
    uint32_t off_xc = 0, off_ts = 0, off_out = 0, smem = 0, smem_hess = 0;
};

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    double* d_in = nullptr;
    double* d_out = nullptr;   // mu | var | deriv | hess, sized for the flags of the first use
    double* h_in = nullptr;    // pinned staging (pageable callers only)
    double* h_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0, h_in_cap = 0, h_out_cap = 0;
    // pending copy-out for the staging path
    bool pending = false;
    int64_t pend_n0 = 0, pend_n = 0;
};

}  // namespace

struct gpe_model {
    int device = 0, M = 0, D = 0, DP = 0, sms = 0;
    double b = 0;
    double sqrt_w[32];
    bool has_invQ = false;
    bool symmetric = false;      // GPE_OPT_SYMMETRIC_VARIANCE: s_tiled holds the upper-triangular fold of invQ
    FullPlan full;
    FullPlan full_small;         // 16-point tiles for small batches (valid only if it can share the main plan's operands)
    MeanPlan mean;
    double* d_xchunks_full = nullptr;
    double* d_stiled = nullptr;
    double* d_ptiled = nullptr;  // Hessian operand of the fused kernel (null: direct Hessian kernel only)
    // 1024 < M <= GPE_MAX_TRAIN: two-launch variance path (predict_var_large.cuh); d_stiled then has Mp / 4 k-blocks
    bool large_valid = false;
    int large_Mp = 0, large_kblk = 0, large_nstage = 0;
    int64_t large_chunk = 0;             // points per sub-batch = capacity of the K* scratch
    double* d_kscratch = nullptr;        // [large_chunk / 16][large_kblk][16][4], pads stay zero
    cudaEvent_t scratch_free = nullptr;  // the scratch is shared by every stream that predicts with this model
    double centre[32];
    bool hess_fused_ok = false;
    double* d_xchunks_mean = nullptr;
    Tf32Plan tf, tfx;            // single-precision tcgen05 paths: fast (1 x TF32) and precise (3 x TF32); lazy
    float* d_xa_f32 = nullptr;
    uint32_t* d_bslabs = nullptr;
    uint32_t* d_bslabs_lo = nullptr;
    std::vector<double> h_inputs, h_invQt, h_invQ;  // host copy of the model for the lazy FP32 packing
    double h_expx[33];
    std::mutex host_mu;          // host-pointer calls on one model share its slots (and the lazy FP32 packing): one at a time
    Slot slots[3];               // host-streaming pipeline: the direct (pinned caller) path uses two, the staged path three
};

struct gpe_multi {
    std::vector<gpe_model*> models;   // one resident copy of the GP per device
};

struct gpe_bank {
    int device = 0, E = 0, M = 0, D = 0, W = 0;
    std::vector<gpe_model*> models;
    MeanBankEntry* d_entries = nullptr;  // per-emulator phase-A data for the one-launch bank mean / Hessian kernel
    double* d_basis = nullptr;   // basis pre-tiled as [ceil(E/4)][Wp][4], Wp = W rounded up to 256
    int Wp = 0;
    // gpe_bank_forward (host in, host out, one synchronisation): stream + staging buffers, grown on demand
    std::mutex fwd_mu;
    cudaStream_t fwd_st = nullptr;
    double *fwd_h = nullptr, *fwd_d = nullptr;
    size_t fwd_h_cap = 0, fwd_d_cap = 0;
    // gpe_bank_cost: per-chunk mu / deriv scratch, shared by every stream that reduces with this bank
    std::mutex cost_mu;
    double* cost_d = nullptr;
    size_t cost_cap = 0;
    cudaEvent_t cost_free = nullptr;
};

namespace {

int pick_dp(int D) {
    for (int i = 0; i < kNumDp; ++i)
        if (kDpList[i] >= D) return kDpList[i];
    return -1;
}

cudaError_t launch_full(int DP, int cfg, const FullParams& p, int grid, size_t smem, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (DP) {
        case 2: return launch_full_dp2(cfg, p, grid, smem, st);
        case 4: return launch_full_dp4(cfg, p, grid, smem, st);
        case 6: return launch_full_dp6(cfg, p, grid, smem, st);
        case 8: return launch_full_dp8(cfg, p, grid, smem, st);
        case 10: return launch_full_dp10(cfg, p, grid, smem, st);
        case 12: return launch_full_dp12(cfg, p, grid, smem, st);
        case 16: return launch_full_dp16(cfg, p, grid, smem, st);
        case 24: return launch_full_dp24(cfg, p, grid, smem, st);
        case 32: return launch_full_dp32(cfg, p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_mean(int DP, bool hess, const MeanParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (DP) {
        case 2: return launch_mean_dp2(hess, p, grid, smem, st);
        case 4: return launch_mean_dp4(hess, p, grid, smem, st);
        case 6: return launch_mean_dp6(hess, p, grid, smem, st);
        case 8: return launch_mean_dp8(hess, p, grid, smem, st);
        case 10: return launch_mean_dp10(hess, p, grid, smem, st);
        case 12: return launch_mean_dp12(hess, p, grid, smem, st);
        case 16: return launch_mean_dp16(hess, p, grid, smem, st);
        case 24: return launch_mean_dp24(hess, p, grid, smem, st);
        case 32: return launch_mean_dp32(hess, p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

// Decide tile shape, pipeline depth and the shared-memory carve-up of the fused kernel for (M, D).
//   cfg 3: Mp <= 256 -> TN = 32, 4 warps, TWO CTAs per SM: DFMA (phase A) and DMMA (phase B) share one FP64 pipe,
//          so two co-resident CTAs whose phases drift apart keep it busy (ncu: 77% -> see profiles/).  To fit
//          2 x ~100 KB the phase-A chunk buffer overlays the B-operand ring (alias_x).
//   cfg 0: same Mp range, TN = 64, 8 warps, 1 CTA/SM (kept selectable with GPE_FULL_CFG=0 for comparison)
//   cfg 1 / 2: Mp <= 512 / 1024, TN = 32 / 16, 8 warps, 1 CTA/SM
// small_Mp > 0: the low-latency plan for small batches -- 16-point tiles (cfg 2) on the padded width of the main
// plan, so that both share s_tiled / xchunks.
FullPlan plan_full(int M, int D, int DP, int small_Mp = 0) {
    FullPlan f;
    const int m32 = (M + 31) / 32 * 32, m64 = (M + 63) / 64 * 64;
    int force = -1;
    if (const char* e = getenv("GPE_FULL_CFG")) force = atoi(e);
    // the kernels also hold static shared memory (exp table 512 B; HESS variants a 1 KB index table + 256 B): leave room
    uint32_t smem_cap = kSmemMax - 2048;
    if (small_Mp > 0) {
        f.cfg = 2; f.TN = 16; f.WC = 8; f.GH = 4; f.Mp = small_Mp; f.nt_act = small_Mp / 64;
    } else if (m32 <= 256) {
        f.Mp = m32; f.nt_act = m32 / 32;
        if (force == 3) { f.cfg = 3; f.TN = 32; f.WC = 4; f.GH = 1; f.alias_x = 1; smem_cap = (233472 - 2 * 1024) / 2; }
        else if (force == 4) { f.cfg = 4; f.TN = 64; f.WC = 8; f.GH = 2; f.Mp = m64; f.nt_act = m64 / 64; }
        else { f.cfg = 0; f.TN = 64; f.WC = 4; f.GH = 1; }
    } else if (m64 <= 512) {
        f.cfg = 1; f.TN = 32; f.WC = 8; f.GH = 2; f.Mp = m64; f.nt_act = m64 / 64;
    } else if (m64 <= 1024) {
        f.cfg = 2; f.TN = 16; f.WC = 8; f.GH = 4; f.Mp = m64; f.nt_act = m64 / 64;
    } else {
        return f;  // invalid: variance contraction unsupported for this M
    }
    f.ctas_per_sm = (f.cfg == 3) ? 2 : 1;
    f.kblk = (M + 3) / 4;
    f.kbps = (f.cfg == 0) ? 2 : 1;  // k-blocks (32 * Mp bytes each) per ring stage == template KB of the cfg
    f.nit = (f.kblk + f.kbps - 1) / f.kbps;
    f.stage_bytes = (uint32_t)f.kbps * f.Mp * 32u;

    uint32_t off = 0;
    f.off_bar = off; off += 192;
    f.off_sqw = off; off += 256;
    f.off_ks = off; off += (uint32_t)f.TN * (f.Mp + 4) * 8u;
    f.ts_bytes = align_up((uint32_t)f.TN * (D + 1) * 8u, 16);  // x2: current rows / output staging + prefetch
    f.off_ts = off; off += 2 * f.ts_bytes;
    f.off_pa = off; off += (f.GH > 1) ? align_up((uint32_t)f.GH * f.TN * (D + 1) * 8u, 16) : 0;
    f.off_vred = off; off += (uint32_t)f.WC * f.TN * 8u;
    if (DP <= kHessFusedMaxDp && f.cfg <= 2) {   // room for the centred test rows of the fused Hessian
        f.off_hts = off; off += align_up((uint32_t)f.TN * D * 8u, 16);
        f.NC = (DP * (DP + 1) / 2 + 7) / 8 * 8;
        f.kbh = (int)(f.stage_bytes / ((uint32_t)f.NC * 32u));
        f.nit_h = f.kbh > 0 ? (f.kblk + f.kbh - 1) / f.kbh : 0;
    }
    off = align_up(off, 128);
    const uint32_t fixed = off;
    const int m4 = (M + 3) / 4 * 4;
    const uint32_t row = (uint32_t)(DP + 1) * 8u;
    if (f.alias_x) {
        // ring and chunk buffer share [fixed, fixed + max(ring, chunk))
        if (fixed + 2 * f.stage_bytes > smem_cap) return f;
        f.nstage = (fixed + 4 * f.stage_bytes <= smem_cap) ? 4 : 2;
        const int jc_max = (int)((smem_cap - fixed) / row) / 4 * 4;
        if (jc_max < 32) return f;
        f.JC = std::min(jc_max, m4); f.nchunks = (M + f.JC - 1) / f.JC;
        f.off_bst = fixed; f.off_xc = fixed;
        f.smem = fixed + std::max((uint32_t)f.nstage * f.stage_bytes, (uint32_t)f.JC * row);
        f.valid = f.smem <= smem_cap;
        return f;
    }
    // resident training set if it fits beside a ring of >= 3 stages, else chunks of <= 256 points; the ring then
    // takes every stage that still fits (at most 8: the barrier arrays)
    {
        const uint32_t avail = smem_cap - fixed;
        const uint32_t resident = (uint32_t)m4 * row;
        if (avail >= resident + 3 * f.stage_bytes) {
            f.JC = m4; f.nchunks = 1;
        } else {
            if (avail < 2 * f.stage_bytes + 64 * row) return f;
            const int jc_max = (int)((avail - 2 * f.stage_bytes) / row) / 4 * 4;
            f.JC = std::min(std::min(jc_max, 256), m4); f.nchunks = (M + f.JC - 1) / f.JC;
        }
        f.nstage = (int)std::min<uint32_t>(8, (avail - (uint32_t)f.JC * row) / f.stage_bytes);
        if (f.nstage < 2) return f;
    }
    if (f.nstage == 0) return f;
    f.off_bst = fixed;
    f.off_xc = fixed + (uint32_t)f.nstage * f.stage_bytes;
    f.smem = f.off_xc + (uint32_t)f.JC * row;
    f.valid = f.smem <= smem_cap;
    return f;
}

MeanPlan plan_mean(int M, int D, int DP) {
    MeanPlan m;
    const int m4 = (M + 3) / 4 * 4;
    m.JC = std::min(m4, 256);
    m.nchunks = (M + m.JC - 1) / m.JC;
    uint32_t off = 0;
    m.off_xc = off; off += align_up((uint32_t)m.JC * (DP + 1) * 8u, 16);
    m.off_ts = off; off += align_up((uint32_t)kMeanTN * (D + 1) * 8u, 16);
    m.off_out = off;
    m.smem = off;
    // Hessian staging: [TN][D][D] for the triangular kernel (D <= 12), [TN][HR][D] for the row-block kernel
    m.smem_hess = off + (uint32_t)kMeanTN * D * 8u * (uint32_t)(DP <= 12 ? D : (DP <= 16 ? 4 : 2));
    return m;
}

// [nchunks][ JC*DP sqrt(w)-scaled inputs | JC b*alpha ], zero padded
std::vector<double> pack_xchunks(int M, int D, int DP, int JC, int nchunks, const double* inputs,
                                 const double* sqrt_w, const double* invQt, double b) {
    std::vector<double> buf((size_t)nchunks * JC * (DP + 1), 0.0);
    for (int j = 0; j < M; ++j) {
        const int c = j / JC, jl = j - c * JC;
        double* base = buf.data() + (size_t)c * JC * (DP + 1);
        for (int d = 0; d < D; ++d) base[(size_t)jl * DP + d] = sqrt_w[d] * inputs[(size_t)j * D + d];
        base[(size_t)JC * DP + jl] = b * invQt[j];
    }
    return buf;
}

int check_device(int device, int* sms) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(GPE_ERR_NO_DEVICE, "no CUDA device available (%s); libgpemu has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(GPE_ERR_INVALID, "device %d out of range [0, %d)", device, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GPE_ERR_NO_DEVICE, "device %d is sm_%d%d; libgpemu kernels are built for sm_100a only", device,
                    prop.major, prop.minor);
    *sms = prop.multiProcessorCount;
    return GPE_OK;
}

void free_slot(Slot& s) {
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.done) cudaEventDestroy(s.done);
    if (s.st) cudaStreamDestroy(s.st);
    s = Slot();
}

// Operand of the fused Hessian (predict_full.cuh, phase C): p_tiled[kb][col][c] = b alpha_j x'_jd x'_je for
// j = 4 kb + c, col = index of (d <= e) in the row-major upper triangle, x' = sqrt(w) x - centre (mid-range).
// The expansion sum k a (x'_d - t'_d)(x'_e - t'_e) = S2 - t'_d g_e - t'_e g_d - t'_d t'_e S0 loses about
// log10(max |x'|^2) digits to cancellation, so it is only enabled while max |x'|^2 <= kHessFusedMaxR2
// (training inputs within ~45 length scales of the centre: error amplification <= 2e3 * 2^-53 ~ 2e-13);
// beyond that, and for the shapes the fused kernel does not cover, the direct kernel runs.
constexpr double kHessFusedMaxR2 = 2000.0;

int build_hessian_operand(gpe_model* m, const double* inputs, const double* invQt) {
    const FullPlan& f = m->full;
    const int M = m->M, D = m->D;
    memset(m->centre, 0, sizeof(m->centre));
    m->hess_fused_ok = false;
    if (getenv("GPE_HESS_DIRECT") != nullptr) return GPE_OK;   // dev aid: always use the direct Hessian kernels
    if (!f.valid || f.kbh < 1 || m->symmetric || f.NC > M || f.NC > f.Mp) return GPE_OK;
    double r2max = 0.0;
    for (int d = 0; d < D; ++d) {
        double lo = inputs[d], hi = inputs[d];
        for (int j = 1; j < M; ++j) { lo = std::min(lo, inputs[(size_t)j * D + d]); hi = std::max(hi, inputs[(size_t)j * D + d]); }
        m->centre[d] = m->sqrt_w[d] * 0.5 * (lo + hi);
        const double r = m->sqrt_w[d] * 0.5 * (hi - lo);
        r2max = std::max(r2max, r * r);
    }
    if (!(r2max <= kHessFusedMaxR2)) return GPE_OK;
    std::vector<double> pt((size_t)f.kblk * f.NC * 4, 0.0);
    std::vector<double> xp(D);
    for (int j = 0; j < M; ++j) {
        for (int d = 0; d < D; ++d) xp[d] = m->sqrt_w[d] * inputs[(size_t)j * D + d] - m->centre[d];
        const double ba = m->b * invQt[j];
        int col = 0;
        for (int d = 0; d < D; ++d)
            for (int e = d; e < D; ++e, ++col) pt[((size_t)(j >> 2) * f.NC + col) * 4 + (j & 3)] = ba * xp[d] * xp[e];
    }
    cudaError_t e = cudaMalloc((void**)&m->d_ptiled, pt.size() * 8);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_ptiled, pt.data(), pt.size() * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(GPE_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(e));
    m->hess_fused_ok = true;
    return GPE_OK;
}

// Launch the kernels for device-resident data on `st`.  Output strides allow bank (point-major) layouts.
int predict_device(gpe_model* m, const double* testing, int64_t N, double* mu, double* var, double* deriv,
                   double* hess, int64_t ld_mu, int64_t ld_var, int64_t ld_deriv, int64_t ld_hess,
                   cudaStream_t st, int64_t call_N = -1) {
    // call_N: size of the API call this launch is a chunk of (the host pipeline); plan choices that change the
    // summation order depend on it, not on the chunk, so that one call is internally consistent
    if (call_N < 0) call_N = N;
    if (N == 0) return GPE_OK;
    bool mean_done = false;
    static const bool force_full = getenv("GPE_FORCE_FULL") != nullptr;  // dev aid: time phase A alone
    // the Hessian rides on the fused kernel when the model qualifies (build_hessian_operand); without a variance
    // request that only pays for the 64-point tiles of cfg 0 (measured: 4.6e8 vs 3.6e8 points/s at M = 250, but
    // 0.98e8 vs 1.10e8 at M = 1000), so larger M keeps the direct kernel for Hessian-only calls
    const bool fuse_hess = hess != nullptr && m->hess_fused_ok && (var != nullptr || m->full.cfg == 0);
    if (var != nullptr && m->has_invQ && !m->full.valid && m->large_valid) {
        // 1024 < M <= GPE_MAX_TRAIN: per sub-batch, K* + mean + gradient (k_predict_mean2<DP, true>) into the scratch,
        // then the column-pass contraction (k_var_large).  The scratch is per model: streams take turns.
        CUDA_TRY(cudaStreamWaitEvent(st, m->scratch_free, 0));
        for (int64_t n0 = 0; n0 < N; n0 += m->large_chunk) {
            const int64_t n = std::min(m->large_chunk, N - n0);
            MeanParams p;
            memset(&p, 0, sizeof(p));
            p.testing = testing + n0 * m->D; p.N = n;
            p.mu = mu ? mu + n0 * ld_mu : nullptr;
            p.deriv = deriv ? deriv + n0 * ld_deriv : nullptr;
            p.ld_mu = ld_mu; p.ld_deriv = ld_deriv;
            p.xchunks = m->d_xchunks_mean; p.M = m->M; p.D = m->D; p.JC = m->mean.JC; p.nchunks = m->mean.nchunks;
            p.off_xc = m->mean.off_xc; p.off_ts = m->mean.off_ts; p.off_out = m->mean.off_out;
            p.kstar = m->d_kscratch; p.kblk = m->large_kblk;
            memcpy(p.sqrt_w, m->sqrt_w, sizeof(p.sqrt_w));
            const int64_t mtiles = (n + kMeanTN - 1) / kMeanTN;
            CUDA_TRY(launch_mean(m->DP, false, p, dim3((unsigned)std::min<int64_t>(mtiles, (int64_t)m->sms * 12)), m->mean.smem, st));
            VarLargeParams v;
            memset(&v, 0, sizeof(v));
            v.kstar = m->d_kscratch; v.s_tiled = m->d_stiled; v.var = var + n0 * ld_var; v.ld_var = ld_var; v.N = n;
            v.Mp = m->large_Mp; v.kblk = m->large_kblk; v.npass = (m->large_Mp + kVlPass - 1) / kVlPass;
            v.nstage = m->large_nstage; v.b = m->b;
            const size_t smem = 128 + kVlWarps * kVlTN * 8 + (size_t)v.nstage * kVlStageBytes;
            const int64_t vtiles = (n + kVlTN - 1) / kVlTN;
            g_launches.fetch_add(1, std::memory_order_relaxed);
            CUDA_TRY(launch_var_large(v, (int)std::min<int64_t>(vtiles, m->sms), smem, st));
        }
        CUDA_TRY(cudaEventRecord(m->scratch_free, st));
        mean_done = true;
        var = nullptr;
    }
    if (var != nullptr || fuse_hess || (force_full && m->full.valid)) {
        if (var != nullptr && !m->has_invQ) return fail(GPE_ERR_INVALID, "variance requested but the model was created without invQ");
        if (!m->full.valid)
            return fail(GPE_ERR_UNSUPPORTED, "variance contraction supports M <= %d (got M = %d)", GPE_MAX_TRAIN, m->M);
        // up to three waves of 16-point tiles finish sooner than one wave of 64-point tiles
        const bool small = m->full_small.valid && !fuse_hess && call_N <= 3 * 16 * (int64_t)m->sms;
        const FullPlan& f = small ? m->full_small : m->full;
        FullParams p;
        memset(&p, 0, sizeof(p));
        p.testing = testing; p.N = N; p.mu = mu; p.var = var; p.deriv = deriv;
        p.ld_mu = ld_mu; p.ld_var = ld_var; p.ld_deriv = ld_deriv;
        p.xchunks = m->d_xchunks_full; p.s_tiled = m->d_stiled;
        p.M = m->M; p.D = m->D; p.Mp = f.Mp; p.nt_act = f.nt_act; p.kblk = f.kblk;
        p.nit = f.nit; p.nstage = f.nstage; p.lag = (f.nstage >= 3) ? 2 : 1; p.symmetric = m->symmetric ? 1 : 0;
        if (const char* e = getenv("GPE_RING_LAG")) p.lag = std::max(1, std::min(atoi(e), f.nstage - 1)); p.JC = f.JC; p.nchunks = f.nchunks; p.b = m->b; p.alias_x = f.alias_x;
        p.off_bar = f.off_bar; p.off_sqw = f.off_sqw; p.off_ks = f.off_ks; p.off_bst = f.off_bst;
        p.off_xc = f.off_xc; p.off_ts = f.off_ts; p.off_pa = f.off_pa; p.off_vred = f.off_vred;
        p.stage_bytes = f.stage_bytes; p.ts_bytes = f.ts_bytes;
        memcpy(p.sqrt_w, m->sqrt_w, sizeof(p.sqrt_w));
        p.trace = g_trace;
        if (fuse_hess) {
            p.hess = hess; p.ld_hess = ld_hess; p.p_tiled = m->d_ptiled; p.kbh = f.kbh; p.nit_h = f.nit_h; p.off_hts = f.off_hts;
            memcpy(p.centre, m->centre, sizeof(p.centre));
            hess = nullptr;   // done by this launch
        }
        const int64_t ntiles = (N + f.TN - 1) / f.TN;
        const int grid = (int)std::min<int64_t>(ntiles, (int64_t)m->sms * f.ctas_per_sm);
        CUDA_TRY(launch_full(m->DP, f.cfg, p, grid, f.smem, st));
        mean_done = true;
    }
    const bool need_mean = !mean_done && (mu != nullptr || deriv != nullptr);
    if (need_mean || hess != nullptr) {
        const bool do_hess = hess != nullptr;
        MeanParams p;
        memset(&p, 0, sizeof(p));
        p.testing = testing; p.N = N;
        p.mu = mean_done ? nullptr : mu;
        p.deriv = mean_done ? nullptr : deriv;
        p.hess = hess;
        p.ld_mu = ld_mu; p.ld_deriv = ld_deriv; p.ld_hess = ld_hess;
        p.xchunks = m->d_xchunks_mean; p.M = m->M; p.D = m->D; p.JC = m->mean.JC; p.nchunks = m->mean.nchunks;
        p.off_xc = m->mean.off_xc; p.off_ts = m->mean.off_ts; p.off_out = m->mean.off_out;
        memcpy(p.sqrt_w, m->sqrt_w, sizeof(p.sqrt_w));
        const int64_t ntiles = (N + kMeanTN - 1) / kMeanTN;
        // CTAs per SM: a multiple of what is resident (3 for the mean + gradient kernel, 2 / 4 for the Hessian ones)
        static const int env_mult = getenv("GPE_MEAN_GRID") ? atoi(getenv("GPE_MEAN_GRID")) : 0;
        const int mult = env_mult > 0 ? env_mult : (do_hess ? 8 : 12);
        const int grid = (int)std::min<int64_t>(ntiles, (int64_t)m->sms * mult);
        const size_t smem = do_hess ? m->mean.smem_hess : m->mean.smem;
        CUDA_TRY(launch_mean(m->DP, do_hess, p, dim3(grid), smem, st));
    }
    return GPE_OK;
}

bool is_pinned_or_null(const void* p) {
    if (p == nullptr) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int ensure(double** ptr, size_t* cap, size_t need, bool host) {
    if (*cap >= need) return GPE_OK;
    if (*ptr) {
        if (host) cudaFreeHost(*ptr); else cudaFree(*ptr);
        *ptr = nullptr; *cap = 0;
    }
    if (host) CUDA_TRY(cudaMallocHost((void**)ptr, need));
    else CUDA_TRY(cudaMalloc((void**)ptr, need));
    *cap = need;
    return GPE_OK;
}

// Staging copies for pageable callers.  One core moves ~14 GB/s on the GPU boxes, eight ~50 GB/s, and the PCIe link
// 55 GB/s each way, so copies of 1 MB and more are split into >= 256 KB pieces over a small persistent pool (created on first use;
// the submitting thread takes a share of the pieces itself).  Several threads may submit at once: the staged pipeline
// copies results out on its own thread while the caller's thread stages the next inputs.
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool pool;
        return pool;
    }
    void copy(void* dst, const void* src, size_t bytes) {
        constexpr size_t kPiece = 256u << 10;   // waking a worker costs tens of microseconds: not worth it below 1 MB
        const size_t want = bytes >= 4 * kPiece ? bytes / kPiece : 1;
        const unsigned np = (unsigned)std::min<size_t>(workers_.size() + 1, std::max<size_t>(want, 1));
        if (np <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t part = (bytes / np + 63) & ~(size_t)63;
        Batch batch;
        unsigned queued = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (unsigned i = 1; i < np; ++i) {
                const size_t off = (size_t)i * part;
                if (off >= bytes) break;
                q_.push_back(Task{(char*)dst + off, (const char*)src + off, std::min(part, bytes - off), &batch});
                ++queued;
            }
            batch.remaining = (int)queued;
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(part, bytes));
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return batch.remaining == 0; });
    }

private:
    struct Batch { int remaining = 0; };
    struct Task { char* dst; const char* src; size_t len; Batch* batch; };
    CopyPool() {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned n = std::min(7u, std::max(1u, hw / 2) - (hw >= 4 ? 1u : 0u));
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
            if (stop_) return;
            Task t = q_.front();
            q_.pop_front();
            lk.unlock();
            memcpy(t.dst, t.src, t.len);
            lk.lock();
            if (--t.batch->remaining == 0) done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<Task> q_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
};

void par_memcpy(void* dst, const void* src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }

// Host-resident caller: stream chunks through a few slots so the H2D copy of chunk i+1, the kernels of chunk i and
// the D2H copy of chunk i-1 overlap.
//   pinned caller buffers   : DMA'd directly, two slots, 2^18-point chunks;
//   pageable caller buffers : staged through page-locked slot buffers -- inputs and outputs independently, so a
//                             caller with pageable inputs and page-locked result arrays only pays for the copy-in.  The caller's thread stages inputs and
//                             enqueues (copy-in -> H2D -> kernels -> D2H -> event), a second thread waits for each
//                             chunk's event and copies its results out, both copying through the CopyPool; three
//                             slots keep the GPU busy while either side is late.  Chunks are about a quarter of the
//                             call (whole waves of 64-point tiles, at most 2^18 points) so that mid-sized calls
//                             (1e5 points) overlap too.
// T = double (FP64 path) or float (single-precision path); `launch(d_in, n, d_mu, d_var, d_deriv, d_hess, stream)`
// enqueues the kernels for one chunk.
template <typename T, typename Launch>
int predict_host_t(gpe_model* m, const T* testing, int64_t N, T* mu, T* var, T* deriv, T* hess, Launch launch) {
    constexpr size_t ES = sizeof(T);
    const int D = m->D;
    const int64_t per_out = (mu ? 1 : 0) + (var ? 1 : 0) + (deriv ? D : 0) + (hess ? (int64_t)D * D : 0);
    // inputs and outputs are staged independently: page-locked caller memory is DMA'd directly on either side
    // (a pointer query on pageable memory costs ~10 us: a small call with pageable inputs does not ask about its
    // outputs -- staging a few KB is cheaper than finding out)
    const bool in_direct = is_pinned_or_null(testing);
    const bool tiny = N <= 3 * 64 * (int64_t)m->sms;
    const bool out_direct = (in_direct || !tiny) && is_pinned_or_null(mu) && is_pinned_or_null(var) &&
                            is_pinned_or_null(deriv) && is_pinned_or_null(hess);
    const bool direct = in_direct && out_direct;
    int64_t CH = std::min<int64_t>(kPipeChunk, std::max<int64_t>(N, 1));
    if (!direct) {
        const int64_t wave = 64 * (int64_t)m->sms;
        CH = std::min<int64_t>(kPipeChunk, std::max<int64_t>(2 * wave, ((N + 3) / 4 + wave - 1) / wave * wave));
        if (N <= 3 * wave) CH = std::max<int64_t>(N, 1);   // too small to be worth a second thread
    }
    const int ns = direct ? 2 : 3;
    for (int i = 0; i < ns; ++i) {
        Slot& s = m->slots[i];
        if (!s.st) CUDA_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        if (!s.done) CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        int rc = ensure(&s.d_in, &s.d_in_cap, (size_t)CH * D * ES, false);
        if (rc) return rc;
        rc = ensure(&s.d_out, &s.d_out_cap, (size_t)CH * std::max<int64_t>(per_out, 1) * ES, false);
        if (rc) return rc;
        if (!in_direct) {
            rc = ensure(&s.h_in, &s.h_in_cap, (size_t)CH * D * ES, true);
            if (rc) return rc;
        }
        if (!out_direct) {
            rc = ensure(&s.h_out, &s.h_out_cap, (size_t)CH * std::max<int64_t>(per_out, 1) * ES, true);
            if (rc) return rc;
        }
        s.pending = false;
    }
    struct Outs { T *mu, *var, *der, *hes; };
    auto carve = [&](Slot& s, int64_t n) {
        T* o = reinterpret_cast<T*>(s.d_out);
        Outs r;
        r.mu = mu ? o : nullptr;    if (mu) o += n;
        r.var = var ? o : nullptr;  if (var) o += n;
        r.der = deriv ? o : nullptr; if (deriv) o += n * D;
        r.hes = hess ? o : nullptr;
        return r;
    };
    auto d2h_direct = [&](Slot& s, const Outs& o, int64_t n0, int64_t n) -> int {
        if (mu) CUDA_TRY(cudaMemcpyAsync(mu + n0, o.mu, (size_t)n * ES, cudaMemcpyDeviceToHost, s.st));
        if (var) CUDA_TRY(cudaMemcpyAsync(var + n0, o.var, (size_t)n * ES, cudaMemcpyDeviceToHost, s.st));
        if (deriv) CUDA_TRY(cudaMemcpyAsync(deriv + n0 * D, o.der, (size_t)n * D * ES, cudaMemcpyDeviceToHost, s.st));
        if (hess) CUDA_TRY(cudaMemcpyAsync(hess + n0 * D * D, o.hes, (size_t)n * D * D * ES, cudaMemcpyDeviceToHost, s.st));
        return GPE_OK;
    };
    if (direct) {
        int which = 0;
        for (int64_t n0 = 0; n0 < N; n0 += CH, which ^= 1) {
            Slot& s = m->slots[which];
            const int64_t n = std::min(CH, N - n0);
            T* const d_in = reinterpret_cast<T*>(s.d_in);
            const Outs o = carve(s, n);
            CUDA_TRY(cudaMemcpyAsync(d_in, testing + n0 * D, (size_t)n * D * ES, cudaMemcpyHostToDevice, s.st));
            int rc = launch(d_in, n, o.mu, o.var, o.der, o.hes, s.st);
            if (rc) return rc;
            rc = d2h_direct(s, o, n0, n);
            if (rc) return rc;
        }
        for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamSynchronize(m->slots[i].st));
        return GPE_OK;
    }

    // ---- staged path (inputs and / or outputs in pageable memory) -------------------------------------------
    static const bool pipe_trace = getenv("GPE_PIPE_TRACE") != nullptr;   // dev aid: host-side time split of a call
    double t_wait = 0, t_out = 0, t_in = 0, t_block = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto copy_out = [&](Slot& s) -> cudaError_t {   // wait for the slot's chunk, then scatter its results to the caller
        double t0 = now();
        cudaError_t e = cudaEventSynchronize(s.done);
        if (e != cudaSuccess) return e;
        t_wait += now() - t0; t0 = now();
        if (out_direct) return cudaSuccess;          // the D2H copies went straight into the caller's arrays
        const int64_t n0 = s.pend_n0, n = s.pend_n;
        const T* src = reinterpret_cast<const T*>(s.h_out);
        if (mu) { par_memcpy(mu + n0, src, (size_t)n * ES); src += n; }
        if (var) { par_memcpy(var + n0, src, (size_t)n * ES); src += n; }
        if (deriv) { par_memcpy(deriv + n0 * D, src, (size_t)n * D * ES); src += n * D; }
        if (hess) { par_memcpy(hess + n0 * D * D, src, (size_t)n * D * D * ES); }
        t_out += now() - t0;
        return cudaSuccess;
    };
    auto stage_in = [&](Slot& s, int64_t n0, int64_t n) -> int {   // copy-in + enqueue of one chunk on the slot
        T* const d_in = reinterpret_cast<T*>(s.d_in);
        const Outs o = carve(s, n);
        if (in_direct) {
            CUDA_TRY(cudaMemcpyAsync(d_in, testing + n0 * D, (size_t)n * D * ES, cudaMemcpyHostToDevice, s.st));
        } else {
            double t0 = now();
            par_memcpy(s.h_in, testing + n0 * D, (size_t)n * D * ES);
            t_in += now() - t0;
            CUDA_TRY(cudaMemcpyAsync(d_in, s.h_in, (size_t)n * D * ES, cudaMemcpyHostToDevice, s.st));
        }
        int rc = launch(d_in, n, o.mu, o.var, o.der, o.hes, s.st);
        if (rc) return rc;
        if (out_direct) {
            rc = d2h_direct(s, o, n0, n);
            if (rc) return rc;
        } else {
            CUDA_TRY(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)n * per_out * ES, cudaMemcpyDeviceToHost, s.st));
        }
        CUDA_TRY(cudaEventRecord(s.done, s.st));
        s.pend_n0 = n0; s.pend_n = n;
        return GPE_OK;
    };
    const int64_t nchunks = (N + CH - 1) / CH;
    int rc = GPE_OK;
    static const bool no_zero_copy = getenv("GPE_NO_ZERO_COPY") != nullptr;   // dev aid: time the copy-based small path
    static const int64_t zc_max = getenv("GPE_ZERO_COPY_MAX") ? atoll(getenv("GPE_ZERO_COPY_MAX")) : kZeroCopyMax;   // dev aid
    if (N <= zc_max && !in_direct && !out_direct && !no_zero_copy) {
        // Small calls (the reference is typically called with ONE point): the two cudaMemcpyAsync of the staged
        // path cost more than the kernel.  The staging buffers are page-locked, hence mapped into the device's address
        // space (UVA): the kernels read the test rows from and write the results to host memory directly -- one launch
        // and one synchronisation instead of copy, launch, copy, synchronise.
        Slot& s = m->slots[0];
        double t0 = now();
        memcpy(s.h_in, testing, (size_t)N * D * ES);
        t_in += now() - t0;
        T* o = reinterpret_cast<T*>(s.h_out);
        T* const z_mu = mu ? o : nullptr;    if (mu) o += N;
        T* const z_var = var ? o : nullptr;  if (var) o += N;
        T* const z_der = deriv ? o : nullptr; if (deriv) o += N * D;
        T* const z_hes = hess ? o : nullptr;
        rc = launch(reinterpret_cast<T*>(s.h_in), N, z_mu, z_var, z_der, z_hes, s.st);
        if (rc) return rc;
        t0 = now();
        CUDA_TRY(cudaStreamSynchronize(s.st));
        t_wait += now() - t0; t0 = now();
        if (mu) memcpy(mu, z_mu, (size_t)N * ES);
        if (var) memcpy(var, z_var, (size_t)N * ES);
        if (deriv) memcpy(deriv, z_der, (size_t)N * D * ES);
        if (hess) memcpy(hess, z_hes, (size_t)N * D * D * ES);
        t_out += now() - t0;
    } else if (nchunks == 1) {
        rc = stage_in(m->slots[0], 0, N);
        if (rc) return rc;
        CUDA_TRY(copy_out(m->slots[0]));
    } else if (out_direct) {
        // only the inputs are staged: no second thread, a slot is reused once its previous chunk has finished
        for (int64_t c = 0; c < nchunks; ++c) {
            Slot& s = m->slots[c % 3];
            if (c >= 3) {
                const double t0 = now();
                CUDA_TRY(cudaEventSynchronize(s.done));
                t_block += now() - t0;
            }
            const int64_t n0 = c * CH;
            rc = stage_in(s, n0, std::min(CH, N - n0));
            if (rc) { cudaDeviceSynchronize(); return rc; }
        }
        for (int i = 0; i < 3; ++i) CUDA_TRY(cudaStreamSynchronize(m->slots[i].st));
    } else {
        // chunk c lives in slot c % 3.  `staged` / `drained` count chunks handed to / finished by the output thread.
        std::mutex mx;
        std::condition_variable cv;
        int64_t staged = 0, drained = 0;
        bool abort = false;
        cudaError_t out_err = cudaSuccess;
        const int device = m->device;
        std::thread out_thread([&] {
            cudaSetDevice(device);
            for (int64_t c = 0; c < nchunks; ++c) {
                {
                    std::unique_lock<std::mutex> lk(mx);
                    cv.wait(lk, [&] { return staged > c || abort; });
                    if (staged <= c) return;
                }
                const cudaError_t e = copy_out(m->slots[c % 3]);
                std::lock_guard<std::mutex> lk(mx);
                if (e != cudaSuccess) { out_err = e; abort = true; cv.notify_all(); return; }
                drained = c + 1;
                cv.notify_all();
            }
        });
        for (int64_t c = 0; c < nchunks && rc == GPE_OK; ++c) {
            {
                const double t0 = now();
                std::unique_lock<std::mutex> lk(mx);
                cv.wait(lk, [&] { return drained + 3 > c || abort; });   // the slot's previous chunk has left it
                if (abort) break;
                t_block += now() - t0;
            }
            const int64_t n0 = c * CH;
            rc = stage_in(m->slots[c % 3], n0, std::min(CH, N - n0));
            std::lock_guard<std::mutex> lk(mx);
            if (rc == GPE_OK) staged = c + 1; else abort = true;
            cv.notify_all();
        }
        out_thread.join();
        if (rc) {                     // stage_in failed: the error text is already set on this thread
            cudaDeviceSynchronize();
            return rc;
        }
        if (out_err != cudaSuccess) return fail(GPE_ERR_CUDA, "result copy-out failed: %s", cudaGetErrorString(out_err));
    }
    if (pipe_trace)
        fprintf(stderr, "[gpemu pipe] N=%lld in %lld chunks of %lld (in %s, out %s): out-thread wait %.3f ms, copy-out %.3f ms | "
                        "copy-in %.3f ms, caller blocked on a slot %.3f ms\n",
                (long long)N, (long long)nchunks, (long long)CH, in_direct ? "direct" : "staged",
                out_direct ? "direct" : "staged", t_wait * 1e3, t_out * 1e3, t_in * 1e3, t_block * 1e3);
    return GPE_OK;
}

// call_N: size of the user's call when this is one device's share of it (gpe_multi_predict), else N
int predict_host(gpe_model* m, const double* testing, int64_t N, double* mu, double* var, double* deriv,
                 double* hess, int64_t call_N = -1) {
    const int64_t D = m->D;
    if (call_N < 0) call_N = N;
    return predict_host_t<double>(m, testing, N, mu, var, deriv, hess,
                                  [&](double* d_in, int64_t n, double* a, double* b, double* c, double* h, cudaStream_t st) {
                                      return predict_device(m, d_in, n, a, b, c, h, 1, 1, D, D * D, st, call_N);
                                  });
}

// ---- PCA back-projection: out (R, W) = A (R, E) . basis (E, W), A addressed with three strides ------------
// A skinny FP64 GEMM (K = E <= 32) whose output write is the HBM-bound part (W = 2101 doubles per row) and whose
// 2 E W flop per row sit right at the FP64 ridge, so it runs on the FP64 tensor path.  CTA = 64 rows, 8 warps as
// 2 (rows) x 4 (columns); a warp keeps its A fragments (32 rows x E) in registers for the whole sweep over W and
// produces 32 x 32 output tiles with DMMA.8x8x4.  The basis is pre-tiled at bank creation as [ks][Wp][4]
// (b_tiled[ks][w][c] = basis[4 ks + c][w], zero padded), the same conflict-free B-fragment image the variance
// kernel uses, streamed in 128-column groups by TMA bulk copies through a two-stage mbarrier ring.  Output tiles
// go through a per-warp smem buffer (16 rows x pitch 40 doubles = conflict-free 16-byte fragment stores) so
// every global store instruction writes 256 contiguous bytes of one output row (rows are only 8-byte aligned:
// W is odd).  Two CTAs per SM (<= 128 registers, 80 KB smem) so that one CTA's store burst drains -- HBM absorbs
// ~25 B/clk per SM -- while the other's DMMAs run; ncu showed the single-CTA version lg_throttle-bound.
constexpr int kProjThreads = 256, kProjRows = 64, kProjCols = 128;
constexpr int kProjPitch = 40;  // doubles; = 8 (mod 16) so a quarter-warp's 16-byte fragment stores hit 32 banks
template <int KS>   // k-steps of 4: E <= 4 KS
__global__ void __launch_bounds__(kProjThreads, 2) k_project(const double* __restrict__ A, int64_t R, int RD, int64_t ldn,
                                                             int64_t lde, int64_t ldd, const double* __restrict__ b_tiled,
                                                             int E, int W, int Wp, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char psm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(psm);          // [2]
    double* stage = reinterpret_cast<double*>(psm + 128);       // [2][KS][128][4]
    double* ctile = stage + 2 * (size_t)KS * kProjCols * 4;     // [8 warps][16][kProjPitch]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = warp >> 2, wc = warp & 3;
    const int64_t r0 = (int64_t)blockIdx.x * kProjRows + wr * 32;
    const int ngroups = Wp / kProjCols;
    constexpr uint32_t ks_bytes = kProjCols * 4 * 8, stage_doubles = (uint32_t)KS * kProjCols * 4;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto load_group = [&](int g, int st) {   // thread 0
        mbar_arrive_expect_tx(&full[st], (uint32_t)KS * ks_bytes);
        for (int ks = 0; ks < KS; ++ks)
            tma_bulk_g2s(stage + (size_t)st * stage_doubles + (size_t)ks * kProjCols * 4,
                         b_tiled + ((size_t)ks * Wp + (size_t)g * kProjCols) * 4, ks_bytes, &full[st]);
    };
    if (tid == 0) {
        load_group(0, 0);
        if (ngroups > 1) load_group(1, 1);
    }
    // A fragments: lane holds A[row = 8 i + lane / 4][k = 4 ks + lane % 4]
    double a[4][KS];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + 8 * i + (lane >> 2);
        const int64_t base = (r < R) ? (r / RD) * ldn + (r % RD) * ldd : 0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int e = 4 * ks + (lane & 3);
            a[i][ks] = (r < R && e < E) ? A[base + (int64_t)e * lde] : 0.0;
        }
    }
    const int nrow = (int)max((int64_t)0, min((int64_t)32, R - r0));
    double* ct = ctile + (size_t)warp * (16 * kProjPitch);
    uint32_t par = 0;
    for (int g = 0; g < ngroups; ++g) {
        const int st = g & 1;
        mbar_wait(&full[st], (par >> st) & 1u);
        par ^= 1u << st;
        const double* bs = stage + (size_t)st * stage_doubles + (size_t)(wc * 32 + (lane >> 2)) * 4 + (lane & 3);
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double bf = bs[(size_t)ks * kProjCols * 4 + j * 32];
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i][ks], bf);
            }
        }
        const int w = g * kProjCols + wc * 32 + lane;
        double* o = out + r0 * W + w;
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // rows 0..15, then 16..31 of the warp tile
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<double2*>(ct + (8 * i + (lane >> 2)) * kProjPitch + 8 * j + 2 * (lane & 3)) =
                        make_double2(acc[2 * half + i][j][0], acc[2 * half + i][j][1]);
            __syncwarp();
            if (w < W) {
#pragma unroll 4
                for (int rr = 0; rr < 16; ++rr)
                    if (16 * half + rr < nrow) o[(int64_t)(16 * half + rr) * W] = ct[rr * kProjPitch + lane];
            }
        }
        __syncthreads();   // every warp is done with this stage: refill it with group g + 2
        if (tid == 0 && g + 2 < ngroups) load_group(g + 2, st);
    }
}

template <int KS>
cudaError_t launch_project(const double* A, int64_t R, int RD, int64_t ldn, int64_t lde, int64_t ldd,
                           const double* b_tiled, int E, int W, int Wp, double* out, cudaStream_t st) {
    const size_t psmem = 128 + 2 * (size_t)KS * kProjCols * 4 * 8 + 8 * 16 * kProjPitch * 8;
    cudaError_t e = cudaFuncSetAttribute(k_project<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    if (e != cudaSuccess) return e;
    k_project<KS><<<(unsigned)((R + kProjRows - 1) / kProjRows), kProjThreads, psmem, st>>>(A, R, RD, ldn, lde, ldd,
                                                                                          b_tiled, E, W, Wp, out);
    return cudaGetLastError();
}

// ---- single-precision (tcgen05 / TF32) path ------------------------------------------------------------------
uint32_t host_tf32_rna(float x) {   // round-to-nearest (ties away) to 10 mantissa bits, as cvt.rna.tf32.f32
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return u;
    return (u + 0x1000u) & 0xFFFFE000u;
}

int ensure_tf32(gpe_model* m) {
    if (m->tf.valid) return GPE_OK;
    const int M = m->M, D = m->D;
    if (M > 1024) return fail(GPE_ERR_UNSUPPORTED, "the single-precision tensor-core path supports M <= 1024 (got %d)", M);
    int DP = -1;
    for (int dp : kTfDpList) if (dp >= D) { DP = dp; break; }
    if (DP < 0) return fail(GPE_ERR_UNSUPPORTED, "D = %d not supported by the single-precision path", D);
    const int Mp = (M + 63) / 64 * 64, nslab = (M + 31) / 32;
    // ring = true: K* slabs through a 2-deep A ring, column passes (predict_tf32_big.cuh); x3: hi + lo operands
    auto make_plan = [&](bool ring, bool x3, int pass_cols) {
        Tf32Plan t;
        t.DP = DP; t.Mp = Mp; t.nslab = nslab; t.big = ring; t.pass_cols = std::min(pass_cols, Mp);
        t.bstage_bytes = (uint32_t)(ring ? t.pass_cols : Mp) * 128u;
        uint32_t off = 0;
        t.off_bar = off; off += 128;
        t.off_tmem = off; off += 16;
        off = align_up(off, 1024);
        t.off_a = off; off += (uint32_t)(ring ? (x3 ? 4 : 2) : nslab) * kTfTN * 128u;   // A ring or the whole K* tile
        t.off_b = off; off += 2 * t.bstage_bytes;
        t.off_x = off; off += align_up((uint32_t)Mp * (DP + 1) * 4u, 16);
        t.off_out = off; off += align_up((uint32_t)kTfTN * (D + 1) * 4u, 16);
        t.off_vred = off; off += 2u * kTfTN * 4u;
        t.smem = off;
        t.valid = t.smem <= kSmemMax;
        return t;
    };
    Tf32Plan fast = (Mp <= 256) ? make_plan(false, false, Mp) : make_plan(true, false, 512);
    if (!fast.valid && Mp > 256) fast = make_plan(true, false, 256);
    // 3xTF32: M <= 256 keeps the resident-K* kernel (K*_lo lives in tensor memory), larger M the ring kernel
    Tf32Plan prec = (Mp <= 256) ? make_plan(false, false, Mp) : make_plan(true, true, 512);
    for (int pc : {256, 128, 64}) if (!prec.valid && Mp > 256) prec = make_plan(true, true, pc);
    if (!fast.valid || !prec.valid)
        return fail(GPE_ERR_UNSUPPORTED, "single-precision path does not fit shared memory for M = %d, D = %d", M, D);

    std::vector<float> xa((size_t)Mp * (DP + 1), 0.f);
    for (int j = 0; j < M; ++j) {
        for (int d = 0; d < D; ++d) xa[(size_t)j * DP + d] = (float)(m->sqrt_w[d] * m->h_inputs[(size_t)j * D + d]);
        xa[(size_t)Mp * DP + j] = (float)(m->b * m->h_invQt[j]);
    }
    CUDA_TRY(cudaMalloc((void**)&m->d_xa_f32, xa.size() * 4));
    CUDA_TRY(cudaMemcpy(m->d_xa_f32, xa.data(), xa.size() * 4, cudaMemcpyHostToDevice));
    if (!m->h_invQ.empty()) {
        // B operand: slab s holds invQ[j][32 s .. 32 s + 31] for every output column j as a [Mp][128 B] image with
        // the 16-byte chunk index XOR-ed by (j % 8): exactly what a SWIZZLE_128B K-major UMMA descriptor reads.
        // hi = rna_tf32(x); lo = rna_tf32(x - hi) for the 3 x TF32 split.
        std::vector<uint32_t> bh((size_t)nslab * Mp * 32, 0u), bl((size_t)nslab * Mp * 32, 0u);
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < M; ++i) {
                const int s = i >> 5, c = (i & 31) >> 2, e = i & 3;
                const size_t at = (size_t)s * Mp * 32 + (size_t)j * 32 + (size_t)((c ^ (j & 7)) << 2) + e;
                const float x = (float)m->h_invQ[(size_t)j * M + i];
                const uint32_t hi = host_tf32_rna(x);
                float hf;
                memcpy(&hf, &hi, 4);
                bh[at] = hi;
                bl[at] = host_tf32_rna(x - hf);
            }
        CUDA_TRY(cudaMalloc((void**)&m->d_bslabs, bh.size() * 4));
        CUDA_TRY(cudaMemcpy(m->d_bslabs, bh.data(), bh.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc((void**)&m->d_bslabs_lo, bl.size() * 4));
        CUDA_TRY(cudaMemcpy(m->d_bslabs_lo, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice));
    }
    m->tf = fast;
    m->tfx = prec;
    std::vector<double>().swap(m->h_invQ);   // the FP64 host copy was only needed for this packing
    return GPE_OK;
}

int predict_device_f32(gpe_model* m, const float* testing, int64_t N, float* mu, float* var, float* deriv,
                       bool fast, cudaStream_t st) {
    if (N == 0) return GPE_OK;
    int rc = ensure_tf32(m);
    if (rc) return rc;
    if (var && !m->d_bslabs) return fail(GPE_ERR_INVALID, "variance requested but the model was created without invQ");
    const bool x3 = !fast && var != nullptr;   // without the variance both modes are plain FP32: use the lighter plan
    const Tf32Plan& t = x3 ? m->tfx : m->tf;
    const int64_t ntiles = (N + kTfTN - 1) / kTfTN;
    const int grid = (int)std::min<int64_t>(ntiles, m->sms);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (t.big) {
        Tf32BigParams p;
        memset(&p, 0, sizeof(p));
        p.testing = testing; p.N = N; p.mu = mu; p.var = var; p.deriv = deriv;
        p.ld_mu = 1; p.ld_var = 1; p.ld_deriv = m->D;
        p.xa = m->d_xa_f32; p.bslabs = m->d_bslabs; p.bslabs_lo = m->d_bslabs_lo;
        p.M = m->M; p.D = m->D; p.Mp = t.Mp; p.nslab = t.nslab; p.pass_cols = t.pass_cols; p.b = (float)m->b;
        p.off_bar = t.off_bar; p.off_a = t.off_a; p.off_b = t.off_b; p.off_x = t.off_x; p.off_out = t.off_out;
        p.off_vred = t.off_vred; p.off_tmem = t.off_tmem; p.bstage_bytes = t.bstage_bytes;
        for (int d = 0; d < 32; ++d) p.sqrt_w[d] = (float)m->sqrt_w[d];
        CUDA_TRY(launch_tf32_big(t.DP, x3, p, grid, t.smem, st));
        return GPE_OK;
    }
    Tf32Params p;
    memset(&p, 0, sizeof(p));
    p.testing = testing; p.N = N; p.mu = mu; p.var = var; p.deriv = deriv;
    p.ld_mu = 1; p.ld_var = 1; p.ld_deriv = m->D;
    p.xa = m->d_xa_f32; p.bslabs = m->d_bslabs; p.bslabs_lo = m->d_bslabs_lo;
    p.M = m->M; p.D = m->D; p.Mp = t.Mp; p.nslab = t.nslab; p.b = (float)m->b;
    p.off_bar = t.off_bar; p.off_a = t.off_a; p.off_b = t.off_b; p.off_x = t.off_x; p.off_out = t.off_out;
    p.off_vred = t.off_vred; p.off_tmem = t.off_tmem; p.bstage_bytes = t.bstage_bytes;
    for (int d = 0; d < 32; ++d) p.sqrt_w[d] = (float)m->sqrt_w[d];
    CUDA_TRY(launch_tf32(t.DP, x3, p, grid, t.smem, st));
    return GPE_OK;
}

uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    const uint64_t* p = (const uint64_t*)data;
    const size_t n = bytes / 8;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

std::mutex g_wrap_mu;
gpe_model* g_wrap_model = nullptr;
uint64_t g_wrap_key = 0;

}  // namespace

namespace gpe {

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int require_device(int device, int* sms) { return check_device(device, sms); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace gpe

// =================================================================================================
extern "C" {

const char* gpe_last_error(void) { return g_err; }
int gpe_version(void) { return GPE_VERSION; }
int64_t gpe_launch_count(void) { return g_launches.load(); }

// Developer aid (not part of the public header): CTA 0 of the fused kernel records clock64() at its phase
// boundaries for its first 64 tiles into `device_buf` (64 x 8 int64); pass NULL to switch tracing off.
void gpe_debug_trace(long long* device_buf) { g_trace = device_buf; }

int gpe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gpe_model_create(int device, int M, int D, const double* inputs, const double* expX, const double* invQt,
                     const double* invQ, gpe_model** out) {
    unsigned options = 0;
    if (const char* e = getenv("GPE_SYMMETRIC_VARIANCE")) options |= atoi(e) ? GPE_OPT_SYMMETRIC_VARIANCE : 0;
    return gpe_model_create_ex(device, M, D, inputs, expX, invQt, invQ, options, out);
}

int gpe_model_create_ex(int device, int M, int D, const double* inputs, const double* expX, const double* invQt,
                        const double* invQ, unsigned options, gpe_model** out) {
    if (!out) return fail(GPE_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!inputs || !expX || !invQt) return fail(GPE_ERR_INVALID, "inputs, expX and invQt must be non-NULL");
    if (M < 1) return fail(GPE_ERR_INVALID, "M must be >= 1 (got %d)", M);
    if (D < 1 || D > GPE_MAX_INPUTS) return fail(GPE_ERR_INVALID, "D must be in [1, %d] (got %d)", GPE_MAX_INPUTS, D);
    int sms = 0;
    int rc = check_device(device, &sms);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    gpe_model* m = new gpe_model();
    m->device = device; m->M = M; m->D = D; m->DP = pick_dp(D); m->sms = sms;
    m->b = expX[D];
    for (int d = 0; d < 32; ++d) m->sqrt_w[d] = (d < D) ? std::sqrt(expX[d]) : 0.0;
    m->has_invQ = invQ != nullptr;
    m->symmetric = (options & GPE_OPT_SYMMETRIC_VARIANCE) != 0;
    m->h_inputs.assign(inputs, inputs + (size_t)M * D);
    m->h_invQt.assign(invQt, invQt + M);
    if (invQ && M <= 1024) m->h_invQ.assign(invQ, invQ + (size_t)M * M);
    for (int d = 0; d <= D; ++d) m->h_expx[d] = expX[d];
    m->mean = plan_mean(M, D, m->DP);
    {
        std::vector<double> xc = pack_xchunks(M, D, m->DP, m->mean.JC, m->mean.nchunks, inputs, m->sqrt_w, invQt, m->b);
        cudaError_t e = cudaMalloc((void**)&m->d_xchunks_mean, xc.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(m->d_xchunks_mean, xc.data(), xc.size() * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { gpe_model_destroy(m); return fail(GPE_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(e)); }
    }
    if (invQ) {
        m->full = plan_full(M, D, m->DP);
        if (m->full.valid) {
            const FullPlan& f = m->full;
            std::vector<double> xc = pack_xchunks(M, D, m->DP, f.JC, f.nchunks, inputs, m->sqrt_w, invQt, m->b);
            // s_tiled[kb][j][c] = invQ[j][4 kb + c]
            std::vector<double> st((size_t)f.kblk * f.Mp * 4, 0.0);
            for (int j = 0; j < M; ++j)
                for (int i = 0; i < M; ++i) {
                    double v = invQ[(size_t)j * M + i];
                    if (m->symmetric)   // k^T S k = k^T T k with T_ij = S_ij + S_ji (i < j), S_jj (i == j), 0 (i > j)
                        v = (i < j) ? invQ[(size_t)j * M + i] + invQ[(size_t)i * M + j] : (i == j ? v : 0.0);
                    st[((size_t)(i >> 2) * f.Mp + j) * 4 + (i & 3)] = v;
                }
            cudaError_t e = cudaMalloc((void**)&m->d_xchunks_full, xc.size() * 8);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_xchunks_full, xc.data(), xc.size() * 8, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_stiled, st.size() * 8);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_stiled, st.data(), st.size() * 8, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { gpe_model_destroy(m); return fail(GPE_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(e)); }
            rc = build_hessian_operand(m, inputs, invQt);
            if (rc) { gpe_model_destroy(m); return rc; }
            // low-latency plan for small batches: a 64-point tile costs ~45 us at M = 250 even for one point
            if (f.cfg == 0 && f.Mp % 64 == 0 && getenv("GPE_NO_SMALL_PLAN") == nullptr) {
                FullPlan fs = plan_full(M, D, m->DP, f.Mp);
                if (fs.valid && fs.JC == f.JC && fs.nchunks == f.nchunks && fs.kblk == f.kblk) m->full_small = fs;
            }
        } else if (M <= GPE_MAX_TRAIN) {
            // beyond the fused kernel: s_tiled with the contraction padded to Mp (the K* scratch pads are zero)
            m->large_Mp = (M + 63) / 64 * 64;
            m->large_kblk = m->large_Mp / 4;
            m->large_nstage = 6;
            m->large_chunk = 16 * (int64_t)m->sms * 4;
            const size_t Mp = (size_t)m->large_Mp;
            std::vector<double> st((size_t)m->large_kblk * Mp * 4, 0.0);
            for (int j = 0; j < M; ++j)
                for (int i = 0; i < M; ++i) st[((size_t)(i >> 2) * Mp + j) * 4 + (i & 3)] = invQ[(size_t)j * M + i];
            const size_t scratch = (size_t)(m->large_chunk / 16) * m->large_kblk * 64 * 8;
            cudaError_t e = cudaMalloc((void**)&m->d_stiled, st.size() * 8);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_stiled, st.data(), st.size() * 8, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_kscratch, scratch);
            if (e == cudaSuccess) e = cudaMemset(m->d_kscratch, 0, scratch);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->scratch_free, cudaEventDisableTiming);
            if (e != cudaSuccess) { gpe_model_destroy(m); return fail(GPE_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(e)); }
            m->large_valid = true;
        }
    }
    *out = m;
    return GPE_OK;
}

int gpe_model_destroy(gpe_model* m) {
    if (!m) return GPE_OK;
    cudaSetDevice(m->device);
    for (auto& s : m->slots) free_slot(s);
    if (m->d_xchunks_full) cudaFree(m->d_xchunks_full);
    if (m->d_stiled) cudaFree(m->d_stiled);
    if (m->d_ptiled) cudaFree(m->d_ptiled);
    if (m->d_kscratch) cudaFree(m->d_kscratch);
    if (m->scratch_free) cudaEventDestroy(m->scratch_free);
    if (m->d_xchunks_mean) cudaFree(m->d_xchunks_mean);
    if (m->d_xa_f32) cudaFree(m->d_xa_f32);
    if (m->d_bslabs) cudaFree(m->d_bslabs);
    if (m->d_bslabs_lo) cudaFree(m->d_bslabs_lo);
    delete m;
    return GPE_OK;
}

int gpe_predict(gpe_model* m, const double* testing, int64_t N, double* mu, double* var, double* deriv, double* hess,
                unsigned flags, void* stream) {
    NvtxRange nvtx_range("gpe_predict");
    if (!m) return fail(GPE_ERR_INVALID, "model is NULL");
    if (N < 0) return fail(GPE_ERR_INVALID, "N must be >= 0");
    if (N == 0) return GPE_OK;
    if (!testing) return fail(GPE_ERR_INVALID, "testing is NULL");
    if (!(flags & GPE_WANT_MU)) mu = nullptr;
    if (!(flags & GPE_WANT_VAR)) var = nullptr;
    if (!(flags & GPE_WANT_DERIV)) deriv = nullptr;
    if (!(flags & GPE_WANT_HESS)) hess = nullptr;
    if ((flags & GPE_WANT_MU) && !mu) return fail(GPE_ERR_INVALID, "GPE_WANT_MU set but mu is NULL");
    if ((flags & GPE_WANT_VAR) && !var) return fail(GPE_ERR_INVALID, "GPE_WANT_VAR set but var is NULL");
    if ((flags & GPE_WANT_DERIV) && !deriv) return fail(GPE_ERR_INVALID, "GPE_WANT_DERIV set but deriv is NULL");
    if ((flags & GPE_WANT_HESS) && !hess) return fail(GPE_ERR_INVALID, "GPE_WANT_HESS set but hess is NULL");
    if (!mu && !var && !deriv && !hess) return fail(GPE_ERR_INVALID, "no output requested");
    CUDA_TRY(cudaSetDevice(m->device));
    if (flags & GPE_HOST_PTRS) {
        std::lock_guard<std::mutex> lock(m->host_mu);
        return predict_host(m, testing, N, mu, var, deriv, hess);
    }
    return predict_device(m, testing, N, mu, var, deriv, hess, 1, 1, m->D, (int64_t)m->D * m->D,
                          (cudaStream_t)stream);
}

int gpe_predict_f32(gpe_model* m, const float* testing, int64_t N, float* mu, float* var, float* deriv, unsigned flags,
                    void* stream) {
    if (!m) return fail(GPE_ERR_INVALID, "model is NULL");
    if (N < 0) return fail(GPE_ERR_INVALID, "N must be >= 0");
    if (N == 0) return GPE_OK;
    if (!testing) return fail(GPE_ERR_INVALID, "testing is NULL");
    if (flags & GPE_WANT_HESS) return fail(GPE_ERR_UNSUPPORTED, "the single-precision path has no Hessian output");
    if (!(flags & GPE_WANT_MU)) mu = nullptr;
    if (!(flags & GPE_WANT_VAR)) var = nullptr;
    if (!(flags & GPE_WANT_DERIV)) deriv = nullptr;
    if ((flags & GPE_WANT_MU) && !mu) return fail(GPE_ERR_INVALID, "GPE_WANT_MU set but mu is NULL");
    if ((flags & GPE_WANT_VAR) && !var) return fail(GPE_ERR_INVALID, "GPE_WANT_VAR set but var is NULL");
    if ((flags & GPE_WANT_DERIV) && !deriv) return fail(GPE_ERR_INVALID, "GPE_WANT_DERIV set but deriv is NULL");
    if (!mu && !var && !deriv) return fail(GPE_ERR_INVALID, "no output requested");
    CUDA_TRY(cudaSetDevice(m->device));
    // variance mode: 3xTF32 split unless asked otherwise -- except for M > 256, where the FP32 accumulation of the
    // long sums dominates the error anyway (1.1e-5 vs 1.5e-5 at M = 1000) and the split costs 3.6x: there the
    // single pass is the default and GPE_F32_FORCE_3X opts in
    const bool fast = (flags & GPE_F32_FAST_TF32) != 0 || (m->M > 256 && !(flags & GPE_F32_FORCE_3X));
    std::lock_guard<std::mutex> lock(m->host_mu);   // also covers the lazy packing of the FP32 operands
    if (!(flags & GPE_HOST_PTRS)) return predict_device_f32(m, testing, N, mu, var, deriv, fast, (cudaStream_t)stream);
    // host pointers: the same overlapped pipeline as the FP64 path
    return predict_host_t<float>(m, testing, N, mu, var, deriv, (float*)nullptr,
                                 [&](float* d_in, int64_t n, float* a, float* b, float* c, float*, cudaStream_t st) {
                                     return predict_device_f32(m, d_in, n, a, b, c, fast, st);
                                 });
}

int gpe_predict_wrap(const double* expX, const double* inputs, const double* invQt, const double* invQ,
                     const double* testing, double* result, double* error, double* deriv, int Npredict, int Ntrain,
                     int Ninputs, int theta_size) {
    if (theta_size < Ninputs + 1) return fail(GPE_ERR_INVALID, "theta_size %d < Ninputs + 1", theta_size);
    if (!expX || !inputs || !invQt || !invQ || !testing || !result || !error || !deriv)
        return fail(GPE_ERR_INVALID, "NULL array argument");
    if (Npredict < 0) return fail(GPE_ERR_INVALID, "Npredict < 0");
    std::lock_guard<std::mutex> lock(g_wrap_mu);
    uint64_t key = 1469598103934665603ull ^ ((uint64_t)Ntrain << 32 | (uint32_t)Ninputs);
    key = fnv1a(expX, (size_t)(Ninputs + 1) * 8, key);
    key = fnv1a(inputs, (size_t)Ntrain * Ninputs * 8, key);
    key = fnv1a(invQt, (size_t)Ntrain * 8, key);
    key = fnv1a(invQ, (size_t)Ntrain * Ntrain * 8, key);
    if (!g_wrap_model || key != g_wrap_key) {
        if (g_wrap_model) { gpe_model_destroy(g_wrap_model); g_wrap_model = nullptr; }
        int rc = gpe_model_create(0, Ntrain, Ninputs, inputs, expX, invQt, invQ, &g_wrap_model);
        if (rc) return rc;
        g_wrap_key = key;
    }
    if (Npredict == 0) return GPE_OK;
    std::vector<double> nd((size_t)Npredict * Ninputs);
    int rc = gpe_predict(g_wrap_model, testing, Npredict, result, error, nd.data(), nullptr,
                         GPE_WANT_MU | GPE_WANT_VAR | GPE_WANT_DERIV | GPE_HOST_PTRS, nullptr);
    if (rc) return rc;
    // the legacy extension hands deriv back as (Ninputs, Npredict); its caller transposes (GaussianProcess.py:321)
    for (int n = 0; n < Npredict; ++n)
        for (int d = 0; d < Ninputs; ++d) deriv[(size_t)d * Npredict + n] = nd[(size_t)n * Ninputs + d];
    return GPE_OK;
}

// ---- one call, G devices: host-resident test points are split into contiguous ranges, one host thread per
// device drives that device's streaming pipeline (no steady-state exchange) ------------
int gpe_multi_create(int n_devices, const int* devices, int M, int D, const double* inputs, const double* expX,
                     const double* invQt, const double* invQ, unsigned options, gpe_multi** out) {
    if (!out) return fail(GPE_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_devices < 1 || !devices) return fail(GPE_ERR_INVALID, "need at least one device");
    gpe_multi* mm = new gpe_multi();
    for (int i = 0; i < n_devices; ++i) {
        gpe_model* m = nullptr;
        int rc = gpe_model_create_ex(devices[i], M, D, inputs, expX, invQt, invQ, options, &m);
        if (rc) { gpe_multi_destroy(mm); return rc; }
        mm->models.push_back(m);
    }
    *out = mm;
    return GPE_OK;
}

int gpe_multi_destroy(gpe_multi* mm) {
    if (!mm) return GPE_OK;
    for (gpe_model* m : mm->models) gpe_model_destroy(m);
    delete mm;
    return GPE_OK;
}

int gpe_multi_predict(gpe_multi* mm, const double* testing, int64_t N, double* mu, double* var, double* deriv,
                      double* hess, unsigned flags) {
    if (!mm) return fail(GPE_ERR_INVALID, "handle is NULL");
    if (N < 0) return fail(GPE_ERR_INVALID, "N must be >= 0");
    if (N == 0) return GPE_OK;
    const int G = (int)mm->models.size();
    const int64_t D = mm->models[0]->D;
    std::vector<int> rcs(G, GPE_OK);
    std::vector<std::string> msgs(G);
    std::vector<std::thread> th;
    const int64_t base = N / G, rem = N % G;
    for (int g = 0; g < G; ++g) {
        const int64_t lo = g * base + std::min<int64_t>(g, rem), n = base + (g < rem ? 1 : 0);
        if (n == 0) continue;
        th.emplace_back([=, &rcs, &msgs] {
            gpe_model* m = mm->models[g];
            const bool w_mu = flags & GPE_WANT_MU, w_var = flags & GPE_WANT_VAR, w_der = flags & GPE_WANT_DERIV,
                       w_hes = flags & GPE_WANT_HESS;
            if ((w_mu && !mu) || (w_var && !var) || (w_der && !deriv) || (w_hes && !hess) ||
                !(w_mu || w_var || w_der || w_hes) || !testing) {
                rcs[g] = fail(GPE_ERR_INVALID, "output flag set with a NULL array, or nothing requested");
            } else if (cudaSetDevice(m->device) != cudaSuccess) {
                rcs[g] = fail(GPE_ERR_CUDA, "cudaSetDevice(%d) failed", m->device);
            } else {
                // the plan (tile size) follows the size of the whole call, so G devices reproduce one device bit for bit
                std::lock_guard<std::mutex> lock(m->host_mu);
                rcs[g] = predict_host(m, testing + lo * D, n, w_mu ? mu + lo : nullptr, w_var ? var + lo : nullptr,
                                      w_der ? deriv + lo * D : nullptr, w_hes ? hess + lo * D * D : nullptr, N);
            }
            if (rcs[g]) msgs[g] = gpe_last_error();   // thread-local in the worker: carry it out
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], "device %d: %s", mm->models[g]->device, msgs[g].c_str());
    return GPE_OK;
}

int gpe_bank_create(int device, int E, int M, int D, const double* inputs, const double* expX, const double* invQt,
                    const double* invQ, const double* basis, int W, gpe_bank** out) {
    if (!out) return fail(GPE_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (E < 1) return fail(GPE_ERR_INVALID, "E must be >= 1");
    if (basis && W < 1) return fail(GPE_ERR_INVALID, "basis given but W < 1");
    gpe_bank* b = new gpe_bank();
    b->device = device; b->E = E; b->M = M; b->D = D; b->W = basis ? W : 0;
    for (int e = 0; e < E; ++e) {
        gpe_model* m = nullptr;
        int rc = gpe_model_create(device, M, D, inputs, expX + (size_t)e * (D + 1), invQt + (size_t)e * M,
                                  invQ ? invQ + (size_t)e * M * M : nullptr, &m);
        if (rc) { gpe_bank_destroy(b); return rc; }
        b->models.push_back(m);
    }
    {
        std::vector<MeanBankEntry> ent(E);
        for (int e2 = 0; e2 < E; ++e2) {
            ent[e2].xchunks = b->models[e2]->d_xchunks_mean;
            memcpy(ent[e2].sqrt_w, b->models[e2]->sqrt_w, sizeof(ent[e2].sqrt_w));
        }
        cudaError_t e = cudaMalloc((void**)&b->d_entries, sizeof(MeanBankEntry) * E);
        if (e == cudaSuccess) e = cudaMemcpy(b->d_entries, ent.data(), sizeof(MeanBankEntry) * E, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { gpe_bank_destroy(b); return fail(GPE_ERR_CUDA, "bank upload failed: %s", cudaGetErrorString(e)); }
    }
    if (basis) {
        const int ks_e = (E + 3) / 4;
        const int ks_n = ks_e <= 3 ? 3 : (ks_e <= 5 ? 5 : 8);   // the k-step count of the kernel instantiation used
        b->Wp = (W + kProjCols - 1) / kProjCols * kProjCols;
        std::vector<double> bt((size_t)ks_n * b->Wp * 4, 0.0);
        for (int e2 = 0; e2 < E; ++e2)
            for (int w = 0; w < W; ++w) bt[((size_t)(e2 >> 2) * b->Wp + w) * 4 + (e2 & 3)] = basis[(size_t)e2 * W + w];
        cudaError_t e = cudaMalloc((void**)&b->d_basis, bt.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(b->d_basis, bt.data(), bt.size() * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { gpe_bank_destroy(b); return fail(GPE_ERR_CUDA, "basis upload failed: %s", cudaGetErrorString(e)); }
    }
    *out = b;
    return GPE_OK;
}

int gpe_bank_destroy(gpe_bank* b) {
    if (!b) return GPE_OK;
    for (gpe_model* m : b->models) gpe_model_destroy(m);
    if (b->d_basis) { cudaSetDevice(b->device); cudaFree(b->d_basis); }
    if (b->d_entries) { cudaSetDevice(b->device); cudaFree(b->d_entries); }
    if (b->fwd_h) cudaFreeHost(b->fwd_h);
    if (b->fwd_d) { cudaSetDevice(b->device); cudaFree(b->fwd_d); }
    if (b->fwd_st) cudaStreamDestroy(b->fwd_st);
    if (b->cost_d) { cudaSetDevice(b->device); cudaFree(b->cost_d); }
    if (b->cost_free) cudaEventDestroy(b->cost_free);
    delete b;
    return GPE_OK;
}

int gpe_bank_predict(gpe_bank* b, const double* testing, int64_t N, double* mu, double* var, double* deriv,
                     double* hess, unsigned flags, void* stream) {
    NvtxRange nvtx_range("gpe_bank_predict");
    if (!b) return fail(GPE_ERR_INVALID, "bank is NULL");
    if (N < 0) return fail(GPE_ERR_INVALID, "N must be >= 0");
    if (N == 0) return GPE_OK;
    if (flags & GPE_HOST_PTRS) return fail(GPE_ERR_UNSUPPORTED, "gpe_bank_predict takes device pointers");
    if (!testing) return fail(GPE_ERR_INVALID, "testing is NULL");
    if (!(flags & GPE_WANT_MU)) mu = nullptr;
    if (!(flags & GPE_WANT_VAR)) var = nullptr;
    if (!(flags & GPE_WANT_DERIV)) deriv = nullptr;
    if (!(flags & GPE_WANT_HESS)) hess = nullptr;
    if (!mu && !var && !deriv && !hess) return fail(GPE_ERR_INVALID, "no output requested");
    CUDA_TRY(cudaSetDevice(b->device));
    const int64_t E = b->E, D = b->D;
    const bool with_var = var != nullptr;
    if (with_var) {   // the variance contraction is per emulator: one fused launch each (also yields mean + gradient,
                      // and the Hessian when every emulator qualifies for the fused Hessian)
        bool fuse_hess = hess != nullptr;
        for (int64_t e = 0; e < E && fuse_hess; ++e) fuse_hess = b->models[e]->hess_fused_ok;
        for (int64_t e = 0; e < E; ++e) {
            int rc = predict_device(b->models[e], testing, N, mu ? mu + e : nullptr, var + e, deriv ? deriv + e * D : nullptr,
                                    fuse_hess ? hess + e * D * D : nullptr, E, E, E * D, E * D * D, (cudaStream_t)stream);
            if (rc) return rc;
        }
        if (fuse_hess) hess = nullptr;
    } else if (hess != nullptr && N >= 16384) {
        // Hessian without variance on a large batch: per-emulator fused launches (phase A + phase C) beat the
        // one-launch direct kernel when every emulator qualifies and tiles are 64 points (cfg 0)
        bool fuse_hess = true;
        for (int64_t e = 0; e < E && fuse_hess; ++e) fuse_hess = b->models[e]->hess_fused_ok && b->models[e]->full.cfg == 0;
        if (fuse_hess) {
            for (int64_t e = 0; e < E; ++e) {
                int rc = predict_device(b->models[e], testing, N, mu ? mu + e : nullptr, nullptr, deriv ? deriv + e * D : nullptr,
                                        hess + e * D * D, E, E, E * D, E * D * D, (cudaStream_t)stream);
                if (rc) return rc;
            }
            return GPE_OK;
        }
    }
    if (hess != nullptr || (!with_var && (mu != nullptr || deriv != nullptr))) {
        // mean / gradient / Hessian of ALL emulators in one launch (blockIdx.y = emulator)
        gpe_model* m0 = b->models[0];
        const bool do_hess = hess != nullptr;
        if (E > 65535) return fail(GPE_ERR_UNSUPPORTED, "bank size %d exceeds the grid limit", (int)E);
        MeanParams p;
        memset(&p, 0, sizeof(p));
        p.testing = testing; p.N = N;
        p.mu = with_var ? nullptr : mu;
        p.deriv = with_var ? nullptr : deriv;
        p.hess = hess;
        p.ld_mu = E; p.ld_deriv = E * D; p.ld_hess = E * D * D;
        p.eo_mu = 1; p.eo_deriv = D; p.eo_hess = D * D;
        p.bank = b->d_entries;
        p.M = m0->M; p.D = m0->D; p.JC = m0->mean.JC; p.nchunks = m0->mean.nchunks;
        p.off_xc = m0->mean.off_xc; p.off_ts = m0->mean.off_ts; p.off_out = m0->mean.off_out;
        const int64_t ntiles = (N + kMeanTN - 1) / kMeanTN;
        const int gx = (int)std::min<int64_t>(ntiles, std::max<int64_t>(1, (int64_t)m0->sms * 8 / E));
        const size_t smem = do_hess ? m0->mean.smem_hess : m0->mean.smem;
        CUDA_TRY(launch_mean(m0->DP, do_hess, p, dim3(gx, (unsigned)E), smem, (cudaStream_t)stream));
    }
    return GPE_OK;
}

// Least-squares reduction over the emulators of a bank (outputs consumed on the fly).  One warp per
// point: lanes stride over the emulators for the residuals, then lane d accumulates column d of the gradient.
__global__ void __launch_bounds__(128) k_bank_cost(const double* __restrict__ mu, const double* __restrict__ deriv,
                                                   const double* __restrict__ obs, int64_t obs_ld,
                                                   const double* __restrict__ weights, double* __restrict__ cost,
                                                   double* __restrict__ grad, int64_t n, int E, int D) {
    extern __shared__ double wr_s[];                     // [4 warps][E] weighted residuals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* wr = wr_s + (size_t)w * E;
    for (int64_t pt = (int64_t)blockIdx.x * 4 + w; pt < n; pt += (int64_t)gridDim.x * 4) {
        double c = 0.0;
        for (int e = lane; e < E; e += 32) {
            const double r = mu[pt * E + e] - obs[pt * obs_ld + e];
            const double we = weights ? weights[e] : 1.0;
            wr[e] = we * r;
            c = fma(we * r, r, c);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && cost) cost[pt] = 0.5 * c;
        __syncwarp();
        if (grad != nullptr && lane < D) {
            const double* dp = deriv + pt * E * D + lane;
            double g = 0.0;
            for (int e = 0; e < E; ++e) g = fma(wr[e], dp[(size_t)e * D], g);
            grad[pt * D + lane] = g;
        }
        __syncwarp();
    }
}

int gpe_bank_cost(gpe_bank* b, const double* testing, int64_t N, const double* obs, int64_t obs_ld, const double* weights,
                  double* cost, double* grad, void* stream) {
    NvtxRange nvtx_range("gpe_bank_cost");
    if (!b) return fail(GPE_ERR_INVALID, "bank is NULL");
    if (N < 0) return fail(GPE_ERR_INVALID, "N must be >= 0");
    if (N == 0) return GPE_OK;
    if (!testing || !obs) return fail(GPE_ERR_INVALID, "testing / obs is NULL");
    if (!cost && !grad) return fail(GPE_ERR_INVALID, "no output requested");
    if (obs_ld != 0 && obs_ld < b->E) return fail(GPE_ERR_INVALID, "obs_ld must be 0 (one observation vector) or >= E");
    CUDA_TRY(cudaSetDevice(b->device));
    const int64_t E = b->E, D = b->D;
    const int64_t per_point = E * (1 + (grad ? D : 0));
    // points per chunk: whole waves of 64-point tiles, scratch bounded by 256 MB
    int64_t chunk = std::max<int64_t>(64, (((int64_t)256 << 20) / (per_point * 8)) / 64 * 64);
    chunk = std::min(chunk, (N + 63) / 64 * 64);
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(b->cost_mu);
    if (!b->cost_free) CUDA_TRY(cudaEventCreateWithFlags(&b->cost_free, cudaEventDisableTiming));
    // the scratch may still be read by a reduction enqueued on another stream
    CUDA_TRY(cudaStreamWaitEvent(st, b->cost_free, 0));
    if (b->cost_cap < (size_t)(chunk * per_point)) {
        if (b->cost_d) {
            CUDA_TRY(cudaEventSynchronize(b->cost_free));
            CUDA_TRY(cudaFree(b->cost_d));
            b->cost_d = nullptr; b->cost_cap = 0;
        }
        CUDA_TRY(cudaMalloc((void**)&b->cost_d, (size_t)(chunk * per_point) * 8));
        b->cost_cap = (size_t)(chunk * per_point);
    }
    double* d_mu = b->cost_d;
    double* d_der = grad ? b->cost_d + chunk * E : nullptr;
    int sms = b->models[0]->sms;
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
        const int64_t n = std::min(chunk, N - n0);
        int rc = gpe_bank_predict(b, testing + n0 * D, n, d_mu, nullptr, d_der, nullptr,
                                  GPE_WANT_MU | (grad ? GPE_WANT_DERIV : 0u), stream);
        if (rc) return rc;
        const int grid = (int)std::min<int64_t>((n + 3) / 4, (int64_t)sms * 16);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        k_bank_cost<<<grid, 128, (size_t)4 * E * 8, st>>>(d_mu, d_der, obs + n0 * obs_ld, obs_ld, weights,
                                                          cost ? cost + n0 : nullptr, grad ? grad + n0 * D : nullptr, n,
                                                          (int)E, (int)D);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(b->cost_free, st));
    return GPE_OK;
}

int gpe_bank_project(gpe_bank* b, const double* mu, const double* deriv, int64_t N, double* fwd, double* deriv_full,
                     void* stream) {
    if (!b) return fail(GPE_ERR_INVALID, "bank is NULL");
    if (!b->d_basis) return fail(GPE_ERR_INVALID, "bank was created without basis functions");
    if (N <= 0) return N == 0 ? GPE_OK : fail(GPE_ERR_INVALID, "N must be >= 0");
    if (b->E > 32) return fail(GPE_ERR_UNSUPPORTED, "back-projection supports E <= 32 (got %d)", b->E);
    CUDA_TRY(cudaSetDevice(b->device));
    const int E = b->E, D = b->D, W = b->W;
    cudaStream_t st = (cudaStream_t)stream;
    auto run = [&](const double* A, int64_t R, int RD, int64_t ldn, int64_t lde, int64_t ldd, double* o) {
        g_launches.fetch_add(1);
        const int ks = (E + 3) / 4;
        if (ks <= 3) return launch_project<3>(A, R, RD, ldn, lde, ldd, b->d_basis, E, W, b->Wp, o, st);
        if (ks <= 5) return launch_project<5>(A, R, RD, ldn, lde, ldd, b->d_basis, E, W, b->Wp, o, st);
        return launch_project<8>(A, R, RD, ldn, lde, ldd, b->d_basis, E, W, b->Wp, o, st);
    };
    if (fwd) {
        if (!mu) return fail(GPE_ERR_INVALID, "fwd requested but mu is NULL");
        CUDA_TRY(run(mu, N, 1, E, 1, 0, fwd));
    }
    if (deriv_full) {
        if (!deriv) return fail(GPE_ERR_INVALID, "deriv_full requested but deriv is NULL");
        CUDA_TRY(run(deriv, N * D, D, (int64_t)E * D, D, 1, deriv_full));
    }
    return GPE_OK;
}

int gpe_bank_forward(gpe_bank* b, const double* testing, int64_t N, double* fwd, double* deriv_full) {
    NvtxRange nvtx_range("gpe_bank_forward");
    if (!b) return fail(GPE_ERR_INVALID, "bank is NULL");
    if (!b->d_basis) return fail(GPE_ERR_INVALID, "bank was created without basis functions");
    if (N < 0) return fail(GPE_ERR_INVALID, "N must be >= 0");
    if (N == 0) return GPE_OK;
    if (!testing || !fwd) return fail(GPE_ERR_INVALID, "testing / fwd is NULL");
    std::lock_guard<std::mutex> lock(b->fwd_mu);
    CUDA_TRY(cudaSetDevice(b->device));
    if (!b->fwd_st) CUDA_TRY(cudaStreamCreateWithFlags(&b->fwd_st, cudaStreamNonBlocking));
    const int64_t E = b->E, D = b->D, W = b->W;
    // per point: testing D | mu E | deriv E*D stay on the device; fwd W [| deriv_full D*W] come back
    const int64_t out_pp = W + (deriv_full ? D * W : 0);
    const int64_t dev_pp = D + E + E * D + out_pp;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)(64u << 20) / (8 * out_pp)));
    int rc = ensure(&b->fwd_d, &b->fwd_d_cap, (size_t)chunk * dev_pp * 8, false);
    if (rc) return rc;
    rc = ensure(&b->fwd_h, &b->fwd_h_cap, (size_t)chunk * (D + out_pp) * 8, true);
    if (rc) return rc;
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
        const int64_t n = std::min(chunk, N - n0);
        double* d_t = b->fwd_d;
        double* d_mu = d_t + n * D;
        double* d_der = d_mu + n * E;
        double* d_out = d_der + n * E * D;          // fwd (n, W) | deriv_full (n, D, W)
        double* h_t = b->fwd_h;
        double* h_out = h_t + n * D;
        par_memcpy(h_t, testing + n0 * D, (size_t)n * D * 8);
        static const bool no_zero_copy = getenv("GPE_NO_ZERO_COPY") != nullptr;
        if (n <= kZeroCopyMax && (size_t)n * out_pp * 8 <= ((size_t)1 << 20) && !no_zero_copy) {
            // a few points (the reference's call is one): the kernels read the test rows from and write the spectra to
            // the page-locked staging buffer itself (mapped, UVA) -- no copy calls around the two launches
            d_t = h_t;
            d_out = h_out;
            rc = gpe_bank_predict(b, d_t, n, d_mu, nullptr, deriv_full ? d_der : nullptr, nullptr,
                                  GPE_WANT_MU | (deriv_full ? GPE_WANT_DERIV : 0u), b->fwd_st);
            if (rc) return rc;
            rc = gpe_bank_project(b, d_mu, deriv_full ? d_der : nullptr, n, d_out, deriv_full ? d_out + n * W : nullptr, b->fwd_st);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(b->fwd_st));
        } else {
        CUDA_TRY(cudaMemcpyAsync(d_t, h_t, (size_t)n * D * 8, cudaMemcpyHostToDevice, b->fwd_st));
        rc = gpe_bank_predict(b, d_t, n, d_mu, nullptr, deriv_full ? d_der : nullptr, nullptr,
                              GPE_WANT_MU | (deriv_full ? GPE_WANT_DERIV : 0u), b->fwd_st);
        if (rc) return rc;
        rc = gpe_bank_project(b, d_mu, deriv_full ? d_der : nullptr, n, d_out, deriv_full ? d_out + n * W : nullptr, b->fwd_st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_out, d_out, (size_t)n * out_pp * 8, cudaMemcpyDeviceToHost, b->fwd_st));
        CUDA_TRY(cudaStreamSynchronize(b->fwd_st));
        }
        par_memcpy(fwd + n0 * W, h_out, (size_t)n * W * 8);
        if (deriv_full) par_memcpy(deriv_full + n0 * D * W, h_out + n * W, (size_t)n * D * W * 8);
    }
    return GPE_OK;
}

}  // extern "C"
