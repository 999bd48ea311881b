// Batched evaluation of the GP training objective and its gradient.
//
// What it replaces: the inner loop of hyper-parameter fitting -- GaussianProcess.loglikelihood
// (gp_emulator/GaussianProcess.py:78-95) = _set_params -> _prepare_likelihood (:52-75: Z, Q = Z + noise I, inv(Q),
// invQ t, log|Q|) followed by partial_devs (:97-125) at the same theta.  The reference evaluates one theta at a time
// (~0.1 s at M = 250, D = 10 in numpy); fitting a MultivariateEmulator runs n_pcs x n_tries independent L-BFGS-B
// descents on the SAME training inputs (multivariate_gp.py:176-186), so their evaluations batch: one CTA per
// (theta, target vector) problem, B problems per launch.
//
// Per problem (one 1024-thread CTA, matrices in an L2-resident workspace, vectors in shared memory):
//   1. Z_ij = b exp(-1/2 sum_d w_d (x_id - x_jd)^2), Q = Z + noise I                      (:61-69)
//   2. in-place Gauss-Jordan inversion of Q without pivoting (Q is symmetric positive definite, so the pivots
//      are the LDL^T pivots: log|Q| = sum log p_k, and a pivot <= 0 is the reference's LinAlgError from
//      np.linalg.cholesky, :73-75).  M rank-1 updates of the whole matrix: M^3 FMA, 16 M^2 bytes of L2
//      traffic per pivot.
//   3. alpha = invQ t, t.alpha, alpha.alpha, trace(invQ)                                   (:71-72, :118-121)
//   4. g_d = -w_d/4 sum_ij (invQ_ij - alpha_i alpha_j) Z_ij (x_id - x_jd)^2,  g_D = 1/2 sum_ij (...) Z_ij,
//      g_{D+1} = noise/2 (trace(invQ) - alpha.alpha)                                       (:108-122)
//   loglik = 1/2 log|Q| + 1/2 t.alpha + M/2 log(2 pi)                                       (:90-92)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "../../include/gpemu.h"
#include "gpe_math.cuh"
#include "host_common.h"

namespace gpe {

constexpr int kTrainThreads = 1024;
constexpr int kTrainWarps = kTrainThreads / 32;

struct TrainParams {
    const double* x;        // (M, D) training inputs
    const double* targets;  // (T, M)
    const double* thetas;   // (B, D + 2)
    const int* tidx;        // (B) row of `targets` each problem fits
    double* work;           // (B, 2, M, M): [0] Q -> invQ in place, [1] Z
    double* loglik;         // (B)
    double* grad;           // (B, D + 2)
    int* status;            // (B) 0 ok, 1 Q not positive definite / non-finite
    int M, D;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA, result in every thread (kTrainWarps == 32: one partial per lane in the second stage).
__device__ __forceinline__ double block_sum(double v, double* red, int lane, int wid) {
    v = warp_sum(v);
    __syncthreads();  // `red` is free again
    if (lane == 0) red[wid] = v;
    __syncthreads();
    return warp_sum(red[lane]);
}

__global__ void __launch_bounds__(kTrainThreads, 1) k_train_eval(const TrainParams p) {
    extern __shared__ double sm[];
    const int M = p.M, D = p.D, nb = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* xs = sm;              // [M][D]
    double* tt = xs + M * D;      // [M] targets
    double* alpha = tt + M;       // [M]
    double* rowk = alpha + M;     // [M] pivot row, staged
    double* colk = rowk + M;      // [M] pivot column, staged
    double* ew = colk + M;        // [D + 2] exp(theta)
    double* red = ew + 40;        // [32]
    double* A = p.work + (size_t)nb * 2 * M * M;
    double* Z = A + (size_t)M * M;

    for (int i = tid; i < M * D; i += kTrainThreads) xs[i] = p.x[i];
    const double* t = p.targets + (size_t)p.tidx[nb] * M;
    for (int i = tid; i < M; i += kTrainThreads) tt[i] = t[i];
    if (tid < D + 2) ew[tid] = exp(p.thetas[(size_t)nb * (D + 2) + tid]);
    __syncthreads();
    const double bb = ew[D], noise = ew[D + 1];

    // 1. covariance of the training set
    for (int i = wid; i < M; i += kTrainWarps)
        for (int j = lane; j < M; j += 32) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) {
                const double df = xs[i * D + d] - xs[j * D + d];
                r2 += ew[d] * (df * df);
            }
            const double z = bb * exp_neg(-0.5 * r2);
            Z[(size_t)i * M + j] = z;
            A[(size_t)i * M + j] = (i == j) ? z + noise : z;
        }
    __syncthreads();

    // 2. Gauss-Jordan, pivot by pivot; two rows per warp pass so 16 loads are in flight per thread
    double logdet = 0.0;
    bool bad = false;
    for (int k = 0; k < M; ++k) {
        for (int i = tid; i < M; i += kTrainThreads) {
            rowk[i] = A[(size_t)k * M + i];
            colk[i] = A[(size_t)i * M + k];
        }
        __syncthreads();
        const double piv = rowk[k];
        if (!(piv > 0.0) || !(piv < 1e300)) { bad = true; break; }   // same value in every thread: uniform exit
        const double ip = 1.0 / piv;
        if (tid == 0) logdet += log(piv);
        for (int i0 = wid; i0 < M; i0 += 2 * kTrainWarps) {
            const int i1 = i0 + kTrainWarps;
            const bool has1 = i1 < M;
            const double c0 = colk[i0] * ip, c1 = has1 ? colk[i1] * ip : 0.0;
            double* a0 = A + (size_t)i0 * M;
            double* a1 = A + (size_t)(has1 ? i1 : i0) * M;
            for (int j0 = lane; j0 < M; j0 += 256) {
                double v0[8], v1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + 32 * u;
                    v0[u] = (j < M) ? a0[j] : 0.0;
                    v1[u] = (has1 && j < M) ? a1[j] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + 32 * u;
                    if (j < M) {
                        const double r = rowk[j];
                        a0[j] = (i0 == k) ? ((j == k) ? ip : r * ip) : ((j == k) ? -c0 : fma(-c0, r, v0[u]));
                        if (has1) a1[j] = (i1 == k) ? ((j == k) ? ip : r * ip) : ((j == k) ? -c1 : fma(-c1, r, v1[u]));
                    }
                }
            }
        }
        __syncthreads();
    }
    const int G = D + 2;
    if (bad) {
        if (tid == 0) {
            p.status[nb] = 1;
            p.loglik[nb] = nan("");
            for (int d = 0; d < G; ++d) p.grad[(size_t)nb * G + d] = nan("");
        }
        return;
    }

    // 3. alpha = invQ t and the scalar sums
    for (int i = wid; i < M; i += kTrainWarps) {
        double s = 0.0;
        for (int j = lane; j < M; j += 32) s = fma(A[(size_t)i * M + j], tt[j], s);
        s = warp_sum(s);
        if (lane == 0) alpha[i] = s;
    }
    __syncthreads();
    double s_ta = 0.0, s_aa = 0.0, s_tr = 0.0;
    for (int i = tid; i < M; i += kTrainThreads) {
        s_ta = fma(tt[i], alpha[i], s_ta);
        s_aa = fma(alpha[i], alpha[i], s_aa);
        s_tr += A[(size_t)i * M + i];
    }
    s_ta = block_sum(s_ta, red, lane, wid);
    s_aa = block_sum(s_aa, red, lane, wid);
    s_tr = block_sum(s_tr, red, lane, wid);

    // 4. gradient: one pass over (invQ - alpha alpha^T) o Z per hyper-parameter (D + 1 passes, L2-resident)
    for (int d = 0; d <= D; ++d) {
        double acc = 0.0;
        for (int i = wid; i < M; i += kTrainWarps) {
            const double ai = alpha[i];
            const double xi = (d < D) ? xs[i * D + d] : 0.0;
            double part = 0.0;
            for (int j = lane; j < M; j += 32) {
                double w = fma(-ai, alpha[j], A[(size_t)i * M + j]) * Z[(size_t)i * M + j];
                if (d < D) {
                    const double df = xi - xs[j * D + d];
                    w *= df * df;
                }
                part += w;
            }
            acc += part;
        }
        acc = block_sum(acc, red, lane, wid);
        if (tid == 0) p.grad[(size_t)nb * G + d] = (d < D) ? -0.25 * ew[d] * acc : 0.5 * acc;
    }
    if (tid == 0) {
        p.grad[(size_t)nb * G + D + 1] = 0.5 * noise * (s_tr - s_aa);
        p.loglik[nb] = 0.5 * logdet + 0.5 * s_ta + 0.5 * (double)M * 1.8378770664093453;  // log(2 pi)
        p.status[nb] = 0;
    }
}

}  // namespace gpe

using namespace gpe;

struct gpe_trainer {
    int device = 0, M = 0, D = 0, T = 0, sms = 0;
    cudaStream_t st = nullptr;
    double* d_x = nullptr;
    double* d_targets = nullptr;
    double* d_work = nullptr;     // capacity `cap` problems
    double* d_theta = nullptr;    // cap x (D + 2)
    double* d_ll = nullptr;
    double* d_grad = nullptr;
    int* d_tidx = nullptr;
    int* d_status = nullptr;
    int cap = 0;
    size_t smem = 0;
    std::mutex mu;                // evaluations on one trainer share its workspace
};

namespace {

#define TR_TRY(expr)                                                                                              \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess)                                                                                    \
            return set_error(GPE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

void free_batch_buffers(gpe_trainer* t) {
    if (t->d_work) cudaFree(t->d_work);
    if (t->d_theta) cudaFree(t->d_theta);
    if (t->d_ll) cudaFree(t->d_ll);
    if (t->d_grad) cudaFree(t->d_grad);
    if (t->d_tidx) cudaFree(t->d_tidx);
    if (t->d_status) cudaFree(t->d_status);
    t->d_work = t->d_theta = t->d_ll = t->d_grad = nullptr;
    t->d_tidx = t->d_status = nullptr;
    t->cap = 0;
}

int reserve(gpe_trainer* t, int nb) {
    if (nb <= t->cap) return GPE_OK;
    free_batch_buffers(t);
    const size_t G = (size_t)t->D + 2, mm = (size_t)t->M * t->M;
    TR_TRY(cudaMalloc((void**)&t->d_work, (size_t)nb * 2 * mm * 8));
    TR_TRY(cudaMalloc((void**)&t->d_theta, (size_t)nb * G * 8));
    TR_TRY(cudaMalloc((void**)&t->d_ll, (size_t)nb * 8));
    TR_TRY(cudaMalloc((void**)&t->d_grad, (size_t)nb * G * 8));
    TR_TRY(cudaMalloc((void**)&t->d_tidx, (size_t)nb * sizeof(int)));
    TR_TRY(cudaMalloc((void**)&t->d_status, (size_t)nb * sizeof(int)));
    t->cap = nb;
    return GPE_OK;
}

}  // namespace

extern "C" {

int gpe_trainer_create(int device, int M, int D, int T, const double* inputs, const double* targets, gpe_trainer** out) {
    if (!out) return set_error(GPE_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (M < 1 || D < 1 || T < 1) return set_error(GPE_ERR_INVALID, "need M >= 1, D >= 1, T >= 1 (got %d, %d, %d)", M, D, T);
    if (D > GPE_MAX_INPUTS) return set_error(GPE_ERR_UNSUPPORTED, "D = %d exceeds GPE_MAX_INPUTS = %d", D, GPE_MAX_INPUTS);
    if (M > GPE_TRAIN_MAX_M) return set_error(GPE_ERR_UNSUPPORTED, "M = %d exceeds GPE_TRAIN_MAX_M = %d", M, GPE_TRAIN_MAX_M);
    if (!inputs || !targets) return set_error(GPE_ERR_INVALID, "inputs / targets is NULL");
    const size_t smem = ((size_t)M * D + 4 * (size_t)M + 40 + 32) * 8;
    if (smem > 232448) return set_error(GPE_ERR_UNSUPPORTED, "M x D = %d x %d needs %zu bytes of shared memory per CTA", M, D, smem);
    int sms = 0;
    int rc = require_device(device, &sms);
    if (rc) return rc;
    TR_TRY(cudaSetDevice(device));
    gpe_trainer* t = new gpe_trainer;
    t->device = device; t->M = M; t->D = D; t->T = T; t->sms = sms; t->smem = smem;
    cudaError_t e = cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->d_x, (size_t)M * D * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->d_targets, (size_t)T * M * 8);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_x, inputs, (size_t)M * D * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_targets, targets, (size_t)T * M * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_train_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        gpe_trainer_destroy(t);
        return set_error(GPE_ERR_CUDA, "trainer setup failed: %s", cudaGetErrorString(e));
    }
    *out = t;
    return GPE_OK;
}

int gpe_trainer_destroy(gpe_trainer* t) {
    if (!t) return GPE_OK;
    cudaSetDevice(t->device);
    free_batch_buffers(t);
    if (t->d_x) cudaFree(t->d_x);
    if (t->d_targets) cudaFree(t->d_targets);
    if (t->st) cudaStreamDestroy(t->st);
    delete t;
    return GPE_OK;
}

int gpe_trainer_eval(gpe_trainer* t, int B, const int* target_index, const double* thetas, double* loglik, double* grad,
                     int* status) {
    if (!t) return set_error(GPE_ERR_INVALID, "trainer is NULL");
    if (B < 0) return set_error(GPE_ERR_INVALID, "B must be >= 0");
    if (B == 0) return GPE_OK;
    if (!thetas || !loglik || !grad || !status) return set_error(GPE_ERR_INVALID, "thetas / loglik / grad / status is NULL");
    if (target_index)
        for (int i = 0; i < B; ++i)
            if (target_index[i] < 0 || target_index[i] >= t->T)
                return set_error(GPE_ERR_INVALID, "target_index[%d] = %d out of range [0, %d)", i, target_index[i], t->T);
    std::lock_guard<std::mutex> lock(t->mu);
    TR_TRY(cudaSetDevice(t->device));
    const size_t G = (size_t)t->D + 2, mm = (size_t)t->M * t->M;
    // problems per launch: whole waves of CTAs, workspace bounded by 4 GB
    const int by_mem = (int)std::max<size_t>(1, ((size_t)4 << 30) / (2 * mm * 8));
    const int chunk = std::min(B, std::min(by_mem, 8 * t->sms));
    int rc = reserve(t, chunk);
    if (rc) return rc;
    std::vector<int> zeros;
    if (!target_index) zeros.assign((size_t)chunk, 0);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        TR_TRY(cudaMemcpyAsync(t->d_theta, thetas + (size_t)b0 * G, (size_t)nb * G * 8, cudaMemcpyHostToDevice, t->st));
        TR_TRY(cudaMemcpyAsync(t->d_tidx, target_index ? target_index + b0 : zeros.data(), (size_t)nb * sizeof(int),
                               cudaMemcpyHostToDevice, t->st));
        TrainParams p;
        p.x = t->d_x; p.targets = t->d_targets; p.thetas = t->d_theta; p.tidx = t->d_tidx; p.work = t->d_work;
        p.loglik = t->d_ll; p.grad = t->d_grad; p.status = t->d_status; p.M = t->M; p.D = t->D;
        // per function AND per device: set on every launch so multi-device processes stay correct
        TR_TRY(cudaFuncSetAttribute(k_train_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem));
        k_train_eval<<<nb, kTrainThreads, t->smem, t->st>>>(p);
        count_launch();
        TR_TRY(cudaGetLastError());
        TR_TRY(cudaMemcpyAsync(loglik + b0, t->d_ll, (size_t)nb * 8, cudaMemcpyDeviceToHost, t->st));
        TR_TRY(cudaMemcpyAsync(grad + (size_t)b0 * G, t->d_grad, (size_t)nb * G * 8, cudaMemcpyDeviceToHost, t->st));
        TR_TRY(cudaMemcpyAsync(status + b0, t->d_status, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, t->st));
        TR_TRY(cudaStreamSynchronize(t->st));
    }
    return GPE_OK;
}

}  // extern "C"
