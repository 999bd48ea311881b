// Single-precision GP prediction on tcgen05 + TMEM for 256 < M <= 1024.
//
// Same arithmetic as predict_tf32.cuh, but neither the K* tile (128 x 1024 FP32 = 512 KB) nor the accumulator
// (128 x 1024 > the 512 columns of tensor memory) fits on chip, so the contraction G = K* . invQ^T is organised as
//   * COLUMN PASSES of up to 512 output columns: the accumulator of one pass fills all 512 TMEM columns
//     (two UMMA N = 256 sub-blocks);
//   * inside a pass the compute warps sweep ALL training points, producing K* one 32-wide K slab at a time
//     into a two-deep shared-memory ring in the swizzled UMMA A-operand layout (TF32, cvt.rna); the control warp
//     multiplies each slab against the matching invQ slab (TMA ring, 64 KB stages) as soon as both have landed;
//     tcgen05.commit hands the A and B buffers back.  Pass 0 also accumulates the mean and gradient sums;
//   * the epilogue of a pass reads the accumulator rows with tcgen05.ld and multiplies them with K*_nj
//     RECOMPUTED on the fly for the pass's columns (distance + exp, ~16 issue slots per element with FFMA2) --
//     there is no on-chip home for a 128 x 512 FP32 copy beside the rings.
// X3 = true: 3xTF32 split.  K* and invQ are each split into hi = rna_tf32(x) and lo = rna_tf32(x - hi); the MMAs
// accumulate hi.hi + lo.hi + hi.lo (the dropped lo.lo term is 2^-22 relative), which restores FP32-grade accuracy
// (variance error ~1e-6 instead of ~1e-4) and meets the reference's own FP32 pass bar of 1e-5
// (tests/benchmark.py:56).  The A ring then holds hi and lo slabs; the B ring alternates hi and lo stages.
// Cost per point ~ (passes + 0.7) x the phase-A work, i.e. ~3x the small-M kernel per training point at M = 1000;
// still tensor-core cheap: the MMAs hide under the CUDA-core work.
#pragma once
#include "predict_tf32.cuh"

namespace gpe {

struct Tf32BigParams {
    const float* testing;
    int64_t N;
    float* mu;
    float* var;
    float* deriv;
    int64_t ld_mu, ld_var, ld_deriv;
    const float* xa;         // [Mp][DP] scaled inputs, then [Mp] b*alpha
    const uint32_t* bslabs;  // [nslab][Mp][32] TF32, swizzled per row
    const uint32_t* bslabs_lo;  // X3: the lo parts, same layout
    int M, D, Mp, nslab;     // Mp = ceil64(M) <= 1024, nslab = ceil(M / 32)
    int pass_cols;           // 512 or 256: output columns per pass
    float b;
    uint32_t off_bar, off_a, off_b, off_x, off_out, off_vred, off_tmem;
    uint32_t bstage_bytes;   // pass_cols * 128
    float sqrt_w[32];
};

template <int DP, bool X3>
__global__ void __launch_bounds__(kTfThreads, 1) k_predict_tf32_big(const Tf32BigParams p) {
    constexpr int ABUF = (X3 ? 2 : 1) * kTfTN * 128;   // bytes per A ring buffer: hi slab (+ lo slab)
    constexpr int TN = kTfTN;
    constexpr int NC = kTfComputeWarps * 32;
    extern __shared__ __align__(1024) unsigned char smem_tfb[];
    unsigned char* const smem = smem_tfb;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);  // [2]
    uint64_t* b_empty = b_full + 2;                                    // [2]
    uint64_t* a_ready = b_full + 4;                                    // [2]
    uint64_t* a_empty = b_full + 6;                                    // [2]
    uint64_t* acc_ready = b_full + 8;
    uint64_t* acc_empty = b_full + 9;
    uint64_t* x_bar = b_full + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_tmem);
    unsigned char* At = smem + p.off_a;   // 2 x [hi | lo][128 rows][128 B]
    unsigned char* Bt = smem + p.off_b;   // 2 x [pass_cols rows][128 B]
    float* Xs = reinterpret_cast<float*>(smem + p.off_x);
    float* outs = reinterpret_cast<float*>(smem + p.off_out);
    float* vred = reinterpret_cast<float*>(smem + p.off_vred);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.D, Mp = p.Mp, nslab = p.nslab, DV = D + 1;
    const int PW = p.pass_cols;
    const int npass = (Mp + PW - 1) / PW;
    const int64_t ntiles = (p.N + TN - 1) / TN;
    const bool want_var = p.var != nullptr;
    // the carve-up is ascending: barriers, TMEM slot, A, B ring, training set, output staging, variance partials
    smem_guard(umax2(umax2(p.off_vred + 2u * TN * 4u, p.off_out + (uint32_t)TN * (uint32_t)DV * 4u),
                     umax2(p.off_x + (uint32_t)Mp * (uint32_t)(DP + 1) * 4u, p.off_b + 2u * p.bstage_bytes)));

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
            mbar_init(&a_ready[i], NC);
            mbar_init(&a_empty[i], 1);
        }
        mbar_init(acc_ready, 1);
        mbar_init(acc_empty, NC);
        mbar_init(x_bar, 1);
        fence_mbar_init();
    }
    if (warp == kTfComputeWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (tid == 0) {
        const uint32_t bytes = (uint32_t)Mp * (DP + 1) * 4u;
        mbar_arrive_expect_tx(x_bar, bytes);
        tma_bulk_g2s(Xs, p.xa, bytes, x_bar);
    }

    if (warp == kTfComputeWarps) {
        // =============================== control warp: TMA producer + MMA issuer ===============================
        if (lane == 0 && want_var) {
            uint32_t bfull_par = 0, bempty_par = 0, aready_par = 0;
            uint32_t accE_par = 0;
            int64_t loads = 0, used = 0;   // global B ring counters (stage = counter & 1)
            int64_t aslabs = 0;            // global A ring counter (buffer = counter & 1)
            bool first_acc = true;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ps = 0; ps < npass; ++ps) {
                    const int c_lo = ps * PW, width = min(PW, Mp - c_lo);   // this pass's output columns
                    const uint32_t bbytes = (uint32_t)width * 128u;
                    // B ring uses: one per slab (X3: two per slab, hi then lo); use u -> stage u & 1
                    const int uses_per_pass = nslab * (X3 ? 2 : 1);
                    auto load_b = [&](int use) {
                        const int slab = X3 ? (use >> 1) : use;
                        const uint32_t* src = (X3 && (use & 1)) ? p.bslabs_lo : p.bslabs;
                        const int st = (int)(loads & 1);
                        if (loads >= 2) {
                            mbar_wait(&b_empty[st], (bempty_par >> st) & 1u);
                            bempty_par ^= 1u << st;
                        }
                        mbar_arrive_expect_tx(&b_full[st], bbytes);
                        tma_bulk_g2s(Bt + (size_t)st * p.bstage_bytes, src + ((size_t)slab * Mp + c_lo) * 32, bbytes,
                                     &b_full[st]);
                        ++loads;
                    };
                    load_b(0);
                    if (uses_per_pass > 1) load_b(1);
                    if (!first_acc) {   // the previous pass's epilogue must have drained the accumulator
                        mbar_wait(acc_empty, accE_par);
                        accE_par ^= 1;
                    }
                    first_acc = false;
                    auto mma_block = [&](uint32_t a_addr, uint32_t b_addr, bool first) {
                        for (int q = 0; q * 256 < width; ++q) {   // UMMA N <= 256: up to two column sub-blocks
                            const int nq = min(256, width - q * 256);
                            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nq >> 3) << 17) |
                                                   ((uint32_t)(TN >> 4) << 24);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_tf32(tmem_d + (uint32_t)(q * 256), umma_desc_sw128(a_addr + k * 32),
                                          umma_desc_sw128(b_addr + q * (256 * 128) + k * 32), idesc, !(first && k == 0));
                        }
                    };
                    int use = 0;
                    for (int s = 0; s < nslab; ++s) {
                        const int ab = (int)(aslabs & 1);
                        mbar_wait(&a_ready[ab], (aready_par >> ab) & 1u);
                        aready_par ^= 1u << ab;
                        const uint32_t a_hi = smem_u32(At + (size_t)ab * ABUF);
                        const uint32_t a_lo = a_hi + TN * 128;
                        {   // B hi stage: hi.hi (+ lo.hi)
                            const int st = (int)(used & 1);
                            mbar_wait(&b_full[st], (bfull_par >> st) & 1u);
                            bfull_par ^= 1u << st;
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(Bt + (size_t)st * p.bstage_bytes);
                            mma_block(a_hi, b_addr, s == 0);
                            if (X3) mma_block(a_lo, b_addr, false);
                            umma_commit(&b_empty[st]);
                            ++used; ++use;
                            if (use + 1 < uses_per_pass) load_b(use + 1);
                        }
                        if (X3) {   // B lo stage: hi.lo
                            const int st = (int)(used & 1);
                            mbar_wait(&b_full[st], (bfull_par >> st) & 1u);
                            bfull_par ^= 1u << st;
                            tc_fence_after();
                            mma_block(a_hi, smem_u32(Bt + (size_t)st * p.bstage_bytes), false);
                            umma_commit(&b_empty[st]);
                            ++used; ++use;
                            if (use + 1 < uses_per_pass) load_b(use + 1);
                        }
                        umma_commit(&a_empty[ab]);
                        ++aslabs;
                    }
                    umma_commit(acc_ready);
                }
            }
        }
    } else {
        // ======================================= compute warps ===============================================
        const int row = tid & (TN - 1);
        const int h = tid >> 7;
        const int sw = row & 7;
        uint32_t accR_par = 0;
        int64_t slab_uses = 0;   // global slab counter, mirrors the control warp's `used`
        mbar_wait(x_bar, 0);
        const float* al = Xs + Mp * DP;

        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t n0 = tile * TN;
            const int npts = (int)min((int64_t)TN, p.N - n0);
            const int64_t nrow = n0 + min(row, npts - 1);
            f32x2 ts2[DP / 2];
#pragma unroll
            for (int q = 0; q < DP / 2; ++q) {
                const int d = 2 * q;
                const float a = (d < D) ? __ldg(p.testing + nrow * D + d) * p.sqrt_w[d] : 0.f;
                const float b2 = (d + 1 < D) ? __ldg(p.testing + nrow * D + d + 1) * p.sqrt_w[d + 1] : 0.f;
                ts2[q] = pack2(a, b2);
            }
            // exp(-r2/2) of training point j for this thread's test row; optionally returns the differences
            auto kstar = [&](int j, f32x2 (&u2)[DP / 2]) -> float {
                const ulonglong2* xr = reinterpret_cast<const ulonglong2*>(Xs + j * DP);
                f32x2 racc = 0ull;
#pragma unroll
                for (int d4 = 0; d4 < DP / 4; ++d4) {
                    const ulonglong2 x = xr[d4];
                    u2[2 * d4] = sub2(x.x, ts2[2 * d4]);
                    u2[2 * d4 + 1] = sub2(x.y, ts2[2 * d4 + 1]);
                    racc = fma2(u2[2 * d4], u2[2 * d4], racc);
                    racc = fma2(u2[2 * d4 + 1], u2[2 * d4 + 1], racc);
                }
                float r_lo, r_hi;
                unpack2(racc, r_lo, r_hi);
                return ex2_approx((r_lo + r_hi) * -0.72134752044448170368f);
            };

            float mu = 0.f, vsum = 0.f;
            f32x2 g2[DP / 2];
#pragma unroll
            for (int q = 0; q < DP / 2; ++q) g2[q] = 0ull;

            const int npass_run = want_var ? npass : 1;
            for (int ps = 0; ps < npass_run; ++ps) {
                // ---- sweep all training points: K* slabs into the A ring (+ mean / gradient in pass 0) ----------
                for (int s = 0; s < nslab; ++s) {
                    const int buf = (int)(slab_uses & 1);
                    if (want_var && slab_uses >= 2) {   // the MMAs that read this buffer two slabs ago are done
                        mbar_wait(&a_empty[buf], (uint32_t)(((slab_uses >> 1) & 1) ^ 1));
                    }
                    unsigned char* arow = At + (size_t)buf * ABUF + row * 128;
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int cc = h + 2 * ci;          // 16-byte chunk inside the slab
                        float k4[4];
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const int j = 32 * s + 4 * cc + q4;
                            f32x2 u2[DP / 2];
                            const float k = kstar(j, u2);
                            k4[q4] = k;
                            if (ps == 0) {
                                const float cj = k * al[j];
                                mu += cj;
                                const f32x2 cj2 = pack2(cj, cj);
#pragma unroll
                                for (int q = 0; q < DP / 2; ++q) g2[q] = fma2(cj2, u2[q], g2[q]);
                            }
                        }
                        if (want_var) {
                            uint4 v;
                            v.x = tf32_rna(k4[0]); v.y = tf32_rna(k4[1]); v.z = tf32_rna(k4[2]); v.w = tf32_rna(k4[3]);
                            *reinterpret_cast<uint4*>(arow + ((cc ^ sw) << 4)) = v;
                            if (X3) {   // lo = rna(k - hi): the part of K* the TF32 mantissa dropped
                                uint4 w;
                                w.x = tf32_rna(k4[0] - __uint_as_float(v.x)); w.y = tf32_rna(k4[1] - __uint_as_float(v.y));
                                w.z = tf32_rna(k4[2] - __uint_as_float(v.z)); w.w = tf32_rna(k4[3] - __uint_as_float(v.w));
                                *reinterpret_cast<uint4*>(arow + TN * 128 + ((cc ^ sw) << 4)) = w;
                            }
                        }
                    }
                    if (want_var) {
                        fence_proxy_async();
                        mbar_arrive(&a_ready[buf]);
                        ++slab_uses;
                    }
                }
                if (ps == 0) {
                    // combine the two halves of every row, write mean and gradient
                    float g[DP];
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) unpack2(g2[q], g[2 * q], g[2 * q + 1]);
                    float* dst = outs + row * DV;
                    if (h == 1) {
                        dst[0] = mu;
#pragma unroll
                        for (int d = 0; d < DP; ++d)
                            if (d < D) dst[1 + d] = g[d];
                    }
                    bar_sync_compute();
                    if (h == 0) {
                        dst[0] += mu;
#pragma unroll
                        for (int d = 0; d < DP; ++d)
                            if (d < D) dst[1 + d] += g[d];
                    }
                    bar_sync_compute();
                    if (p.mu != nullptr && tid < npts) p.mu[(n0 + tid) * p.ld_mu] = outs[tid * DV];
                    if (p.deriv != nullptr) {
                        for (int e = tid; e < npts * D; e += NC) {
                            const int r = e / D, d = e - r * D;
                            p.deriv[(n0 + r) * p.ld_deriv + d] = p.sqrt_w[d] * outs[r * DV + 1 + d];
                        }
                    }
                }
                if (want_var) {
                    // ---- epilogue of the pass: accumulator rows x recomputed K* ----------------------------------
                    const int c_lo = ps * PW, width = min(PW, Mp - c_lo);
                    mbar_wait(acc_ready, accR_par);
                    accR_par ^= 1;
                    tc_fence_after();
                    const int q = warp & 3, ch = warp >> 2;
                    const int half_cols = width / 2;   // width is a multiple of 64
                    for (int c0 = ch * half_cols; c0 < (ch + 1) * half_cols; c0 += 32) {
                        float gv[32];
                        tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, gv);
#pragma unroll 8
                        for (int i = 0; i < 32; ++i) {
                            f32x2 u2[DP / 2];
                            vsum = fmaf(gv[i], kstar(c_lo + c0 + i, u2), vsum);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(acc_empty);
                }
            }
            if (want_var) {
                vred[h * TN + row] = vsum;
                bar_sync_compute();
                if (tid < npts) p.var[(n0 + tid) * p.ld_var] = p.b - p.b * p.b * (vred[tid] + vred[TN + tid]);
            }
            bar_sync_compute();   // outs / vred free for the next tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTfComputeWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_d));
    }
}

}  // namespace gpe
