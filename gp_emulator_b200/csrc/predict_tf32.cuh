// FP32 / TF32 GP prediction on the 5th-generation tensor cores (tcgen05 + TMEM), M <= 256.
//
// Same math as predict_full.cuh (reference GaussianProcess.py:232-247) in single precision:
//   * K* (FP32, exp via MUFU.EX2), mean and gradient sums on the CUDA cores;
//   * the variance contraction G = K* . invQ^T as tcgen05.mma.kind::tf32 (UMMA 128 x Mp x 8) with the FP32
//     accumulator tile (128 lanes x Mp columns) in TENSOR MEMORY.  The A operand is the K* tile itself: phase A
//     writes it straight into the K-major SWIZZLE_128B shared-memory image the UMMA descriptor expects (values
//     rounded to TF32 with cvt.rna, so the tensor core's mantissa truncation introduces no bias).  The B
//     operand (invQ, TF32-rounded and pre-swizzled on the host into per-K-slab images) is streamed by TMA bulk
//     copies through an mbarrier ring.  A control warp issues the MMAs slab by slab as soon as phase A has
//     finished the corresponding 32 columns of K*, so the tensor work hides under phase A;
//   * var_n = b - b^2 sum_j G_nj K*_nj: tcgen05.ld of the accumulator rows (one TMEM lane per test row)
//     against the K* tile.
// Precision "tf32" = one MMA per k-step (|rel err| ~ 2^-11 / sqrt(M) on the variance, measured in the tests);
// the mean and gradient are plain FP32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpe_ptx.cuh"

namespace gpe {

constexpr int kTfTN = 128;            // test rows per tile == UMMA M == TMEM lanes
constexpr int kTfComputeWarps = 8;
constexpr int kTfThreads = (kTfComputeWarps + 1) * 32;  // + 1 control warp (TMA producer + MMA issuer)
constexpr int kTfMaxSlabs = 8;        // K <= 256 in slabs of 32 (one 128-byte swizzle atom of FP32 per row)

struct Tf32Params {
    const float* testing;  // (N, D)
    int64_t N;
    float* mu;
    float* var;
    float* deriv;
    int64_t ld_mu, ld_var, ld_deriv;
    const float* xa;        // [Mp][DP] sqrt(w)-scaled inputs (FP32), then [Mp] b*alpha
    const uint32_t* bslabs; // [nslab][Mp rows (j)][32 (i)] TF32 bit patterns, 128B-swizzled smem image per slab
    const uint32_t* bslabs_lo;  // X3: rna_tf32(invQ - hi), same layout
    int M, D, Mp, nslab;   // Mp = ceil64(M) output columns (UMMA N); nslab = ceil(M / 32) K slabs
    float b;
    uint32_t off_bar, off_a, off_b, off_x, off_out, off_vred, off_tmem;
    uint32_t bstage_bytes;  // Mp * 128
    float sqrt_w[32];
};

// Packed FP32 pairs (Blackwell FFMA2 / FADD2: two FP32 operations per issue slot) -- phase A is issue bound.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2, rel. error 2^-22, flushes denormals
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version 1 at bit 46, layout 2 at bit 61);
// 8-row groups are 1024 bytes apart (SBO), LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

// A operand from TENSOR MEMORY (lane = row, 32-bit column = k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

// this thread's TMEM lane (lane base in taddr + lane id), 4 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(a), "r"(b), "r"(c),
                 "r"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void bar_sync_compute() {
    asm volatile("bar.sync 1, %0;\n" ::"n"(kTfComputeWarps * 32) : "memory");
}

// 32 lanes x 32 consecutive columns of TMEM -> 32 registers per thread (thread i <-> lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// DP: input dimension padded to a multiple of 4 (float4 training rows)
// X3: 3xTF32 split.  K*_hi lives in the shared-memory A tile, K*_lo in the upper 256 columns of TENSOR MEMORY (the
// A operand of tcgen05.mma may come from TMEM), invQ_hi / invQ_lo alternate through the B ring:
//   D += K*_hi . B_hi + K*_lo . B_hi + K*_hi . B_lo.
template <int DP, bool X3>
__global__ void __launch_bounds__(kTfThreads, 1) k_predict_tf32(const Tf32Params p) {
    constexpr uint32_t kTmemCols = X3 ? 512 : 256;   // accumulator [0, 256) (+ K*_lo [256, 512))
    constexpr int TN = kTfTN;
    constexpr int NC = kTfComputeWarps * 32;
    extern __shared__ __align__(1024) unsigned char smem_tf[];
    unsigned char* const smem = smem_tf;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);  // [2]
    uint64_t* b_empty = b_full + 2;                                    // [2]
    uint64_t* a_ready = b_full + 4;                                    // [kTfMaxSlabs]
    uint64_t* acc_ready = b_full + 12;
    uint64_t* acc_empty = b_full + 13;
    uint64_t* x_bar = b_full + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_tmem);
    unsigned char* At = smem + p.off_a;   // nslab x [128 rows][128 B], swizzled
    unsigned char* Bt = smem + p.off_b;   // 2 x [Mp rows][128 B], swizzled
    float* Xs = reinterpret_cast<float*>(smem + p.off_x);
    float* outs = reinterpret_cast<float*>(smem + p.off_out);   // [TN][D+1]
    float* vred = reinterpret_cast<float*>(smem + p.off_vred);  // [2][TN]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.D, Mp = p.Mp, nslab = p.nslab, DV = D + 1;
    const int64_t ntiles = (p.N + TN - 1) / TN;
    const bool want_var = p.var != nullptr;
    // the carve-up is ascending: barriers, TMEM slot, A, B ring, training set, output staging, variance partials
    smem_guard(umax2(umax2(p.off_vred + 2u * TN * 4u, p.off_out + (uint32_t)TN * (uint32_t)DV * 4u),
                     umax2(p.off_x + (uint32_t)Mp * (uint32_t)(DP + 1) * 4u, p.off_b + 2u * p.bstage_bytes)));

    if (tid == 0) {
        mbar_init(&b_full[0], 1);
        mbar_init(&b_full[1], 1);
        mbar_init(&b_empty[0], 1);
        mbar_init(&b_empty[1], 1);
        for (int s = 0; s < kTfMaxSlabs; ++s) mbar_init(&a_ready[s], NC);
        mbar_init(acc_ready, 1);
        mbar_init(acc_empty, NC);
        mbar_init(x_bar, 1);
        fence_mbar_init();
    }
    if (warp == kTfComputeWarps) {  // control warp owns the TMEM allocation (256 columns: Mp <= 256 FP32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    // training inputs + alpha: resident for the whole kernel
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)Mp * (DP + 1) * 4u;
        mbar_arrive_expect_tx(x_bar, bytes);
        tma_bulk_g2s(Xs, p.xa, bytes, x_bar);
    }

    // UMMA instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = Mp
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Mp >> 3) << 17) | ((uint32_t)(TN >> 4) << 24);

    if (warp == kTfComputeWarps) {
        // =============================== control warp: TMA producer + MMA issuer ===============================
        if (lane == 0 && want_var) {
            uint32_t full_par = 0, empty_par = 0;   // bit st = parity of the next completion of b_full / b_empty[st]
            uint32_t a_par = 0, accE_par = 0;
            int64_t loads = 0;  // B slabs issued so far (global count); slab g -> stage g & 1
            // B ring uses: one per slab (X3: two per slab, hi then lo); use u -> stage (global count) & 1
            const int uses_per_tile = nslab * (X3 ? 2 : 1);
            auto load_b = [&](int use) {
                const int slab = X3 ? (use >> 1) : use;
                const uint32_t* src = (X3 && (use & 1)) ? p.bslabs_lo : p.bslabs;
                const int st = (int)(loads & 1);
                if (loads >= 2) {  // the MMAs that read this stage two uses ago must have completed
                    mbar_wait(&b_empty[st], (empty_par >> st) & 1u);
                    empty_par ^= 1u << st;
                }
                mbar_arrive_expect_tx(&b_full[st], p.bstage_bytes);
                tma_bulk_g2s(Bt + (size_t)st * p.bstage_bytes, src + (size_t)slab * Mp * 32, p.bstage_bytes, &b_full[st]);
                ++loads;
            };
            int64_t used = 0;  // B ring uses consumed so far
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const bool first_tile = (tile == (int64_t)blockIdx.x);
                load_b(0);
                if (uses_per_tile > 1) load_b(1);
                if (!first_tile) {  // previous tile's epilogue must have drained the accumulator (and K*_lo)
                    mbar_wait(acc_empty, accE_par);
                    accE_par ^= 1;
                }
                int use = 0;
                for (int s = 0; s < nslab; ++s) {
                    mbar_wait(&a_ready[s], a_par);
                    const uint32_t a_addr = smem_u32(At + (size_t)s * (TN * 128));
                    {
                        const int st = (int)(used & 1);
                        mbar_wait(&b_full[st], (full_par >> st) & 1u);
                        full_par ^= 1u << st;
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(Bt + (size_t)st * p.bstage_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)   // 4 k-steps of 8 TF32 (32 bytes) inside the 128-byte swizzle atom
                            umma_tf32(tmem_d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                      (s | k) != 0);
                        if (X3) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // K*_lo from tensor memory: columns 256 + 32 s + 8 k ..
                                umma_tf32_ts(tmem_d, tmem_d + 256u + (uint32_t)(32 * s + 8 * k),
                                             umma_desc_sw128(b_addr + k * 32), idesc, true);
                        }
                        umma_commit(&b_empty[st]);   // arrives when these MMAs have finished reading the stage
                        ++used; ++use;
                        if (use + 1 < uses_per_tile) load_b(use + 1);
                    }
                    if (X3) {
                        const int st = (int)(used & 1);
                        mbar_wait(&b_full[st], (full_par >> st) & 1u);
                        full_par ^= 1u << st;
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(Bt + (size_t)st * p.bstage_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32(tmem_d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, true);
                        umma_commit(&b_empty[st]);
                        ++used; ++use;
                        if (use + 1 < uses_per_tile) load_b(use + 1);
                    }
                }
                umma_commit(acc_ready);
                a_par ^= 1;
            }
        }
    } else {
        // ======================================= compute warps ===============================================
        const int row = tid & (TN - 1);   // test row inside the tile == TMEM lane
        const int h = tid >> 7;           // which half of the 16-byte chunks (4 training points each) this thread does
        const int sw = row & 7;           // swizzle phase of the row
        uint32_t accR_par = 0;
        mbar_wait(x_bar, 0);
        const float* al = Xs + Mp * DP;
        const int nchunk = nslab * 8;   // 16-byte chunks (4 training points) covering the K slabs

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

            float mu = 0.f;
            f32x2 g2[DP / 2];
#pragma unroll
            for (int q = 0; q < DP / 2; ++q) g2[q] = 0ull;

            // phase A: chunk c = 4 consecutive training points = one 16-byte unit of the row's 128-byte swizzle atom.
            // All per-dimension arithmetic runs on packed pairs of dimensions (FADD2 / FFMA2).
            for (int c = h; c < nchunk; c += 2) {
                float k4[4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int j = 4 * c + q4;
                    const ulonglong2* xr = reinterpret_cast<const ulonglong2*>(Xs + j * DP);   // float4 = 2 packed pairs
                    f32x2 u2[DP / 2];
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
                    const float k = ex2_approx((r_lo + r_hi) * -0.72134752044448170368f);   // exp(-r2/2) = 2^(-r2 log2(e)/2)
                    k4[q4] = k;
                    const float cj = k * al[j];
                    mu += cj;
                    const f32x2 cj2 = pack2(cj, cj);
#pragma unroll
                    for (int q = 0; q < DP / 2; ++q) g2[q] = fma2(cj2, u2[q], g2[q]);
                }
                if (want_var) {
                    const int slab = c >> 3, cc = c & 7;
                    uint4 v;
                    v.x = tf32_rna(k4[0]); v.y = tf32_rna(k4[1]); v.z = tf32_rna(k4[2]); v.w = tf32_rna(k4[3]);
                    *reinterpret_cast<uint4*>(At + (size_t)slab * (TN * 128) + row * 128 + ((cc ^ sw) << 4)) = v;
                    if (X3) {   // lo = rna(k - hi) -> this row's TMEM lane, columns 256 + 4 c .. + 3
                        tmem_st4(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + 256u + (uint32_t)(4 * c),
                                 tf32_rna(k4[0] - __uint_as_float(v.x)), tf32_rna(k4[1] - __uint_as_float(v.y)),
                                 tf32_rna(k4[2] - __uint_as_float(v.z)), tf32_rna(k4[3] - __uint_as_float(v.w)));
                    }
                    if (cc >= 6) {   // this thread's last chunk of the slab (cc == 6 for h == 0, 7 for h == 1)
                        if (X3) {
                            tmem_st_wait();
                            tc_fence_before();
                        }
                        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
                        mbar_arrive(&a_ready[slab]);
                    }
                }
            }

            // combine the two halves of every row (half 1 parks its sums in smem, half 0 adds its own), then write
            {
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
            }
            if (p.mu != nullptr && tid < npts) p.mu[(n0 + tid) * p.ld_mu] = outs[tid * DV];
            if (p.deriv != nullptr) {
                for (int e = tid; e < npts * D; e += NC) {
                    const int r = e / D, d = e - r * D;
                    p.deriv[(n0 + r) * p.ld_deriv + d] = p.sqrt_w[d] * outs[r * DV + 1 + d];
                }
            }

            if (want_var) {
                // epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (its rows) and column half w / 4
                mbar_wait(acc_ready, accR_par);
                accR_par ^= 1;
                tc_fence_after();
                const int q = warp & 3, ch = warp >> 2;
                const int erow = 32 * q + lane;
                const int half_cols = Mp / 2;          // Mp is a multiple of 64: whole 32-column slabs per half
                float vsum = 0.f;
                for (int c0 = ch * half_cols; c0 < (ch + 1) * half_cols; c0 += 32) {
                    if ((c0 >> 5) >= nslab) break;   // columns >= 32 * nslab: G is zero there and K* was never written
                    float gv[32];
                    tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, gv);
                    const unsigned char* arow = At + (size_t)(c0 >> 5) * (TN * 128) + erow * 128;
                    float lo[32];
                    if (X3) tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + 256u + (uint32_t)c0, lo);
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        float4 kv = *reinterpret_cast<const float4*>(arow + ((cc ^ (erow & 7)) << 4));
                        if (X3) {   // K* = hi + lo, exact to 2^-22
                            kv.x += lo[4 * cc + 0]; kv.y += lo[4 * cc + 1]; kv.z += lo[4 * cc + 2]; kv.w += lo[4 * cc + 3];
                        }
                        vsum = fmaf(gv[4 * cc + 0], kv.x, vsum);
                        vsum = fmaf(gv[4 * cc + 1], kv.y, vsum);
                        vsum = fmaf(gv[4 * cc + 2], kv.z, vsum);
                        vsum = fmaf(gv[4 * cc + 3], kv.w, vsum);
                    }
                }
                vred[ch * TN + erow] = vsum;
                tc_fence_before();
                mbar_arrive(acc_empty);   // this thread's TMEM reads are done
                bar_sync_compute();
                if (tid < npts) p.var[(n0 + tid) * p.ld_var] = p.b - p.b * p.b * (vred[tid] + vred[TN + tid]);
            }
            bar_sync_compute();   // outs / vred / K* tile free for the next tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTfComputeWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(kTmemCols));
    }
}

}  // namespace gpe
