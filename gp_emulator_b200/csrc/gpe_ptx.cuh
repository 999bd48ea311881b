// Thin inline-PTX wrappers: mbarrier, TMA bulk copies (cp.async.bulk -> SASS UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}

// Make mbarrier.init visible to the async (TMA) proxy and to the other CTAs of a cluster.
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D TMA bulk copy global -> this CTA's shared memory, completion counted in bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Kernel-entry guard (always on; a handful of integer instructions per CTA): the shared-memory carve-up computed on the
// host -- byte offsets travelling in the kernel parameters -- must fit the dynamic shared memory this launch was actually
// given.  A mismatch traps (the launch fails with an error the host reports) instead of silently corrupting a
// neighbouring buffer.  compute-sanitizer is not available on the GPU pool; this and the red-zone test
// (tests/test_gpu_parity.py::test_red_zones_around_every_output) stand in for it.
__device__ __forceinline__ uint32_t dynamic_smem_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void smem_guard(uint32_t extent) {
    if (threadIdx.x == 0 && extent > dynamic_smem_size()) __trap();
}
__device__ __forceinline__ uint32_t umax2(uint32_t a, uint32_t b) { return a > b ? a : b; }

}  // namespace gpe
