// Host-side helpers shared by the translation units of libgpemu.so (defined in gpemu.cu).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

namespace gpe {

// Record the calling thread's error message (read back by gpe_last_error) and return `code`.
int set_error(int code, const char* fmt, ...);
// GPE_OK if `device` is a usable sm_100 device (its SM count goes to *sms), else a status with the message set.
int require_device(int device, int* sms);
// Count one kernel launch (gpe_launch_count).
void count_launch();

// NVTX range around a host-side stage (header-only NVTX 3: a no-op costing nanoseconds unless a profiler is attached;
// `ncu --nvtx` / Nsight Systems then show the library's stages by name).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

}  // namespace gpe
