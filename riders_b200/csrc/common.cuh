// Shared helpers for the riders_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/riders_b200.h"

namespace riders {

void set_error(const char* fmt, ...);
int  check_launch(const char* what);          // cudaGetLastError -> RIDERS_* code
int  sm_count();
void set_reduce_rows(int rows);
void set_reduce_tile_w(int w);               // confidence_reduce.cu: 16 (default) or 32 pixel wide tiles              // confidence_reduce.cu tuning knob (4 or 8 rows per thread)
void set_lsq_mode(int mode);                 // lsq_align.cu: 2 = PDL (default), 3 = one launch + flag barrier, 1 = cooperative, 0 = two kernels
void set_gather_smem_budget(int kb);         // roi_gather.cu tuning knob (riders_debug_set)
void set_tail_tc_dbg(int v);                 // decoder_tail_tc.cu timing experiments
void count_launches(int n);                   // feeds riders_launch_count() (bench.py's gpu_launches)

#define RIDERS_REQUIRE(cond, ...)                         \
    do { if (!(cond)) { riders::set_error(__VA_ARGS__); return RIDERS_EINVAL; } } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel into `bytes` of dynamic shared memory on the CURRENT device.  The attribute is per device (and per
// context), so the "already done" cache is indexed by device: a second GPU in the same process gets its own opt-in.
template <typename Kernel>
inline int ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t (&cache)[64], const char* what) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (bytes <= cache[dev]) return RIDERS_OK;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return check_launch(what);
    cache[dev] = bytes;
    return RIDERS_OK;
}

__device__ __forceinline__ float ld_stream(const float* p) {
    // read-once data: bypass L1 allocation
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

}  // namespace riders
