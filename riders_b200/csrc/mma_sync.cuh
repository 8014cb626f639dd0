// Warp-level TF32 tensor-core helpers shared by decoder_tail.cu and subpixel_conv.cu (sm_100a: HMMA.1688.F32.TF32,
// measured 2.0 cycles per instruction per SM = 512 MAC/cycle/SM, 20 cycles dependent latency).
#pragma once
#include "common.cuh"

namespace riders {
namespace tail {

// cvt.rna.tf32.f32 (round to nearest, ties away) for finite values in ONE integer add: the MMA ignores the low 13
// mantissa bits, so adding half a TF32 ulp to the magnitude bits and letting the hardware truncate is the same rounding.
// (ptxas expands the cvt itself into an |x| < inf test plus this add: ~5 instructions per value.)
__device__ __forceinline__ uint32_t to_tf32(float v) { return __float_as_uint(v) + 0x1000u; }

// 16-byte streaming load under a predicate (no branch: the loads of a batch issue back to back), zeros when off
__device__ __forceinline__ float4 ldg128_if(const float4* p, bool on) {
    float4 v;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %5, 0;\n"
        "mov.f32 %0, 0f00000000;\n mov.f32 %1, 0f00000000;\n mov.f32 %2, 0f00000000;\n mov.f32 %3, 0f00000000;\n"
        "@q ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n"
        "}\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "r"((int)on));
    return v;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&a)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace tail
}  // namespace riders
