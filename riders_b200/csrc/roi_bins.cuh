// RoIPool bin arithmetic shared by the gather kernels (A1).  Restates torchvision's RoIPool
// rules (third-party; the reference calls it from RCNet/networks.py:418-433): float32 products,
// C round() (half away from zero), +1 inclusive extents, float32 bin sizes, floor/ceil edges,
// clipping to the map.  __host__ __device__ so the CPU unit test can execute the same code.
#pragma once
#include <math.h>

namespace riders {

struct RoiAxis {
    int start;        // quantised RoI start on this level
    int extent;       // max(end - start + 1, 1)
    float bin;        // float32(extent) / float32(pooled)
};

__host__ __device__ inline RoiAxis roi_axis(float lo, float hi, float scale, int pooled) {
    RoiAxis a;
#ifdef __CUDA_ARCH__
    a.start = (int)roundf(__fmul_rn(lo, scale));
    const int end = (int)roundf(__fmul_rn(hi, scale));
    a.extent = max(end - a.start + 1, 1);
    a.bin = __fdiv_rn((float)a.extent, (float)pooled);
#else
    volatile float l = lo * scale, h = hi * scale;      // volatile: keep float32 roundings on the host
    a.start = (int)roundf(l);
    const int end = (int)roundf(h);
    a.extent = end - a.start + 1 > 1 ? end - a.start + 1 : 1;
    volatile float b = (float)a.extent / (float)pooled;
    a.bin = b;
#endif
    return a;
}

// clipped [b0, b1) input range of pooled bin i
__host__ __device__ inline void roi_bin_range(const RoiAxis& a, int i, int limit, int& b0, int& b1) {
#ifdef __CUDA_ARCH__
    b0 = (int)floorf(__fmul_rn((float)i, a.bin)) + a.start;
    b1 = (int)ceilf(__fmul_rn((float)(i + 1), a.bin)) + a.start;
#else
    volatile float f0 = (float)i * a.bin, f1 = (float)(i + 1) * a.bin;
    b0 = (int)floorf(f0) + a.start;
    b1 = (int)ceilf(f1) + a.start;
#endif
    b0 = b0 < 0 ? 0 : (b0 > limit ? limit : b0);
    b1 = b1 < 0 ? 0 : (b1 > limit ? limit : b1);
}

}  // namespace riders
