// Nearest-neighbour resize of channels-last maps: the F.interpolate(x, size=shape) in front of every
// decoder up-convolution (utils/net_utils.py:149-190 UpConv2d, called from MultiScaleDecoder,
// RCNet/networks.py:678-779).  Pure copy, bit-exact with torch's nearest mode: source index =
// min(int(floorf(dst * float(in) / float(out))), in - 1).  It sits between the gather kernel's
// pooled skips and the cuDNN convolutions; ATen's NHWC nearest kernel runs at ~0.7 TB/s on the
// (points x C x 240 x 100) activations and was 29 % of the step, this one streams float4 over channels.
#include "common.cuh"

namespace riders {

constexpr int UP_THREADS = 256;

__global__ void __launch_bounds__(UP_THREADS)
upsample_nearest_nhwc_kernel(const float4* __restrict__ x, float4* __restrict__ out,
                             int C4, int h, int w, int H, int W, float scale_h, float scale_w, long long rows) {
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {      // row = n * H + oh
        const int oh = (int)(row % H);
        const long long n = row / H;
        const int sh = min((int)floorf(__fmul_rn((float)oh, scale_h)), h - 1);
        const float4* src = x + ((size_t)n * h + sh) * (size_t)w * C4;
        float4* dst = out + (size_t)row * W * C4;
        const int len = W * C4;
        for (int e = threadIdx.x; e < len; e += UP_THREADS) {
            const int ow = e / C4, c = e - ow * C4;
            const int sw = min((int)floorf(__fmul_rn((float)ow, scale_w)), w - 1);
            __stcs(dst + e, __ldg(src + (size_t)sw * C4 + c));
        }
    }
}

}  // namespace riders

extern "C" int riders_upsample_nearest_nhwc(const float* x, int N, int C, int h, int w, int H, int W,
                                            float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "upsample_nearest: bad shape");
    RIDERS_REQUIRE((C & 3) == 0, "upsample_nearest: C %% 4 != 0 (C=%d)", C);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && out, "upsample_nearest: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0, "upsample_nearest: maps must be 16-byte aligned");
    const long long rows = (long long)N * H;
    const long long want = rows < (long long)sm_count() * 32 ? rows : (long long)sm_count() * 32;
    count_launches(1);
    upsample_nearest_nhwc_kernel<<<(unsigned)want, UP_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), C / 4, h, w, H, W,
        (float)h / (float)H, (float)w / (float)W, rows);
    return check_launch("upsample_nearest_nhwc");
}

// ---------------------------------------------------------------------------------------------
// out[p][0:Ca] = leaky_relu(a[p][:]), out[p][Ca:Ca+Cb] = b[p][:]  for every pixel p (channels-last).
// One pass instead of DecoderBlock's in-place LeakyReLU on the up-convolution output followed by
// torch.cat([y, skip], dim=1) (utils/net_utils.py:474-569): reads y and skip once, writes the
// concatenated map once.  Values are bit-identical (x > 0 ? x : x * slope, then a copy).
// ---------------------------------------------------------------------------------------------
namespace riders {

__global__ void __launch_bounds__(256)
leaky_concat_nhwc_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                         int Ca4, int Cb4, long long pixels, float slope, int act_on_a) {
    const int Ct4 = Ca4 + Cb4;
    const long long total = pixels * Ct4;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long p = e / Ct4;
        const int c = (int)(e - p * Ct4);
        float4 v;
        if (c < Ca4) {
            v = __ldcs(a + p * Ca4 + c);
            if (act_on_a) {
                v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
                v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
            }
        } else {
            v = __ldcs(b + p * Cb4 + (c - Ca4));
        }
        out[e] = v;                                   // consumed right away by the next convolution: keep it in L2
    }
}

}  // namespace riders

extern "C" int riders_leaky_concat_nhwc(const float* a, int Ca, const float* b, int Cb, long long pixels,
                                        float slope, int act_on_a, float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(Ca > 0 && Cb >= 0 && pixels >= 0, "leaky_concat: bad shape");
    RIDERS_REQUIRE((Ca & 3) == 0 && (Cb & 3) == 0, "leaky_concat: channel counts must be multiples of 4 (%d, %d)", Ca, Cb);
    if (pixels == 0) return RIDERS_OK;
    RIDERS_REQUIRE(a && out && (b || Cb == 0), "leaky_concat: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0, "leaky_concat: maps must be 16-byte aligned");
    const long long total = pixels * ((Ca + Cb) / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    count_launches(1);
    leaky_concat_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out),
        Ca / 4, Cb / 4, pixels, slope, act_on_a);
    return check_launch("leaky_concat_nhwc");
}
