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

// ---------------------------------------------------------------------------------------------
// y = act(x + bias[c]) in one pass over a channels-last map (in place when out == x): the bias of a
// convolution whose eval-mode BatchNorm was folded into it, plus utils/net_utils.py's activation
// (LeakyReLU 0.2 everywhere in RC-Net).  ATen would run a broadcast add and then leaky_relu_.
// slope: 0.2 leaky, 0 relu, 1 identity.  bias may be NULL.
// ---------------------------------------------------------------------------------------------
namespace riders {

__global__ void __launch_bounds__(256)
bias_act_nhwc_kernel(const float4* __restrict__ x, const float* __restrict__ bias, float4* __restrict__ out,
                     int C4, long long total4, float slope) {
    // blockDim (256) and the grid stride are multiples of C4 (a power of two <= 256): the channel group is fixed per thread
    const int c = threadIdx.x & (C4 - 1);
    const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total4; e += (long long)gridDim.x * 256) {
        float4 v = x[e];
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        out[e] = v;
    }
}

__global__ void __launch_bounds__(256)
bias_act_generic_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ out,
                        int C, long long total, float slope) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        float v = x[e] + (bias ? bias[e % C] : 0.f);
        out[e] = v > 0.f ? v : v * slope;
    }
}

__global__ void __launch_bounds__(256)
bias_leaky_concat_nhwc_kernel(const float4* __restrict__ a, const float* __restrict__ bias, const float4* __restrict__ b,
                              float4* __restrict__ out, int Ca4, int Cb4, int ld4, long long pixels, float slope) {
    const int Ct4 = Ca4 + Cb4;
    const long long total = pixels * Ct4;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long p = e / Ct4;
        const int c = (int)(e - p * Ct4);
        float4 v;
        if (c < Ca4) {
            v = __ldcs(a + p * Ca4 + c);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        } else {
            v = __ldcs(b + p * Cb4 + (c - Ca4));
        }
        out[p * ld4 + c] = v;                       // ld4 == Ct4: dense; larger: the skip channels are already in place
    }
}

}  // namespace riders

extern "C" int riders_bias_act_nhwc(const float* x, const float* bias, int C, long long pixels, float slope,
                                    float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(C > 0 && pixels >= 0, "bias_act: bad shape");
    if (pixels == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && out, "bias_act: null pointer");
    const long long cap = (long long)sm_count() * 16;
    count_launches(1);
    const int C4 = C / 4;
    if ((C & 3) == 0 && C4 <= 256 && (C4 & (C4 - 1)) == 0 && ((((uintptr_t)x) | ((uintptr_t)out) | ((uintptr_t)bias)) & 15) == 0) {
        const long long total4 = pixels * C4;
        long long blocks = (total4 + 255) / 256;
        if (blocks > cap) blocks = cap;
        bias_act_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(x), bias, reinterpret_cast<float4*>(out), C4, total4, slope);
    } else {
        const long long total = pixels * C;
        long long blocks = (total + 255) / 256;
        if (blocks > cap) blocks = cap;
        bias_act_generic_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, out, C, total, slope);
    }
    return check_launch("bias_act_nhwc");
}

extern "C" int riders_bias_leaky_concat_nhwc(const float* a, const float* bias, int Ca, const float* b, int Cb,
                                             long long pixels, float slope, float* out, void* stream) {
    return riders_bias_leaky_place_nhwc(a, bias, Ca, b, Cb, pixels, slope, out, Ca + Cb, stream);
}

extern "C" int riders_bias_leaky_place_nhwc(const float* a, const float* bias, int Ca, const float* b, int Cb,
                                            long long pixels, float slope, float* out, int out_pixel_stride, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(out_pixel_stride >= Ca + Cb && (out_pixel_stride & 3) == 0, "bias_leaky_place: out_pixel_stride=%d < %d channels or not %%4", out_pixel_stride, Ca + Cb);
    RIDERS_REQUIRE(Ca > 0 && Cb >= 0 && pixels >= 0, "bias_leaky_concat: bad shape");
    RIDERS_REQUIRE((Ca & 3) == 0 && (Cb & 3) == 0, "bias_leaky_concat: channel counts must be multiples of 4 (%d, %d)", Ca, Cb);
    if (pixels == 0) return RIDERS_OK;
    RIDERS_REQUIRE(a && bias && out && (b || Cb == 0), "bias_leaky_concat: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out) | ((uintptr_t)bias)) & 15) == 0, "bias_leaky_concat: 16-byte alignment");
    const long long total = pixels * ((Ca + Cb) / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    count_launches(1);
    bias_leaky_concat_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(a), bias, reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out),
        Ca / 4, Cb / 4, out_pixel_stride / 4, pixels, slope);
    return check_launch("bias_leaky_concat_nhwc");
}

// ---------------------------------------------------------------------------------------------
// Sub-pixel form of "nearest 2x resize -> 3x3 convolution" (decoder blocks whose target size is exactly
// twice the input: 15x6 -> 30x12, 60x25 -> 120x50, 120x50 -> 240x100 in RC-Net).  With U[r][c] = x[r>>1][c>>1]
// the 3x3 convolution of U at output pixel (2i+a, 2j+b) only touches x rows {i-1, i, i+1}: it equals a 3x3
// convolution of x itself with the row taps [W0, W1+W2, 0] (a = 0) or [0, W0+W1, W2] (a = 1), same for columns.
// So ONE convolution of the low-resolution map with 4*Cout phase-major output channels (built once by
// DecoderBlock.fold_upsample_) replaces resize + convolution: the 4x larger resized map is never written or read.
// This kernel is the consumer: out[n][2i+a][2j+b][0:Ca] = act(y4[n][i][j][(2a+b)*Ca + c] + bias[c]), then the skip
// channels (if any) are appended - depth-to-space + bias + LeakyReLU + concat in one pass.
// ---------------------------------------------------------------------------------------------
namespace riders {

__global__ void __launch_bounds__(256)
d2s_bias_act_concat_kernel(const float4* __restrict__ y4, const float* __restrict__ bias, const float4* __restrict__ skip,
                           float4* __restrict__ out, int Ca4, int Cb4, int ld4, int ct_shift, int h, int w, int t, float slope,
                           long long rows, int rows_per_block) {
    // t = 0: y4 is (N, h, w, 4*Ca) from the 3x3 phase kernels; t = 1: (N, h+1, w+1, 4*Ca) from the 2x2 phase kernels with
    // padding 1, where phase (a, b) of low-resolution pixel (i, j) sits at (i + a, j + b)
    // a block streams `rows_per_block` consecutive output rows as one flat run of float4s: (row, position) advance
    // incrementally, the channel split uses a shift when Ct4 is a power of two (RC-Net: 4, 16, 64)
    const int Ct4 = Ca4 + Cb4, W2 = 2 * w, H2 = 2 * h;
    const int len = W2 * Ct4;
    const long long row0 = (long long)blockIdx.x * rows_per_block;
    const int nrows = (int)min((long long)rows_per_block, rows - row0);
    // four elements per thread and iteration: all loads are issued before the first store (latency bound otherwise)
    int r = 0, pos = threadIdx.x;
    while (pos >= len) { pos -= len; ++r; }
    while (r < nrows) {
        float4 v[4];
        float4* dstp[4];
        bool act[4];
        int cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dstp[u] = nullptr;
            if (r < nrows) {
                const unsigned row = (unsigned)(row0 + r);                    // row = n * H2 + Y (rows < 2^31)
                const unsigned n = row / (unsigned)H2, Y = row - n * (unsigned)H2;
                const int X = ct_shift >= 0 ? (pos >> ct_shift) : (pos / Ct4);
                const int c = pos - X * Ct4;
                act[u] = c < Ca4;
                cc[u] = c;
                v[u] = act[u] ? __ldcs(y4 + (((size_t)n * (h + t) + (Y >> 1) + (Y & 1) * t) * (size_t)(w + t) + (X >> 1) + (X & 1) * t) * (4 * Ca4)
                                          + (2 * (Y & 1) + (X & 1)) * Ca4 + c)
                              : __ldcs(skip + ((size_t)row * W2 + X) * Cb4 + (c - Ca4));
                dstp[u] = out + ((size_t)row * W2 + X) * ld4 + c;      // ld4 == Ct4: dense; larger: the skip channels are already in place
                pos += 256;
                while (pos >= len) { pos -= len; ++r; }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!dstp[u]) continue;
            float4 t = v[u];
            if (act[u]) {
                if (bias) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cc[u]);
                    t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w;
                }
                t.x = t.x > 0.f ? t.x : t.x * slope; t.y = t.y > 0.f ? t.y : t.y * slope;
                t.z = t.z > 0.f ? t.z : t.z * slope; t.w = t.w > 0.f ? t.w : t.w * slope;
            }
            *dstp[u] = t;
        }
    }
}

}  // namespace riders

extern "C" int riders_d2s_bias_act_concat_nhwc(const float* y4, const float* bias, int Ca, const float* skip, int Cb,
                                               int N, int h, int w, float slope, float* out, void* stream) {
    return riders_d2s_bias_act_place_nhwc(y4, bias, Ca, skip, Cb, N, h, w, slope, out, Ca + Cb, stream);
}

extern "C" int riders_d2s_bias_act_place_nhwc(const float* y4, const float* bias, int Ca, const float* skip, int Cb,
                                              int N, int h, int w, float slope, float* out, int out_pixel_stride, void* stream) {
    return riders_subpixel_d2s_place_nhwc(y4, 3, bias, Ca, skip, Cb, N, h, w, slope, out, out_pixel_stride, stream);
}

extern "C" int riders_subpixel_d2s_place_nhwc(const float* y4, int taps, const float* bias, int Ca, const float* skip, int Cb,
                                              int N, int h, int w, float slope, float* out, int out_pixel_stride, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(taps == 2 || taps == 3, "subpixel_d2s: taps must be 2 or 3 (got %d)", taps);
    RIDERS_REQUIRE(out_pixel_stride >= Ca + Cb && (out_pixel_stride & 3) == 0, "d2s_bias_act_place: out_pixel_stride=%d < %d channels or not %%4", out_pixel_stride, Ca + Cb);
    RIDERS_REQUIRE(N >= 0 && Ca > 0 && Cb >= 0 && h > 0 && w > 0, "d2s_bias_act_concat: bad shape");
    RIDERS_REQUIRE((Ca & 3) == 0 && (Cb & 3) == 0, "d2s_bias_act_concat: channel counts must be multiples of 4 (%d, %d)", Ca, Cb);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(y4 && out && (skip || Cb == 0), "d2s_bias_act_concat: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)y4) | ((uintptr_t)skip) | ((uintptr_t)out) | ((uintptr_t)bias)) & 15) == 0, "d2s_bias_act_concat: 16-byte alignment");
    const long long rows = (long long)N * 2 * h;
    const int Ct4 = (Ca + Cb) / 4, len = 2 * w * Ct4;
    int ct_shift = -1;
    if ((Ct4 & (Ct4 - 1)) == 0) { ct_shift = 0; while ((1 << ct_shift) < Ct4) ++ct_shift; }
    const int rows_per_block = len >= 8192 ? 1 : (8192 + len - 1) / len;
    const long long blocks = (rows + rows_per_block - 1) / rows_per_block;
    RIDERS_REQUIRE(rows < (1ll << 31), "d2s_bias_act_concat: too many rows for one launch");
    count_launches(1);
    d2s_bias_act_concat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(y4), bias, reinterpret_cast<const float4*>(skip), reinterpret_cast<float4*>(out),
        Ca / 4, Cb / 4, out_pixel_stride / 4, ct_shift, h, w, taps == 2 ? 1 : 0, slope, rows, rows_per_block);
    return check_launch("d2s_bias_act_concat_nhwc");
}


// ---------------------------------------------------------------------------------------------
// The same sub-pixel trick for decoder blocks whose target is 2n + 1 along ONE axis (RC-Net: 30x12 -> 60x25 and
// 7x3 -> 15x6).  torch's nearest resize to 2n + 1 picks source index max(j - 1, 0) >> 1: the resized map is the exact
// 2x upsampling of x' = replicate-pad(x, 1 on the left / top) without its first column / row.  So for X >= 1 (Y >= 1)
// the 3x3 convolution of the resized map is the phase convolution of x' read one output pixel further (k = X + 1);
// only the first column (row), whose left (top) neighbour is the zero padding instead of another copy of x[0], differs:
// it comes from a 3x1 (1x3) phase convolution of x's first column (row) with the kx (ky) in {1, 2} taps pre-summed.
//   y4     (N, 4*Ca, hp, wp)  phase convolution of x' (hp = h + sy, wp = w + sx), channel (2a+b)*Ca + c
//   yfirst sx: (N, 2*Ca, h, 1) channel a*Ca + c  |  sy: (N, 2*Ca, 1, w) channel b*Ca + c
//   out    (N, ld, 2*hp - sy, 2*wp - sx): channels [0, Ca) = act(. + bias), pixel stride ld (skip already in place)
// ---------------------------------------------------------------------------------------------
namespace riders {

__global__ void __launch_bounds__(256)
d2s_shift_bias_act_kernel(const float4* __restrict__ y4, const float4* __restrict__ yfirst, const float* __restrict__ bias,
                          float4* __restrict__ out, int Ca4, int ld4, int hp, int wp, int sy, int sx, int tp, float slope,
                          long long total) {
    const int H = 2 * hp - sy, W = 2 * wp - sx;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % Ca4);
        long long t = e / Ca4;
        const int X = (int)(t % W); t /= W;
        const int Y = (int)(t % H);
        const long long n = t / H;
        float4 v;
        if (sx && X == 0) {
            v = __ldcs(yfirst + (n * (hp - sy) + (Y >> 1)) * (2 * Ca4) + (Y & 1) * Ca4 + c);
        } else if (sy && Y == 0) {
            v = __ldcs(yfirst + (n * (wp - sx) + (X >> 1)) * (2 * Ca4) + (X & 1) * Ca4 + c);
        } else {
            const int ky = Y + sy, kx = X + sx;
            v = __ldcs(y4 + ((n * (hp + tp) + (ky >> 1) + (ky & 1) * tp) * (long long)(wp + tp) + (kx >> 1) + (kx & 1) * tp) * (4 * Ca4)
                       + (2 * (ky & 1) + (kx & 1)) * Ca4 + c);
        }
        if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        out[((n * H + Y) * (long long)W + X) * ld4 + c] = v;
    }
}

}  // namespace riders

extern "C" int riders_d2s_shift_bias_act_place_nhwc(const float* y4, const float* yfirst, const float* bias, int Ca,
                                                    int N, int hp, int wp, int sy, int sx, float slope,
                                                    float* out, int out_pixel_stride, void* stream) {
    return riders_subpixel_d2s_shift_place_nhwc(y4, 3, yfirst, bias, Ca, N, hp, wp, sy, sx, slope, out, out_pixel_stride, stream);
}

extern "C" int riders_subpixel_d2s_shift_place_nhwc(const float* y4, int taps, const float* yfirst, const float* bias, int Ca,
                                                    int N, int hp, int wp, int sy, int sx, float slope,
                                                    float* out, int out_pixel_stride, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(taps == 2 || taps == 3, "subpixel_d2s_shift: taps must be 2 or 3 (got %d)", taps);
    RIDERS_REQUIRE(N >= 0 && Ca > 0 && (Ca & 3) == 0 && hp > 1 && wp > 1, "d2s_shift: bad shape");
    RIDERS_REQUIRE((sy == 0 || sy == 1) && (sx == 0 || sx == 1) && sy + sx == 1, "d2s_shift: exactly one of sy, sx must be 1");
    RIDERS_REQUIRE(out_pixel_stride >= Ca && (out_pixel_stride & 3) == 0, "d2s_shift: out_pixel_stride=%d < %d channels or not %%4", out_pixel_stride, Ca);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(y4 && yfirst && out, "d2s_shift: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)y4) | ((uintptr_t)yfirst) | ((uintptr_t)out) | ((uintptr_t)bias)) & 15) == 0, "d2s_shift: 16-byte alignment");
    const long long total = (long long)N * (2 * hp - sy) * (2 * wp - sx) * (Ca / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    count_launches(1);
    d2s_shift_bias_act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(y4), reinterpret_cast<const float4*>(yfirst), bias, reinterpret_cast<float4*>(out),
        Ca / 4, out_pixel_stride / 4, hp, wp, sy, sx, taps == 2 ? 1 : 0, slope, total);
    return check_launch("d2s_shift_bias_act_place_nhwc");
}


// ---------------------------------------------------------------------------------------------
// act(x + bias) written into a map that is replicate-padded by one column on the left (sx) or one row on top (sy): the
// operand x' of the 2n+1 sub-pixel block above, produced by the pass that applies the previous convolution's bias and
// LeakyReLU anyway (instead of F.pad + a layout copy of the activated map: two more passes over it).
//   x (N, h, w, C) channels-last raw convolution output -> out (N, h + sy, w + sx, C); out[:, sy:, sx:] is act(x + bias).
// ---------------------------------------------------------------------------------------------
namespace riders {

__global__ void __launch_bounds__(256)
bias_act_pad_nhwc_kernel(const float4* __restrict__ x, const float* __restrict__ bias, float4* __restrict__ out,
                         int C4, int h, int w, int sy, int sx, float slope, long long total) {
    const int H = h + sy, W = w + sx;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % C4);
        long long t = e / C4;
        const int X = (int)(t % W); t /= W;
        const int Y = (int)(t % H);
        const long long n = t / H;
        float4 v = __ldg(x + ((n * h + max(Y - sy, 0)) * (long long)w + max(X - sx, 0)) * C4 + c);
        if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        out[e] = v;
    }
}

}  // namespace riders

extern "C" int riders_bias_act_pad_nhwc(const float* x, const float* bias, int C, int N, int h, int w, int sy, int sx,
                                        float slope, float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && C > 0 && (C & 3) == 0 && h > 0 && w > 0, "bias_act_pad: bad shape");
    RIDERS_REQUIRE((sy == 0 || sy == 1) && (sx == 0 || sx == 1), "bias_act_pad: sy, sx must be 0 or 1");
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && out && x != out, "bias_act_pad: null pointer or in-place call");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)out) | ((uintptr_t)bias)) & 15) == 0, "bias_act_pad: 16-byte alignment");
    const long long total = (long long)N * (h + sy) * (w + sx) * (C / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    count_launches(1);
    bias_act_pad_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), bias, reinterpret_cast<float4*>(out), C / 4, h, w, sy, sx, slope, total);
    return check_launch("bias_act_pad_nhwc");
}
