// Decoder head feeding the reduce kernel: logits = conv3x3( act(x + bias) ), C -> 1 channel, zero padding,
// channels-last input.  In the reference this is MultiScaleDecoder.output0 (RCNet/networks.py:678-779:
// Conv2d(n_filters[-1] -> 1, 3x3, no bias, linear)) applied to the LeakyReLU'd output of the last decoder
// convolution.  cuDNN runs the 16 -> 1 convolution with a padded output plus an NHWC->NCHW conversion
// (3.1 ms per 2048 patches) after a separate bias/activation pass over the 16-channel 240x100 maps; here one
// kernel reads the raw 16-channel map once (64 B/pixel), applies bias + activation in registers, and writes
// 4 B/pixel.  HBM roofline: (4 C + 4) bytes per pixel.  fp32 FFMA (more exact than the TF32 path it replaces).
//
// Mapping: a warp owns a vertical strip of S <= 30 output columns (+1 halo lane on each side) and slides
// down the rows keeping three running row sums; the left / right taps come from the neighbouring lanes by
// shuffle (as 6 partial sums per row), so every input value is loaded and activated exactly once per strip.  Weights and bias travel as
// kernel parameters (constant bank operands of the FFMAs).
#include "common.cuh"

namespace riders {

constexpr int HEAD_C = 16;
constexpr int HEAD_WARPS = 4;

struct HeadParams {
    float w[3][3][HEAD_C];      // [ky][kx][c]
    float bias[HEAD_C];
    float slope;
    int sigmoid;                // write torch.sigmoid(logit) (rcnet_model.py:123) instead of the logit
};

__global__ void __launch_bounds__(HEAD_WARPS * 32)
conv3x3_head_kernel(const float4* __restrict__ x, float* __restrict__ out, const __grid_constant__ HeadParams P,
                    int H, int W, int S, int strips, int rows_per_seg, int segs, long long units) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * HEAD_WARPS + warp;   // (patch, row segment, strip); warp-uniform
    if (unit >= units) return;
    const int strip = (int)(unit % strips);
    const int seg = (int)((unit / strips) % segs);
    const long long n = unit / (strips * segs);
    const int r0 = seg * rows_per_seg, r1 = min(H, r0 + rows_per_seg);
    const int col = strip * S - 1 + lane;                            // lane 0 / S+1 are halo columns
    const bool col_ok = lane <= S + 1 && col >= 0 && col < W;
    const bool writes = lane >= 1 && lane <= S && col < W;
    const float4* xin = x + ((size_t)n * H * W) * (HEAD_C / 4);
    float* o = out + (size_t)n * H * W;

    float acc_prev = 0.f, acc_cur = 0.f;                             // sums of output rows h-1 and h
    // software pipeline: the 64 B of row h+1 are in flight while row h is activated and accumulated
    float4 nxt[HEAD_C / 4];
    auto fetch = [&](int h) {
        const bool ok = col_ok && h >= 0 && h < H;
        const float4* p = xin + ((size_t)(ok ? h : 0) * W + (ok ? col : 0)) * (HEAD_C / 4);
#pragma unroll
        for (int q = 0; q < HEAD_C / 4; ++q) nxt[q] = ok ? __ldcs(p + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        return ok;
    };
    bool ok_nxt = fetch(r0 - 1);
    for (int h = r0 - 1; h <= r1; ++h) {
        float a[HEAD_C];
        const bool ok = ok_nxt;
#pragma unroll
        for (int q = 0; q < HEAD_C / 4; ++q) {
            a[4 * q + 0] = nxt[q].x + P.bias[4 * q + 0]; a[4 * q + 1] = nxt[q].y + P.bias[4 * q + 1];
            a[4 * q + 2] = nxt[q].z + P.bias[4 * q + 2]; a[4 * q + 3] = nxt[q].w + P.bias[4 * q + 3];
        }
        if (h < r1) ok_nxt = fetch(h + 1);
#pragma unroll
        for (int c = 0; c < HEAD_C; ++c) {
            const float v = a[c] > 0.f ? a[c] : a[c] * P.slope;
            a[c] = ok ? v : 0.f;                                     // zero padding of the ACTIVATED map
        }
        // D[ky][kx] = sum_c w[ky][kx][c] * a[c]: what THIS input pixel adds to the output pixel one step to the
        // right (kx = 0), itself (kx = 1) and one step to the left (kx = 2).  Exchanging the 6 off-centre partial
        // sums instead of the 16 activated inputs cuts the shuffles per row from 32 to 6 (the MIO pipe was the limiter).
        float d[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) d[ky][kx] = 0.f;
#pragma unroll
        for (int c = 0; c < HEAD_C; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) d[ky][kx] = fmaf(P.w[ky][kx][c], a[c], d[ky][kx]);
        // output column x sums D_{x-1}[ky][0] + D_x[ky][1] + D_{x+1}[ky][2]
        const float t0 = __shfl_up_sync(0xffffffffu, d[0][0], 1) + d[0][1] + __shfl_down_sync(0xffffffffu, d[0][2], 1);   // -> row h+1
        const float t1 = __shfl_up_sync(0xffffffffu, d[1][0], 1) + d[1][1] + __shfl_down_sync(0xffffffffu, d[1][2], 1);   // -> row h
        const float t2 = __shfl_up_sync(0xffffffffu, d[2][0], 1) + d[2][1] + __shfl_down_sync(0xffffffffu, d[2][2], 1);   // -> row h-1
        acc_prev += t2;
        if (writes && h - 1 >= r0 && h - 1 < r1)
            __stcs(o + (size_t)(h - 1) * W + col, P.sigmoid ? (1.0f / (1.0f + expf(-acc_prev))) : acc_prev);
        acc_prev = acc_cur + t1;
        acc_cur = t0;
    }
}

}  // namespace riders

extern "C" int riders_conv3x3_head_nhwc(const float* x, int N, int C, int H, int W,
                                        const float* bias_host, float slope, const float* weight_host,
                                        float* out, void* stream) {
    return riders_conv3x3_head_sigmoid_nhwc(x, N, C, H, W, bias_host, slope, weight_host, 0, out, stream);
}

extern "C" int riders_conv3x3_head_sigmoid_nhwc(const float* x, int N, int C, int H, int W,
                                                const float* bias_host, float slope, const float* weight_host,
                                                int apply_sigmoid, float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && H > 0 && W > 0, "conv3x3_head: bad shape");
    RIDERS_REQUIRE(C == HEAD_C, "conv3x3_head: built for %d input channels (got %d)", HEAD_C, C);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && weight_host && out, "conv3x3_head: null pointer");
    RIDERS_REQUIRE((((uintptr_t)x) & 15) == 0, "conv3x3_head: input must be 16-byte aligned");
    HeadParams P;
    for (int c = 0; c < HEAD_C; ++c) {
        P.bias[c] = bias_host ? bias_host[c] : 0.f;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) P.w[ky][kx][c] = weight_host[(c * 3 + ky) * 3 + kx];    // torch (1, C, 3, 3)
    }
    P.slope = slope;
    P.sigmoid = apply_sigmoid ? 1 : 0;
    const int strips = ceil_div(W, 30);
    const int S = ceil_div(W, strips);
    const int segs = ceil_div(H, 64);
    const int rows_per_seg = ceil_div(H, segs);
    const long long units = (long long)N * segs * strips;
    const long long blocks = (units + HEAD_WARPS - 1) / HEAD_WARPS;
    RIDERS_REQUIRE(blocks < (1ll << 31), "conv3x3_head: too many patches for one launch");
    count_launches(1);
    conv3x3_head_kernel<<<(unsigned)blocks, HEAD_WARPS * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), out, P, H, W, S, strips, rows_per_seg, segs, units);
    return check_launch("conv3x3_head_nhwc");
}
