// The host loops on either side of the path, as kernels:
//
// N4  RC-Net input prologue, run_rcnet_zju.py:214-240 + rcnet_transforms.py:100-103, 258-261 + rcnet_main.py:441-444
//       if max(image) > 1:  image = image.int()                 (per frame: the reference batch is one frame)
//       image = image / 255.0                                   (normalized_image_range [0, 1])
//       image = pad(image, (px, py, px, py), 'edge')            px = pw // 2, py = ph // 2
//       point[:, 0] += px;  point[:, 1] += py                   patch centre in padded coordinates
//       box = (x - px, y - py, x + px, y + py)                  of the SHIFTED point; RoI row = (frame, box)
//     One pass writes the padded, normalised frame (NCHW or channels-last, what the backbone wants) and the
//     shifted points + RoI rows: replaces int() + div + replication_pad + layout conversion + the per-point Python loop.
//
// N2  SML input assembly, val_zju.py:181-196 + modules/midas/transforms.py:127-139 (Resize, cv2.INTER_NEAREST on
//     every item of the sample), :185-193 (NormalizeIntermediate), :196-217 (PrepareForNet):
//       sx = min(floor(x * (1 / (ow / W))), W - 1)   in double, as cv2's resizeNN builds its x_ofs table (same for y)
//       x[0] = (int_depth[sy, sx] - mean_d) / std_d ;  x[1] = (int_scales[sy, sx] - mean_s) / std_s      (float32)
//       x[2] = r * 0.299 + g * 0.587 + b * 0.114     (float32, left to right, no FMA contraction)
//       d    = int_depth[sy, sx]                       ('int_depth_no_tf')
//     float32 statement for statement -> bit-exact.
#include "common.cuh"

namespace riders {

constexpr int PR_THREADS = 256;

__global__ void __launch_bounds__(PR_THREADS)
prologue_flag_kernel(const float* __restrict__ image, long long per_frame, int32_t* __restrict__ flags) {
    const int b = blockIdx.y;
    const float* src = image + (size_t)b * per_frame;
    bool big = false;
    for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < per_frame; i += (long long)gridDim.x * PR_THREADS)
        big |= src[i] > 1.0f;
    if (__syncthreads_or(big) && threadIdx.x == 0) flags[b] = 1;      // benign race: every writer stores 1
}

template <bool NHWC>
__global__ void __launch_bounds__(PR_THREADS)
prologue_pad_kernel(const float* __restrict__ image, int B, int C, int H, int W, int px, int py,
                    const int32_t* __restrict__ flags,
                    const float* __restrict__ points, const int32_t* __restrict__ frame_offsets, int R,
                    float* __restrict__ out, float* __restrict__ points_out, float* __restrict__ rois) {
    const int Hp = H + 2 * py, Wp = W + 2 * px;
    const long long total = (long long)B * C * Hp * Wp;
    const long long tid = (long long)blockIdx.x * PR_THREADS + threadIdx.x;
    if (tid < R) {
        // run_rcnet_zju.py:222-232: shift first, then the box around the shifted point (same float32 order)
        const float x = __fadd_rn(points[tid * 3 + 0], (float)px), y = __fadd_rn(points[tid * 3 + 1], (float)py);
        points_out[tid * 3 + 0] = x; points_out[tid * 3 + 1] = y; points_out[tid * 3 + 2] = points[tid * 3 + 2];
        int lo = 0, hi = B;                                           // frame f: offsets[f] <= tid < offsets[f + 1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (frame_offsets[mid] <= (int)tid) lo = mid; else hi = mid; }
        float* r = rois + tid * 5;
        r[0] = (float)lo;
        r[1] = __fsub_rn(x, (float)px); r[2] = __fsub_rn(y, (float)py);
        r[3] = __fadd_rn(x, (float)px); r[4] = __fadd_rn(y, (float)py);
    }
    for (long long e = tid; e < total; e += (long long)gridDim.x * PR_THREADS) {
        int c, xx, yy, b;
        long long t = e;
        if (NHWC) { c = (int)(t % C); t /= C; xx = (int)(t % Wp); t /= Wp; yy = (int)(t % Hp); b = (int)(t / Hp); }
        else      { xx = (int)(t % Wp); t /= Wp; yy = (int)(t % Hp); t /= Hp; c = (int)(t % C); b = (int)(t / C); }
        const int sx = min(max(xx - px, 0), W - 1), sy = min(max(yy - py, 0), H - 1);
        float v = __ldg(image + (((size_t)b * C + c) * H + sy) * W + sx);
        if (flags[b]) v = (float)(int)v;                              // Tensor.int(): truncation toward zero
        out[e] = __fdiv_rn(v, 255.0f);
    }
}

// The same pass for the usual case - 3 channels, channels-last out - without per-element 64-bit index arithmetic: blockIdx =
// (column block, padded row, frame), thread = padded pixel: three plane reads that are coalesced over the warp (the element-wise
// form read the three planes interleaved lane by lane) and one contiguous 12-byte store per thread.  Same arithmetic, bit for bit.
__global__ void __launch_bounds__(PR_THREADS)
prologue_pad3_nhwc_kernel(const float* __restrict__ image, int H, int W, int px, int py, const int32_t* __restrict__ flags,
                          float* __restrict__ out) {
    const int Hp = H + 2 * py, Wp = W + 2 * px;
    const int xx = blockIdx.x * PR_THREADS + threadIdx.x, yy = blockIdx.y, b = blockIdx.z;
    if (xx >= Wp) return;
    const int sx = min(max(xx - px, 0), W - 1), sy = min(max(yy - py, 0), H - 1);
    const float* src = image + ((size_t)b * 3 * H + sy) * W + sx;
    const size_t plane = (size_t)H * W;
    float v0 = __ldg(src), v1 = __ldg(src + plane), v2 = __ldg(src + 2 * plane);
    if (flags[b]) { v0 = (float)(int)v0; v1 = (float)(int)v1; v2 = (float)(int)v2; }      // Tensor.int(): truncation toward zero
    float* o = out + (((size_t)b * Hp + yy) * Wp + xx) * 3;
    o[0] = __fdiv_rn(v0, 255.0f); o[1] = __fdiv_rn(v1, 255.0f); o[2] = __fdiv_rn(v2, 255.0f);
}

// points / RoI rows alone (the pixel part of prologue_pad_kernel is taken by prologue_pad3_nhwc_kernel)
__global__ void __launch_bounds__(PR_THREADS)
prologue_points_kernel(int B, int px, int py, const float* __restrict__ points, const int32_t* __restrict__ frame_offsets, int R,
                       float* __restrict__ points_out, float* __restrict__ rois) {
    const long long tid = (long long)blockIdx.x * PR_THREADS + threadIdx.x;
    if (tid >= R) return;
    const float x = __fadd_rn(points[tid * 3 + 0], (float)px), y = __fadd_rn(points[tid * 3 + 1], (float)py);
    points_out[tid * 3 + 0] = x; points_out[tid * 3 + 1] = y; points_out[tid * 3 + 2] = points[tid * 3 + 2];
    int lo = 0, hi = B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (frame_offsets[mid] <= (int)tid) lo = mid; else hi = mid; }
    float* r = rois + tid * 5;
    r[0] = (float)lo;
    r[1] = __fsub_rn(x, (float)px); r[2] = __fsub_rn(y, (float)py);
    r[3] = __fadd_rn(x, (float)px); r[4] = __fadd_rn(y, (float)py);
}

__global__ void __launch_bounds__(PR_THREADS)
sml_input_kernel(const float* __restrict__ int_depth, const float* __restrict__ int_scales, const float* __restrict__ image,
                 int hwc, int B, int H, int W, int oh, int ow, double ify, double ifx,
                 float mean_d, float std_d, float mean_s, float std_s,
                 float* __restrict__ x, float* __restrict__ d) {
    const long long total = (long long)B * oh * ow;
    for (long long e = (long long)blockIdx.x * PR_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * PR_THREADS) {
        const int ox = (int)(e % ow);
        const long long t = e / ow;
        const int oy = (int)(t % oh), b = (int)(t / oh);
        const int sx = min(__double2int_rd(ox * ifx), W - 1), sy = min(__double2int_rd(oy * ify), H - 1);
        const size_t s = ((size_t)b * H + sy) * W + sx;
        const float dep = __ldg(int_depth + s), sc = __ldg(int_scales + s);
        float r, g, bl;
        if (hwc) { r = __ldg(image + s * 3); g = __ldg(image + s * 3 + 1); bl = __ldg(image + s * 3 + 2); }
        else {
            const size_t plane = (size_t)H * W, s0 = (size_t)b * 3 * plane + (size_t)sy * W + sx;
            r = __ldg(image + s0); g = __ldg(image + s0 + plane); bl = __ldg(image + s0 + 2 * plane);
        }
        const size_t o = (size_t)b * 3 * oh * ow + (size_t)oy * ow + ox, op = (size_t)oh * ow;
        x[o] = __fdiv_rn(__fsub_rn(dep, mean_d), std_d);
        x[o + op] = __fdiv_rn(__fsub_rn(sc, mean_s), std_s);
        x[o + 2 * op] = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(bl, 0.114f));
        d[(size_t)b * op + (size_t)oy * ow + ox] = dep;
    }
}

}  // namespace riders

extern "C" int riders_rcnet_prologue(const float* image, int B, int C, int H, int W, int pad_x, int pad_y, int out_layout,
                                     const float* points, const int32_t* frame_offsets, int R,
                                     float* padded, float* points_shifted, float* rois, int32_t* frame_flags, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && pad_x >= 0 && pad_y >= 0 && R >= 0, "rcnet_prologue: bad shape");
    RIDERS_REQUIRE(out_layout == RIDERS_LAYOUT_NCHW || out_layout == RIDERS_LAYOUT_NHWC, "rcnet_prologue: bad layout %d", out_layout);
    RIDERS_REQUIRE(image && padded && frame_flags, "rcnet_prologue: null pointer");
    RIDERS_REQUIRE(R == 0 || (points && frame_offsets && points_shifted && rois), "rcnet_prologue: null point arrays");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(frame_flags, 0, sizeof(int32_t) * B, st) != cudaSuccess) return check_launch("rcnet_prologue: memset");
    const long long per_frame = (long long)C * H * W;
    const int fb = (int)((per_frame + 4 * PR_THREADS - 1) / (4 * PR_THREADS));
    count_launches(2);
    prologue_flag_kernel<<<dim3(fb < 64 ? fb : 64, B), PR_THREADS, 0, st>>>(image, per_frame, frame_flags);
    const long long total = (long long)B * C * (H + 2 * pad_y) * (W + 2 * pad_x);
    long long blocks = (total + PR_THREADS - 1) / PR_THREADS;
    const long long cap = (long long)sm_count() * 16, need = ((long long)R + PR_THREADS - 1) / PR_THREADS;
    if (blocks > cap) blocks = cap;
    if (blocks < need) blocks = need;
    if (out_layout == RIDERS_LAYOUT_NHWC && C == 3 && H + 2 * pad_y <= 65535 && B <= 65535) {
        if (R > 0) {
            count_launches(1);
            prologue_points_kernel<<<(unsigned)need, PR_THREADS, 0, st>>>(B, pad_x, pad_y, points, frame_offsets, R, points_shifted, rois);
        }
        prologue_pad3_nhwc_kernel<<<dim3(ceil_div(W + 2 * pad_x, PR_THREADS), H + 2 * pad_y, B), PR_THREADS, 0, st>>>(
            image, H, W, pad_x, pad_y, frame_flags, padded);
    } else if (out_layout == RIDERS_LAYOUT_NHWC)
        prologue_pad_kernel<true><<<(unsigned)blocks, PR_THREADS, 0, st>>>(image, B, C, H, W, pad_x, pad_y, frame_flags, points,
                                                                        frame_offsets, R, padded, points_shifted, rois);
    else
        prologue_pad_kernel<false><<<(unsigned)blocks, PR_THREADS, 0, st>>>(image, B, C, H, W, pad_x, pad_y, frame_flags, points,
                                                                         frame_offsets, R, padded, points_shifted, rois);
    return check_launch("rcnet_prologue");
}

extern "C" int riders_sml_input_assemble(const float* int_depth, const float* int_scales, const float* image, int image_hwc,
                                         int B, int H, int W, int out_h, int out_w,
                                         float depth_mean, float depth_std, float scales_mean, float scales_std,
                                         float* x, float* d, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "sml_input_assemble: bad shape");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(int_depth && int_scales && image && x && d, "sml_input_assemble: null pointer");
    RIDERS_REQUIRE(depth_std != 0.0f && scales_std != 0.0f, "sml_input_assemble: zero std");
    // cv2.resize: inv_scale = dsize / ssize (double); resizeNN: ifx = 1 / inv_scale; x_ofs[x] = min(floor(x * ifx), ssize - 1)
    const double ifx = 1.0 / ((double)out_w / (double)W), ify = 1.0 / ((double)out_h / (double)H);
    const long long total = (long long)B * out_h * out_w;
    long long blocks = (total + PR_THREADS - 1) / PR_THREADS;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    count_launches(1);
    sml_input_kernel<<<(unsigned)blocks, PR_THREADS, 0, (cudaStream_t)stream>>>(int_depth, int_scales, image, image_hwc, B, H, W,
                                                                               out_h, out_w, ify, ifx, depth_mean, depth_std,
                                                                               scales_mean, scales_std, x, d);
    return check_launch("sml_input_assemble");
}
