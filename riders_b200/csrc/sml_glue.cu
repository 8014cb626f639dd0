// "Next" rows of SURVEY 8(f) on either side of the alignment kernels:
//
// N2/N3  scale scaffold, val_zju.py:135-179 / train_zju.py:303-316 with the RC-Net -> SML hand-off in memory:
//          q   = uint32(z * 256) / 256, q <= 0 -> 0     the PNG codec between the two scripts
//                                                        (data/data_utils.py:128-143 save_depth, :94-116 load_depth)
//          rcnet_valid = (q < max) * (q > min);  rcnet_inv = 1 / q
//          int_scales = 1;  int_scales[rcnet_valid] = rcnet_inv / int_depth
//          int_scales[sparse_valid] = sparse_inv / int_depth            (radar points win over RC-Net pixels)
//          int_scales = (int_scales - min) / (max - min)                modules/midas/utils.py:223-237
//        float32 arithmetic in the reference's order -> bit-exact.  status = 1 where max - min <= eps
//        (the reference raises ValueError("cannot normalize array, max-min range is 0")).
//
// N4     per-frame validation metrics, utils/eval_utils.py:17-117 as called from val_zju.py:212-231:
//        mask = gt > 0 and min_val < gt < max_val; MAE, RMSE (mm), iMAE, iRMSE (1/km), AbsRel, SqRel, delta<1.25.
//        Element terms in float32 like NumPy, accumulated in fp64 by a fixed tree -> deterministic; the
//        per-frame rows feed riders_b200.dist.reduce_metric_sums (the path's only collective).
#include "common.cuh"
#include <float.h>

namespace riders {

constexpr int GL_THREADS = 256;
constexpr int GL_PX_PER_CTA = 4096;

// ------------------------------------------------------------------------------------ scaffold
__global__ void __launch_bounds__(GL_THREADS)
scaffold_raw_kernel(const float* __restrict__ int_depth, const float* __restrict__ rcnet_depth,
                    const float* __restrict__ sparse_inv, const uint8_t* __restrict__ sparse_valid,
                    int npx, int ctas_per_frame, float min_depth, float max_depth, int png_quantize,
                    float* __restrict__ out, float2* __restrict__ partial) {
    __shared__ float s_mn[GL_THREADS / 32], s_mx[GL_THREADS / 32];
    const int frame = blockIdx.y, cta = blockIdx.x;
    const size_t fo = (size_t)frame * npx;
    const int begin = cta * GL_PX_PER_CTA, end = min(begin + GL_PX_PER_CTA, npx);
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int i = begin + threadIdx.x; i < end; i += GL_THREADS) {
        const float d = int_depth[fo + i];
        float s = 1.0f;
        if (rcnet_depth) {
            float q = rcnet_depth[fo + i];
            if (png_quantize) q = __fdiv_rn(truncf(__fmul_rn(q, 256.0f)), 256.0f);
            if (q <= 0.0f) q = 0.0f;
            if (q < max_depth && q > min_depth) s = __fdiv_rn(__fdiv_rn(1.0f, q), d);
        }
        if (sparse_valid[fo + i]) s = __fdiv_rn(sparse_inv[fo + i], d);
        out[fo + i] = s;
        mn = fminf(mn, s); mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < GL_THREADS / 32; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
        partial[(size_t)frame * ctas_per_frame + cta] = make_float2(mn, mx);
    }
}

__global__ void __launch_bounds__(GL_THREADS)
scaffold_normalize_kernel(float* __restrict__ out, const float2* __restrict__ partial, int npx, int ctas_per_frame,
                          float* __restrict__ minmax, int32_t* __restrict__ status) {
    __shared__ float s_mn[GL_THREADS / 32], s_mx[GL_THREADS / 32];
    __shared__ float s_lo, s_range;
    const int frame = blockIdx.y, cta = blockIdx.x;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int c = threadIdx.x; c < ctas_per_frame; c += GL_THREADS) {
        const float2 p = partial[(size_t)frame * ctas_per_frame + c];
        mn = fminf(mn, p.x); mx = fmaxf(mx, p.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < GL_THREADS / 32; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
        const float range = __fsub_rn(mx, mn);
        s_lo = mn; s_range = range;
        if (cta == 0) {
            if (minmax) { minmax[frame * 2] = mn; minmax[frame * 2 + 1] = mx; }
            if (status) status[frame] = ((double)range > 2.220446049250313e-16) ? 0 : 1;   // np.finfo("float").eps
        }
    }
    __syncthreads();
    const float lo = s_lo, range = s_range;
    if (!((double)range > 2.220446049250313e-16)) return;                  // reference raises; leave the raw scales
    const size_t fo = (size_t)frame * npx;
    const int begin = cta * GL_PX_PER_CTA, end = min(begin + GL_PX_PER_CTA, npx);
    for (int i = begin + threadIdx.x; i < end; i += GL_THREADS)
        out[fo + i] = __fdiv_rn(__fsub_rn(out[fo + i], lo), range);
}

// ------------------------------------------------------------------------------------ metrics
struct Met { double v[8]; };          // sum|d|, sum d^2, sum|id|, sum id^2, sum|d|/t, sum d^2/t, count(delta), count

__global__ void __launch_bounds__(GL_THREADS)
metrics_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int npx, int ctas_per_frame,
                       float min_val, float max_val, double* __restrict__ partial) {
    __shared__ double sm[GL_THREADS / 32][8];
    const int frame = blockIdx.y, cta = blockIdx.x;
    const size_t fo = (size_t)frame * npx;
    const int begin = cta * GL_PX_PER_CTA, end = min(begin + GL_PX_PER_CTA, npx);
    Met m = {{0, 0, 0, 0, 0, 0, 0, 0}};
    for (int i = begin + threadIdx.x; i < end; i += GL_THREADS) {
        const float g = gt[fo + i];
        if (!(g > 0.0f && g > min_val && g < max_val)) continue;       // val_zju.py:212-221
        const float p = pred[fo + i];
        const float s = __fmul_rn(1000.0f, p), t = __fmul_rn(1000.0f, g);          // millimetres
        const float d = __fsub_rn(t, s);
        const float is = __fdiv_rn(1.0f, __fmul_rn(0.001f, p)), it = __fdiv_rn(1.0f, __fmul_rn(0.001f, g));   // 1/km
        const float id = __fsub_rn(it, is);
        const float sq = __fmul_rn(d, d);
        m.v[0] += (double)fabsf(d);
        m.v[1] += (double)sq;
        m.v[2] += (double)fabsf(id);
        m.v[3] += (double)__fmul_rn(id, id);
        m.v[4] += (double)__fdiv_rn(fabsf(__fsub_rn(s, t)), t);
        m.v[5] += (double)__fdiv_rn(__fmul_rn(__fsub_rn(s, t), __fsub_rn(s, t)), t);
        m.v[6] += (fmaxf(__fdiv_rn(g, p), __fdiv_rn(p, g)) < 1.25f) ? 1.0 : 0.0;
        m.v[7] += 1.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m.v[q] += __shfl_down_sync(0xffffffffu, m.v[q], o);
        if (lane == 0) sm[warp][q] = m.v[q];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double a = 0.0;
        for (int w = 0; w < GL_THREADS / 32; ++w) a += sm[w][threadIdx.x];
        partial[((size_t)frame * ctas_per_frame + cta) * 8 + threadIdx.x] = a;
    }
}

__global__ void __launch_bounds__(256)
metrics_final_kernel(const double* __restrict__ partial, int ctas_per_frame, double* __restrict__ metrics) {
    __shared__ double sm[8];
    const int frame = blockIdx.x;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;          // warp q adds column q
    double a = 0.0;
    for (int c = lane; c < ctas_per_frame; c += 32) a += partial[((size_t)frame * ctas_per_frame + c) * 8 + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) sm[q] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double n = sm[7];
        double* o = metrics + (size_t)frame * 8;
        o[0] = sm[0] / n;            // MAE
        o[1] = sqrt(sm[1] / n);      // RMSE
        o[2] = sm[2] / n;            // iMAE
        o[3] = sqrt(sm[3] / n);      // iRMSE
        o[4] = sm[4] / n;            // AbsRel
        o[5] = sm[5] / n;            // SqRel
        o[6] = sm[6] / n;            // delta1
        o[7] = n;                    // evaluated pixels (0 -> the means are NaN like np.mean of an empty array)
    }
}

}  // namespace riders

extern "C" size_t riders_glue_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 256;
    const size_t ctas = (size_t)riders::ceil_div(H * W, riders::GL_PX_PER_CTA);
    return (size_t)B * ctas * 8 * sizeof(double) + 256;
}

extern "C" int riders_scale_scaffold(const float* int_depth, const float* rcnet_depth,
                                     const float* sparse_inv, const uint8_t* sparse_valid,
                                     int B, int H, int W, float min_depth, float max_depth, int png_quantize,
                                     float* int_scales, float* minmax, int32_t* status,
                                     void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0, "scale_scaffold: bad shape");
    RIDERS_REQUIRE((long long)H * W < (1ll << 30) && B <= 65535, "scale_scaffold: frame too large / too many frames");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(int_depth && sparse_inv && sparse_valid && int_scales && workspace, "scale_scaffold: null pointer");
    RIDERS_REQUIRE((((uintptr_t)workspace) & 7) == 0, "scale_scaffold: workspace must be 8-byte aligned");
    const int npx = H * W, ctas = ceil_div(npx, GL_PX_PER_CTA);
    dim3 grid(ctas, B);
    cudaStream_t st = (cudaStream_t)stream;
    count_launches(2);
    scaffold_raw_kernel<<<grid, GL_THREADS, 0, st>>>(int_depth, rcnet_depth, sparse_inv, sparse_valid, npx, ctas,
                                                     min_depth, max_depth, png_quantize, int_scales, (float2*)workspace);
    scaffold_normalize_kernel<<<grid, GL_THREADS, 0, st>>>(int_scales, (const float2*)workspace, npx, ctas, minmax, status);
    return check_launch("scale_scaffold");
}

extern "C" int riders_depth_metrics(const float* pred, const float* gt, int B, int H, int W,
                                    float min_val, float max_val, double* metrics, void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0, "depth_metrics: bad shape");
    RIDERS_REQUIRE((long long)H * W < (1ll << 30) && B <= 65535, "depth_metrics: frame too large / too many frames");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(pred && gt && metrics && workspace, "depth_metrics: null pointer");
    RIDERS_REQUIRE((((uintptr_t)workspace) & 7) == 0, "depth_metrics: workspace must be 8-byte aligned");
    const int npx = H * W, ctas = ceil_div(npx, GL_PX_PER_CTA);
    cudaStream_t st = (cudaStream_t)stream;
    count_launches(2);
    metrics_partial_kernel<<<dim3(ctas, B), GL_THREADS, 0, st>>>(pred, gt, npx, ctas, min_val, max_val, (double*)workspace);
    metrics_final_kernel<<<B, 256, 0, st>>>((const double*)workspace, ctas, metrics);
    return check_launch("depth_metrics");
}
