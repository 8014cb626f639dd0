// A1: per-radar-point feature gather == torchvision.ops.roi_pool on every pyramid level
// (replaces the five calls at RCNet/networks.py:418-433 of the reference).
//
// The op is a pure selection (max over a quantised window), so results are bit-exact.
// Roofline: HBM-write bound - every RoI writes C*oh*ow floats per level (1.39 MB per radar point
// at patch 240x100) while the feature maps (32.6 MB per frame) are read through L2.
//
// Two forward paths:
//   * roi_gather_tma_kernel  (channels-last maps): one CTA = one RoI x a strip of pooled rows.
//     The input window of the strip is staged into shared memory by ONE 4-D TMA tile load
//     (cp.async.bulk.tensor, box = C x BW x BH x 1, out-of-map parts zero-filled and never
//     read because bins are clipped).  Threads then take the bin maxima as float4 over
//     channels and stream them out fully coalesced.  Several CTAs per SM overlap load and math.
//   * roi_gather_generic_kernel (any layout, any RoI size, optional argmax): one thread per
//     pooled element (NCHW) or per 4 channels (NHWC), reading the window through L1/L2.
// Backward (training, RCNet/rcnet_main.py:342-358): deterministic gather formulation - one
// thread per input element visits, in RoI order, the few bins that can contain it and adds
// the gradients whose argmax points at it.  No float atomics.
#include "common.cuh"
#include "roi_bins.cuh"
#include <cuda.h>
#include <float.h>

namespace riders {

struct GatherLevel {
    const float* feat; float* out; int32_t* argmax;
    int C, H, W, oh, ow;
    int ld;                    // floats per output pixel (channels-last only; C unless the output is a channel slice)
    float scale;
    long long first_item;      // prefix of work items (generic kernel)
};
struct GatherParams {
    GatherLevel lv[RIDERS_MAX_LEVELS];
    int n_levels;
    int B, R;
    const float* rois;
    long long total_items;
};

constexpr int GEN_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------------
template <bool NHWC>
__global__ void __launch_bounds__(GEN_THREADS)
roi_gather_generic_kernel(const GatherParams P) {
    const long long item = (long long)blockIdx.x * GEN_THREADS + threadIdx.x;
    if (item >= P.total_items) return;
    int l = 0;
#pragma unroll
    for (int q = 1; q < RIDERS_MAX_LEVELS; ++q)
        if (q < P.n_levels && item >= P.lv[q].first_item) l = q;
    const GatherLevel& L = P.lv[l];
    long long e = item - L.first_item;
    const int CV = NHWC ? (L.C >> 2) : L.C;          // channel groups
    int r, c, i, j;
    if (NHWC) {
        c = (int)(e % CV); e /= CV;
        j = (int)(e % L.ow); e /= L.ow;
        i = (int)(e % L.oh); r = (int)(e / L.oh);
    } else {
        j = (int)(e % L.ow); e /= L.ow;
        i = (int)(e % L.oh); e /= L.oh;
        c = (int)(e % CV); r = (int)(e / CV);
    }
    const float* roi = P.rois + (size_t)r * 5;
    const int b = (int)roi[0];
    const RoiAxis ax = roi_axis(roi[1], roi[3], L.scale, L.ow);
    const RoiAxis ay = roi_axis(roi[2], roi[4], L.scale, L.oh);
    int w0, w1, h0, h1;
    roi_bin_range(ax, j, L.W, w0, w1);
    roi_bin_range(ay, i, L.H, h0, h1);
    const bool empty = (h1 <= h0) || (w1 <= w0) || b < 0 || b >= P.B;

    if (NHWC) {
        float4 best = empty ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
        int4 arg = make_int4(-1, -1, -1, -1);
        if (!empty) {
            const float* base = L.feat + (size_t)b * L.H * L.W * L.C + (size_t)c * 4;
            for (int h = h0; h < h1; ++h)
                for (int w = w0; w < w1; ++w) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(base + ((size_t)h * L.W + w) * L.C));
                    const int idx = h * L.W + w;
                    if (v.x > best.x) { best.x = v.x; arg.x = idx; }
                    if (v.y > best.y) { best.y = v.y; arg.y = idx; }
                    if (v.z > best.z) { best.z = v.z; arg.z = idx; }
                    if (v.w > best.w) { best.w = v.w; arg.w = idx; }
                }
        }
        const size_t px = ((size_t)r * L.oh + i) * L.ow + j;
        __stcs(reinterpret_cast<float4*>(L.out + px * L.ld + (size_t)c * 4), best);
        if (L.argmax) __stcs(reinterpret_cast<int4*>(L.argmax + px * L.C + (size_t)c * 4), arg);
    } else {
        float best = empty ? 0.f : -FLT_MAX;
        int arg = -1;
        if (!empty) {
            const float* base = L.feat + ((size_t)b * L.C + c) * L.H * L.W;
            for (int h = h0; h < h1; ++h)
                for (int w = w0; w < w1; ++w) {
                    const float v = __ldg(base + (size_t)h * L.W + w);
                    if (v > best) { best = v; arg = h * L.W + w; }
                }
        }
        const size_t o = (((size_t)r * L.C + c) * L.oh + i) * L.ow + j;
        __stcs(L.out + o, best);
        if (L.argmax) __stcs(L.argmax + o, arg);
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged channels-last kernel: ONE launch for every pyramid level
// ---------------------------------------------------------------------------------------------
constexpr int TMA_THREADS = 256;
constexpr int TMA_MAX_LEVELS = RIDERS_MAX_LEVELS;

struct TmaLevel {
    const float* feat; float* out;
    int C, H, W, oh, ow;
    int ld;          // floats per output pixel (C, or more when out[l] is a channel slice of a wider map)
    float scale;
    int TI;          // pooled rows per CTA
    int BH, BW;      // TMA box rows / cols
    int cv_shift;    // log2(C / 4); C / 4 is a power of two on this path
    int strips;      // CTAs per RoI
    int first_block; // first linear block index of this level
};
struct TmaPack {
    CUtensorMap map[TMA_MAX_LEVELS];
    TmaLevel lv[TMA_MAX_LEVELS];
    int n_levels, B;
    const float* rois;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ float4 max4(float4 a, const float4 b) {
    a.x = fmaxf(a.x, b.x); a.y = fmaxf(a.y, b.y); a.z = fmaxf(a.z, b.z); a.w = fmaxf(a.w, b.w);
    return a;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Bin tables of one CTA, in BYTES relative to the staged window: x = offset of the first row / column
// of the bin (-1: empty bin), y = step to its second row / column (0 if the bin has one), z = count.
// Maximum over a <= 2 x 2 bin without branches: four loads at {0, dw} x {0, dh} (duplicates are harmless).
__device__ __forceinline__ float4 bin_max_2x2(uint32_t base, int3 hb, int3 wb) {
    const uint32_t p = base + (uint32_t)hb.x + (uint32_t)wb.x;
    const float4 v00 = lds128(p), v01 = lds128(p + wb.y), v10 = lds128(p + hb.y), v11 = lds128(p + hb.y + wb.y);
    float4 r;
    r.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x)); r.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
    r.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z)); r.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
    return r;
}
// any bin size (RoIs much larger than the pooled patch)
__device__ __forceinline__ float4 bin_max_any(uint32_t base, int3 hb, int3 wb, int row_bytes, int px_bytes) {
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int h = 0; h < hb.z; ++h)
        for (int w = 0; w < wb.z; ++w) best = max4(best, lds128(base + hb.x + h * row_bytes + wb.x + w * px_bytes));
    return best;
}

__global__ void __launch_bounds__(TMA_THREADS)
roi_gather_tma_kernel(const __grid_constant__ TmaPack P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int3 wbin[256];          // per pooled column of the RoI
    __shared__ int3 hbin[64];           // per pooled row of this strip
    __shared__ int s_big;               // some bin of this CTA is larger than 2 x 2
    __shared__ int s_ragged;            // some bin is not the regular stride-1 window {i, i+1} x {j, j+1}

    int l = 0;
#pragma unroll
    for (int q = 1; q < TMA_MAX_LEVELS; ++q)
        if (q < P.n_levels && (int)blockIdx.x >= P.lv[q].first_block) l = q;
    const TmaLevel& L = P.lv[l];

    float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int blk = (int)blockIdx.x - L.first_block;
    const int r = blk / L.strips;
    const int i0 = (blk - r * L.strips) * L.TI;
    const int ni = min(L.TI, L.oh - i0);
    const float* roi = P.rois + (size_t)r * 5;
    const int b = (int)roi[0];
    const RoiAxis ax = roi_axis(roi[1], roi[3], L.scale, L.ow);
    const RoiAxis ay = roi_axis(roi[2], roi[4], L.scale, L.oh);

    // unclipped first input row of the strip and first input col of the RoI = window origin
    const int win_h = (int)floorf(__fmul_rn((float)i0, ay.bin)) + ay.start;
    const int win_w = ax.start;
    const int last_h = (int)ceilf(__fmul_rn((float)(i0 + ni), ay.bin)) + ay.start;    // exclusive
    const int last_w = (int)ceilf(__fmul_rn((float)L.ow, ax.bin)) + ax.start;
    const bool bad = b < 0 || b >= P.B;
    const bool fits = (last_h - win_h <= L.BH) && (last_w - win_w <= L.BW) && !bad;   // CTA-uniform
    const int px_bytes = L.C * 4, row_bytes = L.BW * px_bytes;

    if (threadIdx.x == 0) {
        s_big = 0; s_ragged = 0;
        if (fits) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&bar, (uint32_t)L.BH * (uint32_t)row_bytes);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%2, %3, %4, %5}], [%6];"
                ::"r"(smem_u32(tile)), "l"(&P.map[l]), "r"(0), "r"(win_w), "r"(win_h), "r"(b), "r"(smem_u32(&bar))
                : "memory");
        }
    }
    __syncthreads();
    // bin tables (clipped to the map, then made window-relative) while the tile is in flight
    bool big = false, ragged = false;
    for (int j = threadIdx.x; j < L.ow; j += TMA_THREADS) {
        int w0, w1;
        roi_bin_range(ax, j, L.W, w0, w1);
        const int n = w1 - w0;
        wbin[j] = make_int3(n > 0 ? (w0 - win_w) * px_bytes : -1, n > 1 ? px_bytes : 0, n);
        big |= n > 2;
        ragged |= n != 2 || w0 - win_w != j;
    }
    for (int i = threadIdx.x; i < ni; i += TMA_THREADS) {
        int h0, h1;
        roi_bin_range(ay, i0 + i, L.H, h0, h1);
        const int n = h1 - h0;
        hbin[i] = make_int3(n > 0 ? (h0 - win_h) * row_bytes : -1, n > 1 ? row_bytes : 0, n);
        big |= n > 2;
        ragged |= n != 2 || h0 - win_h != i;
    }
    if (big) s_big = 1;
    if (ragged) s_ragged = 1;
    __syncthreads();
    const bool any_big = s_big != 0, regular = s_ragged == 0;

    // thread -> (channel group c4, pixel); pixels advance by PP per pass, no divisions in the loop
    const int CV = 1 << L.cv_shift;
    const int c4 = threadIdx.x & (CV - 1);
    const int PP = TMA_THREADS >> L.cv_shift;
    const int npix = ni * L.ow;
    int p = threadIdx.x >> L.cv_shift;
    int i = p / L.ow, j = p - i * L.ow;
    const int di = PP / L.ow, dj = PP - di * L.ow;
    float4* out_base = reinterpret_cast<float4*>(L.out + (((size_t)r * L.oh + i0) * L.ow) * L.ld) + c4;
    const int LD4 = L.ld >> 2;                                       // float4s per output pixel
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

    if (fits) {
        mbar_wait(&bar, 0);
        const uint32_t base = smem_u32(tile) + (uint32_t)c4 * 16u;
        if (regular) {
            // RC-Net away from the frame border: the RoI extent is the pooled size + 1, so bin (i, j) is exactly the window
            // rows {i, i+1} x cols {j, j+1} - a stride-1 2 x 2 max pool.  The kernel is bound by shared-memory wavefronts
            // (ncu: LSU data pipe 89 % busy), so a thread produces two horizontally adjacent outputs from 6 loads and no
            // bin-table lookups: 24 instead of 44 wavefronts per warp and output pair.
            const int half = (L.ow + 1) >> 1;
            const int items = ni * half;
            const int dI = PP / half, dJ = PP - dI * half;
            int item = threadIdx.x >> L.cv_shift;
            int ii = item / half, jp = item - ii * half;
            const uint32_t px = (uint32_t)px_bytes, rb = (uint32_t)row_bytes;
            for (; item < items; item += PP) {
                const int jj = 2 * jp;
                const bool pair = jj + 1 < L.ow;
                const uint32_t a = base + (uint32_t)ii * rb + (uint32_t)jj * px;
                const float4 m0 = max4(lds128(a), lds128(a + rb));
                const float4 m1 = max4(lds128(a + px), lds128(a + rb + px));
                const float4 m2 = pair ? max4(lds128(a + 2 * px), lds128(a + rb + 2 * px)) : m1;
                float4* o = out_base + ((size_t)ii * L.ow + jj) * LD4;
                __stcs(o, max4(m0, m1));
                if (pair) __stcs(o + LD4, max4(m1, m2));
                ii += dI; jp += dJ;
                if (jp >= half) { jp -= half; ++ii; }
            }
        } else if (!any_big) {
            // hot path (RC-Net: RoI extent = pooled size + 1, every bin is 1 or 2 wide): two pixels per pass
            for (; p < npix; p += 2 * PP) {
                int i1 = i + di, j1 = j + dj;
                if (j1 >= L.ow) { j1 -= L.ow; ++i1; }
                const bool two = p + PP < npix;
                const int3 hb0 = hbin[i], wb0 = wbin[j];
                const int3 hb1 = two ? hbin[i1] : hb0, wb1 = two ? wbin[j1] : wb0;
                const bool e0 = (hb0.x | wb0.x) < 0, e1 = (hb1.x | wb1.x) < 0;
                float4 v0 = bin_max_2x2(base, e0 ? make_int3(0, 0, 0) : hb0, e0 ? make_int3(0, 0, 0) : wb0);
                float4 v1 = bin_max_2x2(base, e1 ? make_int3(0, 0, 0) : hb1, e1 ? make_int3(0, 0, 0) : wb1);
                __stcs(out_base + (size_t)p * LD4, e0 ? zero : v0);
                if (two) __stcs(out_base + (size_t)(p + PP) * LD4, e1 ? zero : v1);
                i = i1 + di; j = j1 + dj;
                if (j >= L.ow) { j -= L.ow; ++i; }
            }
        } else {
            for (; p < npix; p += PP) {
                const int3 hb = hbin[i], wb = wbin[j];
                const bool e = (hb.x | wb.x) < 0;
                __stcs(out_base + (size_t)p * LD4, e ? zero : bin_max_any(base, hb, wb, row_bytes, px_bytes));
                i += di; j += dj;
                if (j >= L.ow) { j -= L.ow; ++i; }
            }
        }
    } else {
        // RoI larger than the staged box (or bad batch index -> zeros): read the map directly
        const int g_row = L.W * L.C;                                   // floats
        const float* src = L.feat + (size_t)(bad ? 0 : b) * L.H * g_row + c4 * 4;
        for (; p < npix; p += PP) {
            const int3 hb = hbin[i], wb = wbin[j];
            float4 best = zero;
            if (!bad && (hb.x | wb.x) >= 0) {
                best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
                const int h0 = hb.x / row_bytes + win_h, w0 = wb.x / px_bytes + win_w;
                for (int h = h0; h < h0 + hb.z; ++h)
                    for (int w = w0; w < w0 + wb.z; ++w)
                        best = max4(best, __ldg(reinterpret_cast<const float4*>(src + (size_t)h * g_row + (size_t)w * L.C)));
            }
            __stcs(out_base + (size_t)p * LD4, best);
            i += di; j += dj;
            if (j >= L.ow) { j -= L.ow; ++i; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// deterministic backward
// ---------------------------------------------------------------------------------------------
struct GatherBwdLevel {
    const float* grad_out; const int32_t* argmax; float* grad_feat;
    int C, H, W, oh, ow;
    float scale;
};

// One thread per feature element sums, in RoI order (deterministic, no atomics), the gradients of the bins whose argmax is that
// element.  The CTA first compacts - in order - the RoIs of its frame whose row range meets the rows its 256 consecutive elements
// cover (axis arithmetic done once per RoI and CTA, kept in shared memory), so a thread visits a handful of candidates instead of
// re-deriving the bin geometry of every RoI of the batch (the first version: 37 us per RoI, 15 % of the training step).
struct BwdCand { int b, ys, yext, xs, xext, r; float ybin, xbin; };
constexpr int BWD_CHUNK = GEN_THREADS;

template <bool NHWC>
__global__ void __launch_bounds__(GEN_THREADS)
roi_gather_bwd_kernel(const GatherBwdLevel L, const float* __restrict__ rois, int R, int B) {
    __shared__ BwdCand s_c[BWD_CHUNK];
    __shared__ int s_warp[GEN_THREADS / 32];
    const long long n = (long long)B * L.C * L.H * L.W;
    const long long first = (long long)blockIdx.x * GEN_THREADS;
    const long long e0 = first + threadIdx.x;
    const bool active = e0 < n;
    long long e = active ? e0 : n - 1;
    int b, c, h, w;
    if (NHWC) { c = (int)(e % L.C); e /= L.C; w = (int)(e % L.W); e /= L.W; h = (int)(e % L.H); b = (int)(e / L.H); }
    else      { w = (int)(e % L.W); e /= L.W; h = (int)(e % L.H); e /= L.H; c = (int)(e % L.C); b = (int)(e / L.C); }
    const int me = h * L.W + w;
    // rows covered by this CTA (when all its elements lie in one H x W plane: one frame for NHWC, one (frame, channel) for NCHW)
    const long long last = min(first + GEN_THREADS - 1, n - 1);
    const long long plane = NHWC ? (long long)L.H * L.W * L.C : (long long)L.H * L.W;
    const long long row_len = NHWC ? (long long)L.W * L.C : (long long)L.W;
    const bool one_plane = first / plane == last / plane;
    const int h_lo = (int)((first % plane) / row_len), h_hi = (int)((last % plane) / row_len);
    const int b_lo = (int)(first / ((long long)L.C * L.H * L.W)), b_hi = (int)(last / ((long long)L.C * L.H * L.W));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int r0 = 0; r0 < R; r0 += BWD_CHUNK) {
        // ---- ordered compaction of the chunk's RoIs that can touch this CTA ------------------------------------------------
        const int r = r0 + threadIdx.x;
        bool keep = false;
        BwdCand cd;
        if (r < R) {
            const float* roi = rois + (size_t)r * 5;
            cd.b = (int)roi[0]; cd.r = r;
            const RoiAxis ax = roi_axis(roi[1], roi[3], L.scale, L.ow);
            const RoiAxis ay = roi_axis(roi[2], roi[4], L.scale, L.oh);
            cd.ys = ay.start; cd.yext = ay.extent; cd.ybin = ay.bin;
            cd.xs = ax.start; cd.xext = ax.extent; cd.xbin = ax.bin;
            keep = cd.b >= b_lo && cd.b <= b_hi && (!one_plane || (cd.ys <= h_hi && cd.ys + cd.yext >= h_lo));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int i = 0; i < GEN_THREADS / 32; ++i) { const int v = s_warp[i]; if (i < warp) off += v; total += v; }
        if (keep) s_c[off + __popc(m & ((1u << lane) - 1))] = cd;
        __syncthreads();
        // ---- this element against the candidates, in RoI order --------------------------------------------------------------
        if (active) {
            for (int k = 0; k < total; ++k) {
                const BwdCand q = s_c[k];
                if (q.b != b) continue;
                const int hr = h - q.ys, wr = w - q.xs;
                if (hr < 0 || wr < 0 || hr > q.yext || wr > q.xext) continue;
                RoiAxis ax, ay;
                ax.start = q.xs; ax.extent = q.xext; ax.bin = q.xbin;
                ay.start = q.ys; ay.extent = q.yext; ay.bin = q.ybin;
                const int ie = (int)((float)hr / ay.bin), je = (int)((float)wr / ax.bin);
                for (int i = max(ie - 2, 0); i <= min(ie + 2, L.oh - 1); ++i) {
                    int h0, h1;
                    roi_bin_range(ay, i, L.H, h0, h1);
                    if (h < h0 || h >= h1) continue;
                    for (int j = max(je - 2, 0); j <= min(je + 2, L.ow - 1); ++j) {
                        int w0, w1;
                        roi_bin_range(ax, j, L.W, w0, w1);
                        if (w < w0 || w >= w1) continue;
                        const size_t o = NHWC ? ((((size_t)q.r * L.oh + i) * L.ow + j) * L.C + c)
                                              : ((((size_t)q.r * L.C + c) * L.oh + i) * L.ow + j);
                        if (L.argmax[o] == me) acc += L.grad_out[o];
                    }
                }
            }
        }
        __syncthreads();                                 // the list is rewritten by the next chunk
    }
    if (active) L.grad_feat[e0] += acc;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

static int g_tma_smem_budget = 48 * 1024;     // per CTA -> 4 CTAs per SM; measured best of {24..100} KB (tunable: riders_debug_set("gather_smem_kb", v))

void set_gather_smem_budget(int kb) { if (kb >= 16 && kb <= 200) g_tma_smem_budget = kb * 1024; }

// choose strip height / box for one level; returns false if the TMA path cannot be used
static bool plan_tma_level(TmaLevel& L, int max_roi_w, int max_roi_h) {
    if (max_roi_w <= 0 || max_roi_h <= 0) return false;
    const int cv = L.C >> 2;
    if ((L.C & 3) || cv > TMA_THREADS || (cv & (cv - 1)) || L.C > 256 || L.ow > 256) return false;
    L.cv_shift = 0;
    while ((1 << L.cv_shift) < cv) ++L.cv_shift;
    const int rw_max = (int)ceilf(max_roi_w * L.scale) + 2;
    const int rh_max = (int)ceilf(max_roi_h * L.scale) + 2;
    const float bh_max = (float)rh_max / (float)L.oh;
    L.BW = rw_max;
    if (L.BW > 256) return false;
    int best = 0;
    for (int ti = 1; ti <= L.oh && ti <= 64; ++ti) {
        const int bh = (int)(ti * bh_max) + 2;
        if (bh > 256) break;
        if ((size_t)bh * L.BW * L.C * 4 + 128 <= (size_t)g_tma_smem_budget) best = ti;
    }
    if (best == 0) return false;
    // even out the strips
    L.strips = ceil_div(L.oh, best);
    L.TI = ceil_div(L.oh, L.strips);
    L.strips = ceil_div(L.oh, L.TI);
    L.BH = (int)(L.TI * bh_max) + 2;
    return true;
}

static int encode_level_map(CUtensorMap* tmap, const TmaLevel& L, int B) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("roi_gather: cuTensorMapEncodeTiled unavailable"); return RIDERS_ECUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)L.C, (cuuint64_t)L.W, (cuuint64_t)L.H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)L.C * 4, (cuuint64_t)L.W * L.C * 4, (cuuint64_t)L.H * L.W * L.C * 4};
    cuuint32_t box[4] = {(cuuint32_t)L.C, (cuuint32_t)L.BW, (cuuint32_t)L.BH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)L.feat, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("roi_gather: cuTensorMapEncodeTiled failed (%d)", (int)cr); return RIDERS_ECUDA; }
    return RIDERS_OK;
}

static int launch_tma_pack(TmaPack& P, int R, cudaStream_t st) {
    size_t smem = 0;
    long long blocks = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        TmaLevel& L = P.lv[l];
        int rc = encode_level_map(&P.map[l], L, P.B);
        if (rc != RIDERS_OK) return rc;
        const size_t need = (size_t)L.BH * L.BW * L.C * 4 + 128;
        if (need > smem) smem = need;
        L.first_block = (int)blocks;
        blocks += (long long)R * L.strips;
    }
    RIDERS_REQUIRE(blocks < (1ll << 31), "roi_gather: too many CTAs for one launch; chunk the RoIs");
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(roi_gather_tma_kernel, (size_t)(smem), configured, "roi_gather: smem attribute")) return rc_smem;
    count_launches(1);
    roi_gather_tma_kernel<<<(unsigned)blocks, TMA_THREADS, smem, st>>>(P);
    return RIDERS_OK;
}

}  // namespace riders

extern "C" int riders_roi_gather_fwd(int n_levels,
                                     const float* const* feat, const int* C, const int* Hl, const int* Wl,
                                     const float* spatial_scale, const int* oh, const int* ow,
                                     int layout, int B, const float* rois, int R,
                                     int max_roi_w, int max_roi_h,
                                     float* const* out, int32_t* const* argmax, void* stream) {
    return riders_roi_gather_fwd_placed(n_levels, feat, C, Hl, Wl, spatial_scale, oh, ow, layout, B, rois, R,
                                        max_roi_w, max_roi_h, out, nullptr, argmax, stream);
}

extern "C" int riders_roi_gather_fwd_placed(int n_levels,
                                            const float* const* feat, const int* C, const int* Hl, const int* Wl,
                                            const float* spatial_scale, const int* oh, const int* ow,
                                            int layout, int B, const float* rois, int R,
                                            int max_roi_w, int max_roi_h,
                                            float* const* out, const int* out_pixel_stride, int32_t* const* argmax,
                                            void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(n_levels > 0 && n_levels <= RIDERS_MAX_LEVELS, "roi_gather: n_levels=%d out of range", n_levels);
    RIDERS_REQUIRE(layout == RIDERS_LAYOUT_NCHW || layout == RIDERS_LAYOUT_NHWC, "roi_gather: bad layout %d", layout);
    RIDERS_REQUIRE(B > 0 && R >= 0, "roi_gather: bad B=%d R=%d", B, R);
    RIDERS_REQUIRE(feat && C && Hl && Wl && spatial_scale && oh && ow && out, "roi_gather: null pointer");
    if (R == 0) return RIDERS_OK;
    RIDERS_REQUIRE(rois, "roi_gather: null rois");
    RIDERS_REQUIRE(R <= 65535, "roi_gather: at most 65535 RoIs per call (got %d); chunk the call", R);
    cudaStream_t st = (cudaStream_t)stream;

    GatherParams P;
    P.n_levels = 0; P.B = B; P.R = R; P.rois = rois;
    static thread_local TmaPack TP;           // 1.5 KB: the levels that take the TMA path, one launch
    TP.n_levels = 0; TP.B = B; TP.rois = rois;
    long long items = 0;
    for (int l = 0; l < n_levels; ++l) {
        RIDERS_REQUIRE(feat[l] && out[l], "roi_gather: null map at level %d", l);
        RIDERS_REQUIRE(C[l] > 0 && Hl[l] > 0 && Wl[l] > 0 && oh[l] > 0 && ow[l] > 0, "roi_gather: bad dims at level %d", l);
        RIDERS_REQUIRE((long long)Hl[l] * Wl[l] < (1ll << 31), "roi_gather: map too large at level %d", l);
        const bool want_arg = argmax && argmax[l];
        const int ld = (out_pixel_stride && out_pixel_stride[l] > 0) ? out_pixel_stride[l] : C[l];
        RIDERS_REQUIRE(ld == C[l] || (layout == RIDERS_LAYOUT_NHWC && ld > C[l] && (ld & 3) == 0),
                       "roi_gather: out_pixel_stride[%d]=%d needs the channels-last layout, >= C and %%4==0", l, ld);
        if (layout == RIDERS_LAYOUT_NHWC) {
            RIDERS_REQUIRE((C[l] & 3) == 0 && (((uintptr_t)feat[l]) & 15) == 0 && (((uintptr_t)out[l]) & 15) == 0,
                           "roi_gather: channels-last path needs C%%4==0 and 16-byte aligned maps (level %d)", l);
            if (!want_arg) {
                TmaLevel& T = TP.lv[TP.n_levels];
                T.feat = feat[l]; T.out = out[l]; T.C = C[l]; T.H = Hl[l]; T.W = Wl[l]; T.oh = oh[l]; T.ow = ow[l]; T.ld = ld;
                T.scale = spatial_scale[l];
                if (plan_tma_level(T, max_roi_w, max_roi_h)) { ++TP.n_levels; continue; }
            }
        }
        GatherLevel& G = P.lv[P.n_levels++];
        G.feat = feat[l]; G.out = out[l]; G.argmax = want_arg ? argmax[l] : nullptr;
        G.C = C[l]; G.H = Hl[l]; G.W = Wl[l]; G.oh = oh[l]; G.ow = ow[l]; G.ld = ld; G.scale = spatial_scale[l];
        G.first_item = items;
        const long long cg = (layout == RIDERS_LAYOUT_NHWC) ? C[l] / 4 : C[l];
        items += (long long)R * cg * oh[l] * ow[l];
    }
    if (TP.n_levels > 0) {
        int rc = launch_tma_pack(TP, R, st);
        if (rc != RIDERS_OK) return rc;
    }
    if (P.n_levels > 0) {
        P.total_items = items;
        const long long blocks = (items + GEN_THREADS - 1) / GEN_THREADS;
        RIDERS_REQUIRE(blocks < (1ll << 31), "roi_gather: too many elements for one launch; chunk the RoIs");
        count_launches(1);
        if (layout == RIDERS_LAYOUT_NHWC)
            roi_gather_generic_kernel<true><<<(unsigned)blocks, GEN_THREADS, 0, st>>>(P);
        else
            roi_gather_generic_kernel<false><<<(unsigned)blocks, GEN_THREADS, 0, st>>>(P);
    }
    return check_launch("roi_gather_fwd");
}

extern "C" int riders_roi_gather_bwd(int n_levels,
                                     const float* const* grad_out, const int32_t* const* argmax,
                                     const int* C, const int* Hl, const int* Wl, const int* oh, const int* ow,
                                     const float* spatial_scale,
                                     int layout, int B, const float* rois, int R,
                                     float* const* grad_feat, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(n_levels > 0 && n_levels <= RIDERS_MAX_LEVELS, "roi_gather_bwd: n_levels=%d out of range", n_levels);
    RIDERS_REQUIRE(layout == RIDERS_LAYOUT_NCHW || layout == RIDERS_LAYOUT_NHWC, "roi_gather_bwd: bad layout %d", layout);
    RIDERS_REQUIRE(grad_out && argmax && C && Hl && Wl && oh && ow && spatial_scale && grad_feat, "roi_gather_bwd: null pointer");
    RIDERS_REQUIRE(B > 0 && R >= 0, "roi_gather_bwd: bad B=%d R=%d", B, R);
    if (R == 0) return RIDERS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 0; l < n_levels; ++l) {
        RIDERS_REQUIRE(grad_out[l] && argmax[l] && grad_feat[l], "roi_gather_bwd: null map at level %d", l);
        GatherBwdLevel L;
        L.grad_out = grad_out[l]; L.argmax = argmax[l]; L.grad_feat = grad_feat[l];
        L.C = C[l]; L.H = Hl[l]; L.W = Wl[l]; L.oh = oh[l]; L.ow = ow[l]; L.scale = spatial_scale[l];
        const long long n = (long long)B * L.C * L.H * L.W;
        const long long blocks = (n + GEN_THREADS - 1) / GEN_THREADS;
        RIDERS_REQUIRE(blocks < (1ll << 31), "roi_gather_bwd: map too large");
        count_launches(1);
        if (layout == RIDERS_LAYOUT_NHWC)
            roi_gather_bwd_kernel<true><<<(unsigned)blocks, GEN_THREADS, 0, st>>>(L, rois, R, B);
        else
            roi_gather_bwd_kernel<false><<<(unsigned)blocks, GEN_THREADS, 0, st>>>(L, rois, R, B);
    }
    return check_launch("roi_gather_bwd");
}
