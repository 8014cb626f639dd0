// A5: thresholded confidence -> quasi-dense depth (RCNet/rcnet_main.py:460-487 of the reference).
//
// The reference materialises a (K, H', W') tensor of zero canvases (1.14 GB per frame at
// K=512), then reduces it three times.  Here every output pixel GATHERS from the crops that
// cover it: crop pixel (i, j) of point k lands on unpadded pixel (y_k - ph + i, x_k - pw + j)
// with (x_k, y_k) = int() of the pad-shifted point, so each confidence is read exactly once,
// coalesced along the crop row, and nothing but the three maps is written.
//
// HBM roofline: reads R*ph*pw*4 B of confidences + 12 B/point, writes 3 maps (12 B/pixel).
//
// Work decomposition: CTA = TW x (8 * 32/TW * ROWS) pixel tile of one frame, 8 warps; a warp covers TW columns x 32/TW row
// groups, each thread ROWS = 4 vertically stacked pixels (amortises the per-point address math over 4 loads).  TW = 32 is the
// round-1 shape (32 x 32 tiles); TW = 16 (16 x 64 tiles, two row groups per warp) wastes fewer lanes on a 100-wide crop
// that straddles the tiles ((100 + 16) / 100 = 1.16 lanes per useful lane instead of (100 + 32) / 100 = 1.32) and still reads
// 64-byte row segments.  The CTA first compacts, in point order (so the float sums run in the reference's k order and the
// first-max rule holds), the points whose crop overlaps the tile.
#include "common.cuh"

namespace riders {

constexpr int CR_THREADS = 256;
constexpr int CHUNK = 256;                 // candidate points examined per round
constexpr int NC = 2;                      // candidates gathered per inner iteration (4 measured slower)

// base = address of crop pixel (0, 0) shifted so that base + row * pw + col addresses the crop pixel under frame
// pixel (row, col): one 64-bit add per candidate and thread instead of re-deriving (i, j) offsets
struct Cand { const float* base; int r0; int c0; float z; int k; };

template <bool LOGITS>
__device__ __forceinline__ float conf_value(float v, float thr) {
    // torch.sigmoid (rcnet_model.py:123): 1 / (1 + exp(-v)) with the SFU exp2 / reciprocal (relative
    // error ~2e-7, far inside the 1e-4 tolerance on confidences); strict < (rcnet_main.py:462).
    // A pixel the crop does not cover is loaded as 0 (confidences) / -inf (logits) -> contributes exactly 0.
    if (LOGITS) v = (1.0f / (1.0f + expf(-v)))   /* full-precision expf + IEEE division: a confidence within 1 ulp of the strict `< thr` cut must not flip */;
    return (v < thr) ? 0.0f : v;
}

// predicated streaming load (no branch around it, so the loads of an iteration issue back to back)
template <bool LOGITS>
__device__ __forceinline__ float ld_stream_if(const float* p, bool on) {
    float v;
    if (LOGITS)
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "setp.ne.b32 q, %2, 0;\n"
            "mov.f32 %0, 0fFF800000;\n"
            "@q ld.global.nc.L1::no_allocate.f32 %0, [%1];\n"
            "}\n" : "=f"(v) : "l"(p), "r"((int)on));
    else
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "setp.ne.b32 q, %2, 0;\n"
            "mov.f32 %0, 0f00000000;\n"
            "@q ld.global.nc.L1::no_allocate.f32 %0, [%1];\n"
            "}\n" : "=f"(v) : "l"(p), "r"((int)on));
    return v;
}

// four vertically stacked crop pixels under ONE predicate, row stride known at compile time (immediate offsets):
// 1 setp + 4 mov + 4 predicated loads instead of 4 x (setp + mov + 64-bit add + load)
template <bool LOGITS, int PW>
__device__ __forceinline__ void ld4_stream_if(const float* p, bool on, float (&v)[4]) {
    if (LOGITS)
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "setp.ne.b32 q, %5, 0;\n"
            "mov.f32 %0, 0fFF800000;\n mov.f32 %1, 0fFF800000;\n mov.f32 %2, 0fFF800000;\n mov.f32 %3, 0fFF800000;\n"
            "@q ld.global.nc.L1::no_allocate.f32 %0, [%4];\n"
            "@q ld.global.nc.L1::no_allocate.f32 %1, [%4+%6];\n"
            "@q ld.global.nc.L1::no_allocate.f32 %2, [%4+%7];\n"
            "@q ld.global.nc.L1::no_allocate.f32 %3, [%4+%8];\n"
            "}\n" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "r"((int)on), "n"(PW * 4), "n"(PW * 8), "n"(PW * 12));
    else
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "setp.ne.b32 q, %5, 0;\n"
            "mov.f32 %0, 0f00000000;\n mov.f32 %1, 0f00000000;\n mov.f32 %2, 0f00000000;\n mov.f32 %3, 0f00000000;\n"
            "@q ld.global.nc.L1::no_allocate.f32 %0, [%4];\n"
            "@q ld.global.nc.L1::no_allocate.f32 %1, [%4+%6];\n"
            "@q ld.global.nc.L1::no_allocate.f32 %2, [%4+%7];\n"
            "@q ld.global.nc.L1::no_allocate.f32 %3, [%4+%8];\n"
            "}\n" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "r"((int)on), "n"(PW * 4), "n"(PW * 8), "n"(PW * 12));
}

// PW > 0: crop width fixed at compile time (RC-Net's 100): immediate row offsets, constant address math
// SELECT = false (no `select` map asked for): the running maximum is one FMNMX instead of compare + two selects - the kernel is
// issue-bound, and that is 2 of ~14 instructions per loaded confidence
template <bool LOGITS, int ROWS, int PW, int TILE_W, bool SELECT>
__global__ void __launch_bounds__(CR_THREADS)
confidence_reduce_kernel(const float* __restrict__ conf, const float* __restrict__ pts,
                         const int* __restrict__ frame_offsets,
                         int H, int W, int ph, int pw_rt, float thr,
                         float* __restrict__ depth, float* __restrict__ response,
                         int32_t* __restrict__ select) {
    const int pw = PW > 0 ? PW : pw_rt;
    constexpr int GROUPS = 32 / TILE_W;                  // row groups per warp
    constexpr int TILE_H = (CR_THREADS / 32) * GROUPS * ROWS;
    __shared__ Cand cand[CHUNK];
    __shared__ int warp_count[CR_THREADS / 32];
    __shared__ int n_cand;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.z;
    const int first_point = frame_offsets[frame];
    const int n_points = frame_offsets[frame + 1] - first_point;
    const size_t frame_px = (size_t)frame * H * W;
    const int tile_c = blockIdx.x * TILE_W, tile_r = blockIdx.y * TILE_H;
    const int col = tile_c + (lane % TILE_W);
    const int row = tile_r + (warp * GROUPS + lane / TILE_W) * ROWS;
    const ptrdiff_t thr_off = (ptrdiff_t)row * pw + col;

    float best[ROWS], sc[ROWS], scz[ROWS];
    int arg[ROWS];
#pragma unroll
    for (int m = 0; m < ROWS; ++m) { best[m] = 0.f; sc[m] = 0.f; scz[m] = 0.f; arg[m] = 0; }

    const size_t crop = (size_t)ph * pw;

    for (int base = 0; base < n_points; base += CHUNK) {
        // ---- ordered compaction of the points overlapping this tile -----------------
        const int k = base + tid;
        bool hit = false;
        Cand c;
        if (k < n_points) {
            const float* p = pts + (size_t)(first_point + k) * 3;
            c.c0 = (int)p[0] - pw;                     // int(x) - 2*(pw/2)  (pw even in RC-Net)
            c.r0 = (int)p[1] - ph;
            c.z = p[2];
            c.k = k;
            c.base = conf + (size_t)(first_point + k) * crop - ((ptrdiff_t)c.r0 * pw + c.c0);
            hit = (c.r0 < tile_r + TILE_H) && (c.r0 + ph > tile_r) &&
                  (c.c0 < tile_c + TILE_W) && (c.c0 + pw > tile_c);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_count[warp] = __popc(ballot);
        __syncthreads();                               // also fences the previous round's reads of cand[]
        int offset = 0;
#pragma unroll
        for (int w = 0; w < CR_THREADS / 32; ++w) offset += (w < warp) ? warp_count[w] : 0;
        if (hit) cand[offset + __popc(ballot & ((1u << lane) - 1u))] = c;
        if (tid == CR_THREADS - 1) n_cand = offset + __popc(ballot);
        __syncthreads();
        const int n = n_cand;

        // ---- gather: NC candidates per iteration, in point order (float sums as in the reference) ----
        for (int i = 0; i < n; i += NC) {
            Cand q[NC];
            float v[NC][ROWS];
            bool any[NC];
#pragma unroll
            for (int u = 0; u < NC; ++u) {
                const bool on = i + u < n;
                q[u] = cand[on ? i + u : i];           // smem broadcast
                const int i0 = row - q[u].r0;          // uniform per row group (per warp for TILE_W = 32)
                const bool in_col = on && (unsigned)(col - q[u].c0) < (unsigned)pw;
                const float* src = q[u].base + thr_off;
                any[u] = __any_sync(0xffffffffu, on && i0 + ROWS > 0 && i0 < ph);
                if (__all_sync(0xffffffffu, i0 >= 0 && i0 + ROWS <= ph)) {      // the usual case: all rows inside the 240-row crop
                    if constexpr (PW > 0 && ROWS == 4) {
                        ld4_stream_if<LOGITS, PW>(src, in_col, v[u]);
                    } else {
#pragma unroll
                        for (int m = 0; m < ROWS; ++m) v[u][m] = ld_stream_if<LOGITS>(src + m * pw, in_col);
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < ROWS; ++m)
                        v[u][m] = ld_stream_if<LOGITS>(src + m * pw, in_col && (unsigned)(i0 + m) < (unsigned)ph);
                }
            }
#pragma unroll
            for (int u = 0; u < NC; ++u) {
                if (!any[u]) continue;                 // warp-uniform: none of this warp's rows meets the crop
#pragma unroll
                for (int m = 0; m < ROWS; ++m) {
                    if (SELECT) {
                        const float x = conf_value<LOGITS>(v[u][m], thr);
                        if (x > best[m]) { best[m] = x; arg[m] = q[u].k; }      // first max wins
                        sc[m] = __fadd_rn(sc[m], x);
                        scz[m] = __fadd_rn(scz[m], __fmul_rn(x, q[u].z));       // (tile*z) then sum, rcnet_main.py:478
                    } else {
                        // the same arithmetic under one predicate: a confidence below the cut contributes exactly 0 (adding +0
                        // changes nothing), so the three updates are simply skipped - one select less per confidence
                        float x = v[u][m];
                        if (LOGITS) x = 1.0f / (1.0f + expf(-x));
                        if (!(x < thr)) {
                            best[m] = fmaxf(best[m], x);                        // same value (x >= 0, a NaN never replaces it)
                            sc[m] = __fadd_rn(sc[m], x);
                            scz[m] = __fadd_rn(scz[m], __fmul_rn(x, q[u].z));
                        }
                    }
                }
            }
        }
    }

    if (col < W) {
#pragma unroll
        for (int m = 0; m < ROWS; ++m) {
            const int r = row + m;
            if (r < H) {
                const size_t o = frame_px + (size_t)r * W + col;
                response[o] = best[m];
                depth[o] = (best[m] == 0.0f) ? 0.0f : __fdiv_rn(scz[m], sc[m]);
                if (SELECT) select[o] = arg[m];
            }
        }
    }
}

static int g_reduce_rows = 4;
static int g_reduce_tile_w = 16;
void set_reduce_rows(int rows) { if (rows == 4 || rows == 8) g_reduce_rows = rows; }
void set_reduce_tile_w(int w) { if (w == 16 || w == 32) g_reduce_tile_w = w; }

template <bool LOGITS, int ROWS, int TW>
static void launch_reduce_tw(const float* conf, const float* pts, const int* frame_offsets, int B, int H, int W, int ph, int pw,
                             float thr, float* depth, float* response, int32_t* select, cudaStream_t st) {
    dim3 grid(ceil_div(W, TW), ceil_div(H, (CR_THREADS / 32) * (32 / TW) * ROWS), B);
    if (ROWS == 4 && pw == 100) {               // RC-Net's patch width (run_rcnet_zju.py:30)
        if (select) confidence_reduce_kernel<LOGITS, ROWS, 100, TW, true><<<grid, CR_THREADS, 0, st>>>(conf, pts, frame_offsets, H, W, ph, pw, thr, depth, response, select);
        else confidence_reduce_kernel<LOGITS, ROWS, 100, TW, false><<<grid, CR_THREADS, 0, st>>>(conf, pts, frame_offsets, H, W, ph, pw, thr, depth, response, select);
    } else {
        if (select) confidence_reduce_kernel<LOGITS, ROWS, 0, TW, true><<<grid, CR_THREADS, 0, st>>>(conf, pts, frame_offsets, H, W, ph, pw, thr, depth, response, select);
        else confidence_reduce_kernel<LOGITS, ROWS, 0, TW, false><<<grid, CR_THREADS, 0, st>>>(conf, pts, frame_offsets, H, W, ph, pw, thr, depth, response, select);
    }
}

template <bool LOGITS, int ROWS>
static void launch_reduce(const float* conf, const float* pts, const int* frame_offsets, int B, int H, int W, int ph, int pw,
                          float thr, float* depth, float* response, int32_t* select, cudaStream_t st) {
    if (g_reduce_tile_w == 16 && ROWS == 4) launch_reduce_tw<LOGITS, ROWS, 16>(conf, pts, frame_offsets, B, H, W, ph, pw, thr, depth, response, select, st);
    else launch_reduce_tw<LOGITS, ROWS, 32>(conf, pts, frame_offsets, B, H, W, ph, pw, thr, depth, response, select, st);
}

}  // namespace riders

extern "C" int riders_confidence_reduce(const float* conf, int conf_is_logits,
                                        const float* pts, const int* frame_offsets,
                                        int B, int H, int W, int ph, int pw, float thr,
                                        float* depth, float* response, int32_t* select,
                                        void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0 && ph > 0 && pw > 0, "confidence_reduce: bad shape B=%d H=%d W=%d ph=%d pw=%d", B, H, W, ph, pw);
    RIDERS_REQUIRE(B <= 65535, "confidence_reduce: at most 65535 frames per call");
    RIDERS_REQUIRE(depth && response && frame_offsets, "confidence_reduce: null pointer");
    RIDERS_REQUIRE((ph % 2 == 0) && (pw % 2 == 0), "confidence_reduce: patch size must be even (got %dx%d)", ph, pw);
    if (B == 0) return RIDERS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    count_launches(1);
    if (g_reduce_rows == 8) {
        if (conf_is_logits) launch_reduce<true, 8>(conf, pts, frame_offsets, B, H, W, ph, pw, thr, depth, response, select, st);
        else launch_reduce<false, 8>(conf, pts, frame_offsets, B, H, W, ph, pw, thr, depth, response, select, st);
    } else {
        if (conf_is_logits) launch_reduce<true, 4>(conf, pts, frame_offsets, B, H, W, ph, pw, thr, depth, response, select, st);
        else launch_reduce<false, 4>(conf, pts, frame_offsets, B, H, W, ph, pw, thr, depth, response, select, st);
    }
    return check_launch("confidence_reduce");
}
