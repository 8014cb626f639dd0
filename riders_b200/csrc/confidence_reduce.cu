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
// Work decomposition: CTA = 32 x 32 pixel tile of one frame, 8 warps, each warp a 4-row x
// 32-col strip, each thread 4 vertically stacked pixels (amortises the per-point address
// math over 4 loads).  The CTA first compacts, in point order (so the float sums run in the
// reference's k order and the first-max rule holds), the points whose crop overlaps the tile.
#include "common.cuh"

namespace riders {

constexpr int TILE_W = 32;
constexpr int TILE_H = 32;
constexpr int ROWS_PER_THREAD = 4;
constexpr int CR_THREADS = 256;
constexpr int CHUNK = 256;                 // candidate points examined per round
constexpr int NC = 2;                      // candidates gathered per inner iteration (8 loads in flight per thread; 4 measured slower)

struct Cand { int r0; int c0; float z; int k; };

template <bool LOGITS>
__device__ __forceinline__ float conf_value(float v, float thr) {
    // torch.sigmoid (rcnet_model.py:123): 1 / (1 + exp(-v)) with the SFU exp2 / reciprocal (relative
    // error ~2e-7, far inside the 1e-4 tolerance on confidences); strict < (rcnet_main.py:462)
    if (LOGITS) v = __fdividef(1.0f, 1.0f + __expf(-v));
    return (v < thr) ? 0.0f : v;
}

// predicated streaming load (no branch around it, so the 8 loads of an iteration issue back to back)
__device__ __forceinline__ float ld_stream_if(const float* p, bool on) {
    float v;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %2, 0;\n"
        "mov.f32 %0, 0f00000000;\n"
        "@q ld.global.nc.L1::no_allocate.f32 %0, [%1];\n"
        "}\n" : "=f"(v) : "l"(p), "r"((int)on));
    return v;
}

// raw values of the 4 vertically stacked crop pixels of candidate q under this thread + their coverage mask
__device__ __forceinline__ unsigned load_candidate(const float* __restrict__ conf, const Cand& q, bool on, int first_point,
                                                   size_t crop, int col, int row, int ph, int pw,
                                                   float (&v)[ROWS_PER_THREAD]) {
    const int j = col - q.c0;
    const int i0 = row - q.r0;
    const bool in_col = on && (unsigned)j < (unsigned)pw;
    const float* src = conf + (size_t)(first_point + q.k) * crop + (ptrdiff_t)i0 * pw + j;
    unsigned mask = 0;
#pragma unroll
    for (int m = 0; m < ROWS_PER_THREAD; ++m) {
        const bool in = in_col && (unsigned)(i0 + m) < (unsigned)ph;
        v[m] = ld_stream_if(src + m * pw, in);
        mask |= in ? (1u << m) : 0u;
    }
    return mask;
}

template <bool LOGITS>
__global__ void __launch_bounds__(CR_THREADS)
confidence_reduce_kernel(const float* __restrict__ conf, const float* __restrict__ pts,
                         const int* __restrict__ frame_offsets,
                         int H, int W, int ph, int pw, float thr,
                         float* __restrict__ depth, float* __restrict__ response,
                         int32_t* __restrict__ select) {
    __shared__ Cand cand[CHUNK];
    __shared__ int warp_count[CR_THREADS / 32];
    __shared__ int n_cand;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.z;
    const int first_point = frame_offsets[frame];
    const int n_points = frame_offsets[frame + 1] - first_point;
    const size_t frame_px = (size_t)frame * H * W;
    const int tile_c = blockIdx.x * TILE_W, tile_r = blockIdx.y * TILE_H;
    const int col = tile_c + lane;
    const int row = tile_r + warp * ROWS_PER_THREAD;

    float best[ROWS_PER_THREAD], sc[ROWS_PER_THREAD], scz[ROWS_PER_THREAD];
    int arg[ROWS_PER_THREAD];
#pragma unroll
    for (int m = 0; m < ROWS_PER_THREAD; ++m) { best[m] = 0.f; sc[m] = 0.f; scz[m] = 0.f; arg[m] = 0; }

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

        // ---- gather: NC candidates per iteration -> NC * 4 independent loads in flight per thread ----
        // A pixel the crop does not cover contributes v = 0, which leaves max / sums bit-identical.
        for (int i = 0; i < n; i += NC) {
            Cand q[NC];
            float v[NC][ROWS_PER_THREAD];
            unsigned msk[NC];
#pragma unroll
            for (int u = 0; u < NC; ++u) {
                const bool on = i + u < n;
                q[u] = cand[on ? i + u : i];           // smem broadcast
                msk[u] = load_candidate(conf, q[u], on, first_point, crop, col, row, ph, pw, v[u]);
            }
#pragma unroll
            for (int u = 0; u < NC; ++u) {             // in candidate (= point) order: float sums as in the reference
#pragma unroll
                for (int m = 0; m < ROWS_PER_THREAD; ++m) {
                    const float c = conf_value<LOGITS>(v[u][m], thr);          // branch-free
                    const float x = (msk[u] >> m) & 1u ? c : 0.0f;
                    if (x > best[m]) { best[m] = x; arg[m] = q[u].k; }          // first max wins
                    sc[m] = __fadd_rn(sc[m], x);
                    scz[m] = __fadd_rn(scz[m], __fmul_rn(x, q[u].z));           // (tile*z) then sum, rcnet_main.py:478
                }
            }
        }
    }

    if (col < W) {
#pragma unroll
        for (int m = 0; m < ROWS_PER_THREAD; ++m) {
            const int r = row + m;
            if (r < H) {
                const size_t o = frame_px + (size_t)r * W + col;
                response[o] = best[m];
                depth[o] = (best[m] == 0.0f) ? 0.0f : __fdiv_rn(scz[m], sc[m]);
                if (select) select[o] = arg[m];
            }
        }
    }
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
    dim3 grid(ceil_div(W, TILE_W), ceil_div(H, TILE_H), B);
    count_launches(1);
    if (conf_is_logits)
        confidence_reduce_kernel<true><<<grid, CR_THREADS, 0, st>>>(conf, pts, frame_offsets, H, W, ph, pw, thr, depth, response, select);
    else
        confidence_reduce_kernel<false><<<grid, CR_THREADS, 0, st>>>(conf, pts, frame_offsets, H, W, ph, pw, thr, depth, response, select);
    return check_launch("confidence_reduce");
}
