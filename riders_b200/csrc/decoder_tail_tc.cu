// Fused tail of RC-Net's decoder on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM): the same three
// passes as decoder_tail.cu (MultiScaleDecoder, RCNet/networks.py:678-779; DecoderBlock, utils/net_utils.py:474-569)
//
//     y      = LeakyReLU(depth_to_space(y4) + bias_a)          last block's up-convolution (sub-pixel form, see upsample.cu)
//     z      = LeakyReLU(conv3x3(y; Wb) + bias_b)              last block's 16 -> 16 convolution
//     logits = conv3x3(z; Wh)        [-> sigmoid]              output0, 16 -> 1, no bias
//
// as a streaming pipeline: a CTA owns a strip of one radar point's map (all 240 rows x up to 124 columns) and walks down it
// row by row; y and z rows live in small shared-memory rings, nothing but y4 (64 B per output pixel) is read and nothing but
// the logits (4 B) is written, and no row is ever computed twice (the mma.sync kernel recomputes a 2-pixel halo per 12 x 50
// tile: +44 % work).
//
// Both convolutions are implicit GEMMs whose A operand is a SHIFTED VIEW of a staged row:
//   * an operand row slot holds 128 (+2) consecutive pixels of one image row as two planes of 16-byte chunks (8 fp16 channels
//     each), pixel p at byte 16 p of its plane - the K-major no-swizzle canonical layout (core matrix = 8 pixels x 16 B,
//     SBO = 128 B between 8-pixel groups, LBO = plane stride between the two K chunks).  An unswizzled descriptor may start at
//     any 16-byte boundary, so the operand of horizontal tap dx is the same slot addressed 16 dx bytes further: M = 128 pixels,
//     K = 16 input channels, no im2col, no copies;
//   * the vertical taps ride in N: for y row r the B operand of tap dx stacks the three kernel rows, N = 3 x 16 output channels,
//     and the three 16-column blocks of D accumulate into the TMEM accumulators of z rows r+1, r, r-1 (a ring of twelve
//     16-column accumulators laid out so that the three targets are adjacent columns).  One z row costs 3-4 MMAs
//     (128 x 48 x 16) instead of the 9 x 2 x 8 mma.sync.m16n8k8 per 16 pixels of the legacy kernel, and its epilogue reads 16
//     finished columns;
//   * the 16 -> 1 head is the same construction with N = 16 (columns 0..2 = the head's three kernel rows): three MMAs per z row,
//     logits[o] = Dh[o-1][0] + Dh[o][1] + Dh[o+1][2] read back one column each.
// Operands are fp16 (11-bit significand, the same as the TF32 the cuDNN flow and decoder_tail.cu use) with fp32 accumulation;
// every staged value passes a running max and a value outside the fp16 range raises `status` (the host then re-runs the
// mma.sync TF32 kernel, ops.decoder_tail).
//
// Warp roles (13 warps, 2 CTAs per SM, 256 TMEM columns each): 0-3 stage y rows (global -> bias + LeakyReLU -> fp16 -> ring),
// 4-7 z epilogue (TMEM -> bias + LeakyReLU + image mask -> fp16 -> z ring), 8-11 head epilogue (TMEM -> sum -> sigmoid ->
// global), 12: one elected lane issues every MMA.  All hand-offs are mbarriers; TMEM and shared-memory rings are freed by
// tcgen05.commit (operands) or by the reading warps (accumulators).
#include "common.cuh"
#include "tcgen05.cuh"
#include <cuda_fp16.h>

namespace riders {
namespace tailtc {
using namespace tcg;

constexpr int TC = 124;                              // output columns per strip (z needs +2, y +4 = 128 = M)
constexpr int NPOS = 136;                            // pixels per operand plane (128 + 2 read by the shifted views, padded)
constexpr int PLANE = NPOS * 16;                     // 2176 B: 8 fp16 channels of one row
constexpr int SLOT = 2 * PLANE;                      // 4352 B: one staged row (16 channels)
constexpr int NYS = 6, NZS = 4;                      // y / z operand ring slots
constexpr int ZR = 12;                               // z accumulator ring: 16 TMEM columns each
constexpr int NDH = 4;                               // head accumulator ring: 16 TMEM columns each
constexpr int LAG = 3;                               // the head of z row q is issued after the convolution of y row q + 2 + LAG
constexpr uint32_t TMEM_COLS = 256, DH_COL0 = 16 * ZR;
constexpr int LOADER_WARPS = 4, W_MMA = 12, THREADS = 13 * 32;
constexpr int WB_BYTES = 3 * 48 * 32, WH_BYTES = 3 * 16 * 32, W_BYTES = WB_BYTES + WH_BYTES;      // 4608 + 1536
constexpr int B_YFULL = 0, B_YFREE = B_YFULL + NYS, B_ZFULL = B_YFREE + NYS, B_ZFREE = B_ZFULL + NZS, B_AFULL = B_ZFREE + NZS,
              B_AFREE = B_AFULL + ZR, B_HFULL = B_AFREE + ZR, B_HFREE = B_HFULL + NDH, NBARS = B_HFREE + NDH;
constexpr int RAWD = 6;                              // raw y4 rows in flight per CTA (cp.async ring, 8 KB each)
constexpr int RAW_BYTES = RAWD * 4 * 128 * 16;
constexpr int SMEM_BYTES = (NYS + NZS) * SLOT + RAW_BYTES + W_BYTES + NBARS * 8 + 16;              // 99184
static_assert(TC + 4 <= 128 && 16 * ZR + 16 * NDH <= (int)TMEM_COLS, "strip / TMEM budget");

struct Params {
    float bias_a[16];
    float bias_b[16];
    float slope;
    int sigmoid;
    int t;                   // 0: y4 from the 3x3 phase kernels (h, w); 1: from the 2x2 phase kernels, (h + 1, w + 1)
    int h, w;
    int tiles_x, row_splits, rows_per;
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(THREADS, 2)
decoder_tail_tc_kernel(const float4* __restrict__ y4, const uint4* __restrict__ wpack, float* __restrict__ out,
                       int* __restrict__ status, const __grid_constant__ Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ys = smem;                                   // [NYS][2][NPOS] x 16 B
    uint8_t* zs = ys + NYS * SLOT;                        // [NZS][2][NPOS] x 16 B
    uint8_t* raw = zs + NZS * SLOT;                       // [RAWD][4 passes][128 loader threads] x 16 B of y4 as fetched
    uint8_t* ws = raw + RAW_BYTES;                        // packed Wb, Wh
    const uint32_t bars = smem_u32(ws + W_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ws + W_BYTES + NBARS * 8);
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = 2 * P.h, W = 2 * P.w;
    int blk = blockIdx.x;
    const int rs = blk % P.row_splits; blk /= P.row_splits;
    const int tx = blk % P.tiles_x;
    const long long n = blk / P.tiles_x;
    const int C0 = tx * TC;
    const int R0 = rs * P.rows_per, R1 = min(H, R0 + P.rows_per);          // output rows of this strip
    const int zlo = max(R0 - 1, 0), zhi = min(R1, H - 1);                  // z rows it needs (rows outside the image are zero)
    const int nz = zhi - zlo + 1, ny = nz + 2;                             // y rows zlo - 1 .. zhi + 1
    const float slope = P.slope;

    for (int i = tid; i < (NYS + NZS) * SLOT / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < W_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(ws)[i] = __ldg(wpack + i);
    if (tid == 0) {
        for (int i = 0; i < NYS; ++i) { mbar_init(bar(B_YFULL + i), LOADER_WARPS); mbar_init(bar(B_YFREE + i), 1); }
        for (int i = 0; i < NZS; ++i) { mbar_init(bar(B_ZFULL + i), 4); mbar_init(bar(B_ZFREE + i), 1); }
        for (int i = 0; i < ZR; ++i)  { mbar_init(bar(B_AFULL + i), 1); mbar_init(bar(B_AFREE + i), 4); }
        for (int i = 0; i < NDH; ++i) { mbar_init(bar(B_HFULL + i), 1); mbar_init(bar(B_HFREE + i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    float amax = 0.f;                                     // running max of |staged value| (fp16 range guard)

    if (warp < LOADER_WARPS) {
        // ===================== y rows: depth-to-space + bias_a + LeakyReLU -> fp16 -> ring =====================
        // lane = (quarter of the 16 channels, pixel of an 8-pixel group): the four quarter-lanes of a pixel read its 64 bytes,
        // the warp's store covers 8 pixels x 2 planes (256 B, two conflict-free wavefronts)
        const int q = lane >> 3, l8 = lane & 7;
        const float4 ba = make_float4(P.bias_a[4 * q + 0], P.bias_a[4 * q + 1], P.bias_a[4 * q + 2], P.bias_a[4 * q + 3]);
        const int pt = P.t, hq = P.h + pt, wq = P.w + pt;
        const int xw = min(TC, W - C0) + 4;               // y positions of this strip that can be non-zero
        const int x0 = warp * 8 + l8;                     // position in pass 0; + 32 per pass
        const int X0 = C0 - 2 + x0;
        const int b = X0 & 1;                             // column phase: fixed per thread (a pass advances 32 columns)
        // float4 offset of (Y, X) = [((Y >> 1) + a pt) wq + (X >> 1) + b pt] * 16 + (2 a + b) * 4 + q
        const float4* src = y4 + (size_t)n * hq * wq * 16 + q + b * 4;
        uint8_t* dst0 = ys + (q >> 1) * PLANE + (q & 1) * 8 + x0 * 16;
        // The raw rows are prefetched RAWD - 1 rows ahead with cp.async (no registers held while in flight); every thread reads
        // back only the 16-byte pieces it fetched itself, so cp.async.wait_group is the only synchronisation needed.
        const uint32_t raw0 = smem_u32(raw) + tid * 16;
        unsigned colmask = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            colmask |= ((unsigned)(X0 + 32 * p) < (unsigned)W && x0 + 32 * p < xw ? 1u : 0u) << p;
        auto issue_row = [&](int iy) {
            const int Y = zlo - 1 + iy, a = Y & 1;
            const bool row_ok = (unsigned)Y < (unsigned)H;
            const float4* rp = src + (row_ok ? (size_t)(((Y >> 1) + a * pt) * wq) * 16 + a * 8 : 0);
            const uint32_t d = raw0 + (iy % RAWD) * (4 * 128 * 16);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const bool ok = row_ok && ((colmask >> p) & 1);
                const float4* g = ok ? rp + (((X0 + 32 * p) >> 1) + b * pt) * 16 : y4;         // zero-filled when !ok (src size 0)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d + p * (128 * 16)), "l"(g), "r"(ok ? 16 : 0) : "memory");
            }
        };
        for (int j = 0; j < RAWD - 1; ++j) {
            if (j < ny) issue_row(j);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int iy = 0; iy < ny; ++iy) {
            if (iy + RAWD - 1 < ny) issue_row(iy + RAWD - 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(RAWD - 1) : "memory");
            const bool row_ok = (unsigned)(zlo - 1 + iy) < (unsigned)H;
            const float4* rsl = reinterpret_cast<const float4*>(raw + (iy % RAWD) * (4 * 128 * 16)) + tid;
            float4 cur[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) cur[p] = rsl[p * 128];
            const int sl = iy % NYS;
            mbar_wait(bar(B_YFREE + sl), ((iy / NYS) & 1) ^ 1);
            uint8_t* dst = dst0 + sl * SLOT;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float4 x = cur[p];
                x.x += ba.x; x.y += ba.y; x.z += ba.z; x.w += ba.w;
                x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);      // LeakyReLU, 0 <= slope <= 1
                x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
                const bool ok = row_ok && ((colmask >> p) & 1);
                if (ok) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
                uint2 o;
                o.x = ok ? pack_h2(x.x, x.y) : 0u;        // zero outside the image: the convolution pads the ACTIVATED map
                o.y = ok ? pack_h2(x.z, x.w) : 0u;
                *reinterpret_cast<uint2*>(dst + p * 32 * 16) = o;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_YFULL + sl));
        }
    } else if (warp < 8) {
        // ===================== z epilogue: TMEM -> bias_b + LeakyReLU + image mask -> fp16 -> z ring =====================
        const int lg = warp - 4, x = lg * 32 + lane;      // TMEM lane = z position; image column C0 - 1 + x
        const bool colok = (unsigned)(C0 - 1 + x) < (unsigned)W;
        const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16);
        for (int rz = 0; rz < nz; ++rz) {
            const int st = rz % ZR;
            mbar_wait(bar(B_AFULL + st), (rz / ZR) & 1);
            tc_fence_after();
            uint32_t r[16];
            TMEM_LD16(trow + 16 * (ZR - 1 - st), r);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_AFREE + st));
            uint32_t hv[8];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                float v0 = __uint_as_float(r[c]) + P.bias_b[c], v1 = __uint_as_float(r[c + 1]) + P.bias_b[c + 1];
                v0 = fmaxf(v0, v0 * slope); v1 = fmaxf(v1, v1 * slope);
                amax = fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1)));
                hv[c >> 1] = colok ? pack_h2(v0, v1) : 0u;                   // the head's zero padding outside the image
            }
            const int zsl = rz % NZS;
            mbar_wait(bar(B_ZFREE + zsl), ((rz / NZS) & 1) ^ 1);
            uint8_t* dst = zs + zsl * SLOT + x * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<uint4*>(dst + PLANE) = make_uint4(hv[4], hv[5], hv[6], hv[7]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_ZFULL + zsl));
        }
    } else if (warp < 12) {
        // ===================== head epilogue: logits[o] = Dh[o-1][0] + Dh[o][1] + Dh[o+1][2] =====================
        const int lg = warp - 8, x = lg * 32 + lane;      // TMEM lane = output position; image column C0 + x
        const bool ok = x < TC && C0 + x < W;
        const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16) + DH_COL0;
        float* op = out + (size_t)n * H * W + (ok ? C0 + x : 0);
        for (int o = R0; o < R1; ++o) {
            const int hq = min(o + 1, zhi) - zlo;
            mbar_wait(bar(B_HFULL + hq % NDH), (hq / NDH) & 1);
            tc_fence_after();
            uint32_t v0 = 0u, v1 = 0u, v2 = 0u;
            const int q = o - zlo;                        // 0 <= q < nz
            if (q >= 1) TMEM_LD1(trow + 16 * ((q - 1) % NDH) + 0, v0);
            TMEM_LD1(trow + 16 * (q % NDH) + 1, v1);
            if (q + 1 < nz) TMEM_LD1(trow + 16 * ((q + 1) % NDH) + 2, v2);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && q >= 1) mbar_arrive(bar(B_HFREE + (q - 1) % NDH));
            const float s = (__uint_as_float(v0) + __uint_as_float(v1)) + __uint_as_float(v2);
            if (ok) __stcs(op + (size_t)o * W, P.sigmoid ? (1.0f / (1.0f + expf(-s))) : s);
        }
    } else {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t ya = smem_u32(ys), za = smem_u32(zs), wa = smem_u32(ws);
            constexpr uint32_t I16 = idesc_f16(16), I32 = idesc_f16(32), I48 = idesc_f16(48);
            auto adesc = [&](uint32_t slot, int dx) { return smem_desc_nosw(slot + 16u * dx, PLANE, 128); };
            auto bdesc = [&](int dx, int block) { return smem_desc_nosw(wa + dx * 1536 + block * 512, 128, 256); };
            auto hdesc = [&](int dx) { return smem_desc_nosw(wa + WB_BYTES + dx * 512, 128, 256); };
            for (int t = 0; t < ny + LAG; ++t) {
                if (t < ny) {
                    // y row t: kernel row dy' of the 3x3 accumulates into z row t - dy' (relative); the first contribution to
                    // z row t (dy' = 0, dx = 0) overwrites its accumulator, whose previous tenant must have been read
                    const int ysl = t % NYS;
                    mbar_wait(bar(B_YFULL + ysl), (t / NYS) & 1);
                    if (t < nz) mbar_wait(bar(B_AFREE + t % ZR), ((t / ZR) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t a0 = ya + ysl * SLOT;
                    const int s0 = t % ZR;
                    if (t >= 2 && t < nz && s0 >= 2) {
                        const uint32_t c0 = tmem + 16 * (ZR - 1 - s0);            // accumulators of z rows t, t-1, t-2: adjacent
                        umma_f16(c0, adesc(a0, 0), bdesc(0, 0), I16, 0);
                        umma_f16(c0 + 16, adesc(a0, 0), bdesc(0, 1), I32, 1);
                        umma_f16(c0, adesc(a0, 1), bdesc(1, 0), I48, 1);
                        umma_f16(c0, adesc(a0, 2), bdesc(2, 0), I48, 1);
                    } else {
                        for (int b = 0; b < 3; ++b) {
                            const int rzb = t - b;
                            if (rzb < 0 || rzb >= nz) continue;
                            const uint32_t cb = tmem + 16 * (ZR - 1 - rzb % ZR);
                            for (int dx = 0; dx < 3; ++dx) umma_f16(cb, adesc(a0, dx), bdesc(dx, b), I16, (b == 0 && dx == 0) ? 0u : 1u);
                        }
                    }
                    umma_commit(bar(B_YFREE + ysl));
                    if (t >= 2) umma_commit(bar(B_AFULL + (t - 2) % ZR));         // z row t - 2 has its three kernel rows
                }
                const int hz = t - 2 - LAG;
                if (hz >= 0 && hz < nz) {
                    const int zsl = hz % NZS, ds = hz % NDH;
                    mbar_wait(bar(B_ZFULL + zsl), (hz / NZS) & 1);
                    mbar_wait(bar(B_HFREE + ds), ((hz / NDH) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t a0 = za + zsl * SLOT, d = tmem + DH_COL0 + 16 * ds;
                    for (int dx = 0; dx < 3; ++dx) umma_f16(d, adesc(a0, dx), hdesc(dx), I16, dx > 0 ? 1u : 0u);
                    umma_commit(bar(B_ZFREE + zsl));
                    umma_commit(bar(B_HFULL + ds));
                }
            }
        }
        __syncwarp();
    }
    if (status && !(amax <= 65504.f)) atomicOr(status, 1);                        // also catches NaN
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tailtc
}  // namespace riders

extern "C" int riders_decoder_tail_tc_nhwc(const float* y4, int taps, int N, int h, int w, const float* bias_a_host,
                                           const void* w_pack, const float* bias_b_host, float slope, int apply_sigmoid,
                                           float* out, int* status, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && h > 0 && w > 0, "decoder_tail_tc: bad shape");
    RIDERS_REQUIRE(taps == 2 || taps == 3, "decoder_tail_tc: taps must be 2 or 3 (got %d)", taps);
    RIDERS_REQUIRE(slope >= 0.f && slope <= 1.f, "decoder_tail_tc: LeakyReLU slope must be in [0, 1] (got %g)", (double)slope);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(y4 && w_pack && out, "decoder_tail_tc: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)y4) | ((uintptr_t)w_pack)) & 15) == 0, "decoder_tail_tc: y4 / w_pack must be 16-byte aligned");
    tailtc::Params P;
    for (int c = 0; c < 16; ++c) {
        P.bias_a[c] = bias_a_host ? bias_a_host[c] : 0.f;
        P.bias_b[c] = bias_b_host ? bias_b_host[c] : 0.f;
    }
    P.slope = slope;
    P.sigmoid = apply_sigmoid ? 1 : 0;
    P.t = taps == 2 ? 1 : 0;
    P.h = h; P.w = w;
    const int H = 2 * h, W = 2 * w;
    P.tiles_x = ceil_div(W, tailtc::TC);
    // a strip is a whole column of the map unless that leaves SMs idle (few points): then the rows are split too
    const long long strips = (long long)N * P.tiles_x;
    const long long want = 4ll * sm_count();
    int splits = strips >= want ? 1 : (int)((want + strips - 1) / strips);
    splits = splits < 1 ? 1 : splits;
    if (splits > H / 16) splits = H / 16 > 0 ? H / 16 : 1;
    P.rows_per = ceil_div(H, splits);
    P.row_splits = ceil_div(H, P.rows_per);
    const long long blocks = strips * P.row_splits;
    RIDERS_REQUIRE(blocks < (1ll << 31), "decoder_tail_tc: too many strips for one launch");
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(tailtc::decoder_tail_tc_kernel, (size_t)(tailtc::SMEM_BYTES), configured, "decoder_tail_tc: shared memory attribute")) return rc_smem;
    count_launches(1);
    tailtc::decoder_tail_tc_kernel<<<(unsigned)blocks, tailtc::THREADS, tailtc::SMEM_BYTES, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(y4), reinterpret_cast<const uint4*>(w_pack), out, status, P);
    return check_launch("decoder_tail_tc_nhwc");
}
