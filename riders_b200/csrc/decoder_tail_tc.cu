// Fused tail of RC-Net's decoder on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM): the same three
// passes as decoder_tail.cu (MultiScaleDecoder, RCNet/networks.py:678-779; DecoderBlock, utils/net_utils.py:474-569)
//
//     y      = LeakyReLU(depth_to_space(y4) + bias_a)          last block's up-convolution (sub-pixel form, see upsample.cu)
//     z      = LeakyReLU(conv3x3(y; Wb) + bias_b)              last block's 16 -> 16 convolution
//     logits = conv3x3(z; Wh)        [-> sigmoid]              output0, 16 -> 1, no bias
//
// as a streaming pipeline: a CTA owns a strip of one radar point's map (all 240 rows x up to 124 columns) and walks down it
// row by row; y and z rows live in small shared-memory rings, nothing but the input map is read (the fp16 operand map of
// upconv_tc.cu / subpixel_conv_tc.cu: 32 B per output pixel, its rows bulk-copied straight into the y ring; or a float y4 from a
// library phase convolution: 64 B, staged through registers) and nothing but the logits (4 B) is written, and no row is ever
// computed twice (the mma.sync kernel recomputes a 2-pixel halo per 12 x 50 tile: +44 % work).
//
// Both convolutions are implicit GEMMs whose A operand is a SHIFTED VIEW of a staged row:
//   * an operand row slot holds 128 (+2) consecutive pixels of one image row as two planes of 16-byte chunks (8 fp16 channels
//     each), pixel p at byte 16 p of its plane - the K-major no-swizzle canonical layout (core matrix = 8 pixels x 16 B,
//     SBO = 128 B between 8-pixel groups, LBO = plane stride between the two K chunks).  An unswizzled descriptor may start at
//     any 16-byte boundary, so the operand of horizontal tap dx is the same slot addressed 16 dx bytes further: M = 128 pixels,
//     K = 16 input channels, no im2col, no copies;
//   * the vertical taps ride in N: for y row r the B operand of tap dx stacks the three kernel rows, N = 3 x 16 output channels,
//     and the three 16-column blocks of D accumulate into the TMEM accumulators of z rows r+1, r, r-1 (a ring of twelve
//     16-column accumulators laid out so that the three targets are adjacent columns; every MMA accumulates, the epilogue
//     zeroes an accumulator right after reading it).  One z row costs 3 MMAs (128 x 48 x 16; 6 where the ring wraps) instead of
//     the 9 x 2 x 8 mma.sync.m16n8k8 per 16 pixels of the legacy kernel, and its epilogue reads 16 finished columns;
//   * the 16 -> 1 head is the same construction with N = 16 (columns 0..2 = the head's three kernel rows): three MMAs per z row,
//     logits[o] = Dh[o-1][0] + Dh[o][1] + Dh[o+1][2] read back one column each.
// Operands are fp16 (11-bit significand, the same as the TF32 the cuDNN flow and decoder_tail.cu use) with fp32 accumulation;
// every staged value passes a running max and a value outside the fp16 range raises `status` (the host then re-runs the
// mma.sync TF32 kernel, ops.decoder_tail).
//
// Warp roles (2 CTAs per SM, 256 TMEM columns each; see Roles below): staging warps (float input: global -> bias + LeakyReLU ->
// fp16 -> ring; fp16 input: one lane bulk-copies rows), z epilogue groups (TMEM -> bias + LeakyReLU + image mask -> fp16 -> z
// ring), head epilogue groups (TMEM -> sum -> sigmoid -> global), one MMA-issuing warp per convolution.  All hand-offs are
// mbarriers; TMEM and shared-memory rings are freed by tcgen05.commit (operands) or by the reading warps (accumulators).
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
constexpr int NYS = 8, NZS = 6;                      // y / z operand ring slots
constexpr int ZR = 12;                               // z accumulator ring: 16 TMEM columns each
constexpr int NDH = 4;                               // head accumulator ring: 16 TMEM columns each
constexpr uint32_t TMEM_COLS = 256, DH_COL0 = 16 * ZR;
constexpr int WB_BYTES = 3 * 48 * 32, WH_BYTES = 3 * 16 * 32, W_BYTES = WB_BYTES + WH_BYTES;      // 4608 + 1536
constexpr int B_YFULL = 0, B_YFREE = B_YFULL + NYS, B_ZFULL = B_YFREE + NYS, B_ZFREE = B_ZFULL + NZS, B_AFULL = B_ZFREE + NZS,
              B_AFREE = B_AFULL + ZR, B_HFULL = B_AFREE + ZR, B_HFREE = B_HFULL + NDH, NBARS = B_HFREE + NDH;
// Warp roles.  Float input (y4 from a library convolution): warps 0-7 stage, 8-11 / 12-15 epilogues, 16 issues the MMAs.
// fp16 input (the finished map of riders_subpixel_conv_tc_half_nhwc): nothing to compute while staging - lane 0 of warp 0
// bulk-copies rows straight into the operand ring, warp 1 issues the MMAs, warps 4-7 / 8-11 are the epilogues.
template <bool HALF_IN> struct Roles {
    // z epilogue groups (four warps each, alternating z rows): a row's epilogue is one serial chain of TMEM / shared-memory
    // round trips per warp, so with the staging arithmetic gone (fp16 input) ONE group is the slowest stage of the pipeline
    static constexpr int EPI1_GROUPS = HALF_IN ? 2 : 1;
    // head epilogue groups: two were built and measured NO faster (0.79 vs 0.76 ms) - by then the shared-memory pipe (tensor-core
    // operand reads: 3.5 + 3 MMAs x ~40 wavefronts per row and CTA) is ~85 % busy, which no amount of warps changes
    static constexpr int EPI2_GROUPS = 1;
    static constexpr int LOADER_WARPS = HALF_IN ? 1 : 8, W_EPI1 = HALF_IN ? 4 : 8, W_EPI2 = W_EPI1 + 4 * EPI1_GROUPS, W_MMA = HALF_IN ? 1 : 16;
    static constexpr int W_MMA_HEAD = HALF_IN ? 2 : 17;             // the head's MMAs have their own issuing warp (see there)
    static constexpr int THREADS = HALF_IN ? (4 + 4 * EPI1_GROUPS + 4 * EPI2_GROUPS) * 32 : 18 * 32;
    static constexpr int YSLOTS = NYS + (HALF_IN ? 1 : 0);          // + a slot that stays zero: the rows outside the image
    static constexpr int SMEM_BYTES = (YSLOTS + NZS) * SLOT + W_BYTES + NBARS * 8 + 16;
};
static_assert(TC + 4 <= 128 && 16 * ZR + 16 * NDH <= (int)TMEM_COLS, "strip / TMEM budget");

struct Params {
    float bias_a[16];
    float bias_b[16];
    float slope;
    int sigmoid;
    int t;                   // 0: y4 from the 3x3 phase kernels (h, w); 1: from the 2x2 phase kernels, (h + 1, w + 1)
    int h, w;
    int tiles_x, row_splits, rows_per;
    int dbg;                 // timing experiments only (riders_debug_set("tail_tc_dbg")): 1 = one head MMA, 2 = one convolution MMA per row
};

__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }

template <bool HALF_IN>
__global__ void __launch_bounds__(Roles<HALF_IN>::THREADS, 2)
decoder_tail_tc_kernel(const float4* __restrict__ y4, const uint4* __restrict__ wpack, float* __restrict__ out,
                       int* __restrict__ status, const __grid_constant__ Params P) {
    constexpr int EPI1_GROUPS = Roles<HALF_IN>::EPI1_GROUPS, EPI2_GROUPS = Roles<HALF_IN>::EPI2_GROUPS, W_MMA_HEAD = Roles<HALF_IN>::W_MMA_HEAD;
    constexpr int LOADER_WARPS = Roles<HALF_IN>::LOADER_WARPS, W_EPI1 = Roles<HALF_IN>::W_EPI1, W_EPI2 = Roles<HALF_IN>::W_EPI2,
                  W_MMA = Roles<HALF_IN>::W_MMA, THREADS = Roles<HALF_IN>::THREADS, YSLOTS = Roles<HALF_IN>::YSLOTS;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ys = smem;                                   // [YSLOTS][2][NPOS] x 16 B
    uint8_t* zs = ys + YSLOTS * SLOT;                     // [NZS][2][NPOS] x 16 B
    uint8_t* ws = zs + NZS * SLOT;                        // packed Wb, Wh
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

    for (int i = tid; i < (YSLOTS + NZS) * SLOT / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < W_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(ws)[i] = __ldg(wpack + i);
    if (tid == 0) {
        for (int i = 0; i < NYS; ++i) { mbar_init(bar(B_YFULL + i), HALF_IN ? 1 : LOADER_WARPS); mbar_init(bar(B_YFREE + i), 1); }
        for (int i = 0; i < NZS; ++i) { mbar_init(bar(B_ZFULL + i), 4); mbar_init(bar(B_ZFREE + i), 1); }
        for (int i = 0; i < ZR; ++i)  { mbar_init(bar(B_AFULL + i), 1); mbar_init(bar(B_AFREE + i), 4); }
        for (int i = 0; i < NDH; ++i) { mbar_init(bar(B_HFULL + i), 1); mbar_init(bar(B_HFREE + i), EPI2_GROUPS == 1 ? 4 : 12); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp >= W_EPI1 && warp < W_EPI1 + 4) {            // z accumulators start from zero (their readers keep them so)
        for (int i = 0; i < ZR; ++i) tmem_zero16(tmem + ((uint32_t)((warp - W_EPI1) * 32) << 16) + 16 * i);
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    __half2 amax2 = __float2half2_rn(0.f);                // running max of |staged fp16 value| (range guard: inf = overflow)

    if (HALF_IN && warp == 0) {
        // ===================== y rows (fp16 input): one bulk copy per plane straight into the operand ring =====================
        if (lane == 0) {
            const __half* yh = reinterpret_cast<const __half*>(y4);
            const int xs = max(C0 - 2, 0), xe = min(W, C0 + TC + 2);              // image columns of this strip (with the halo)
            const uint32_t bytes = (uint32_t)(xe - xs) * 16, dpos = (uint32_t)(xs - (C0 - 2)) * 16;
            const __half* src = yh + ((size_t)n * H * 2 * W + xs) * 8;
            int sl = 0, ph = 1;
            for (int iy = 0; iy < ny; ++iy) {
                const int Y = zlo - 1 + iy;
                if ((unsigned)Y >= (unsigned)H) continue;                          // the MMA reads the zero slot for those
                mbar_wait_sleep(bar(B_YFREE + sl), ph, 100);
                mbar_expect_tx(bar(B_YFULL + sl), 2 * bytes);
                const uint32_t dst = smem_u32(ys + sl * SLOT) + dpos;
                const __half* g = src + (size_t)Y * 2 * W * 8;
                bulk_g2s(dst, g, bytes, bar(B_YFULL + sl));
                bulk_g2s(dst + PLANE, g + (size_t)W * 8, bytes, bar(B_YFULL + sl));
                if (++sl == NYS) { sl = 0; ph ^= 1; }
            }
        }
    } else if (!HALF_IN && warp < LOADER_WARPS) {
        // ===================== y rows: depth-to-space + bias_a + LeakyReLU -> fp16 -> ring =====================
        // lane = (quarter of the 16 channels, pixel of an 8-pixel group): the four quarter-lanes of a pixel read its 64 bytes,
        // the warp's store covers 8 pixels x 2 planes (256 B, two conflict-free wavefronts).  8 warps x 8 pixels x 2 passes.
        const int q = lane >> 3, l8 = lane & 7;
        const float4 ba = make_float4(P.bias_a[4 * q + 0], P.bias_a[4 * q + 1], P.bias_a[4 * q + 2], P.bias_a[4 * q + 3]);
        const int pt = P.t, hq = P.h + pt, wq = P.w + pt;
        const int xw = min(TC, W - C0) + 4;               // y positions of this strip that can be non-zero
        const int x0 = warp * 8 + l8;                     // position in pass 0; + 64 in pass 1
        const int X0 = C0 - 2 + x0;
        const int b = X0 & 1;                             // column phase: fixed per thread (a pass advances 64 columns)
        // float4 offset of (Y, X) = [((Y >> 1) + a pt) wq + (X >> 1) + b pt] * 16 + (2 a + b) * 4 + q
        const float4* src = y4 + (size_t)n * hq * wq * 16 + q + b * 4;
        const bool ok0 = (unsigned)X0 < (unsigned)W && x0 < xw, ok1 = (unsigned)(X0 + 64) < (unsigned)W && x0 + 64 < xw;
        const float4* src0 = src + ((X0 >> 1) + b * pt) * 16;                    // column part of the two passes
        const float4* src1 = src + (((X0 + 64) >> 1) + b * pt) * 16;
        uint8_t* dst0 = ys + (q >> 1) * PLANE + (q & 1) * 8 + x0 * 16;
        // Rows go global -> registers -> operand ring (a staging copy in shared memory costs more shared-memory bandwidth than
        // the MMAs themselves: the kernel is bound by that pipe).  Two rows are in flight per thread (three register buffers,
        // the loop is unrolled by three so that they stay registers) and a bulk L2 prefetch runs PF rows ahead, so the loads in
        // flight only have to cover the L2 latency.
        const size_t img = (size_t)n * hq * wq * 16;
        const int lc0 = max((C0 - 2) >> 1, 0), lc1 = min(((C0 - 2 + xw - 1) >> 1) + pt, wq - 1);      // low-resolution columns used
        auto prefetch_row = [&](int Y) {                  // the low-resolution row behind output rows Y (and its sibling)
            if ((unsigned)Y >= (unsigned)H) return;
            const int L = (Y >> 1) + (Y & 1) * pt;
            const float4* g = y4 + img + ((size_t)L * wq + lc0) * 16;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"((lc1 - lc0 + 1) * 256) : "memory");
        };
        constexpr int PF = 10;
        if (tid == 0)
            for (int Y = zlo - 1; Y < zlo - 1 + PF; ++Y)
                if (Y == zlo - 1 || (Y & 1) == pt) prefetch_row(Y);              // each low-resolution row once
        // row part of the address, kept incrementally: from an even row to the next one the phase advances by 8 float4 (and the
        // low-resolution row by pt), from an odd one the phase goes back and the low-resolution row advances by 1 - pt
        const int d_even = pt * wq * 16 + 8, d_odd = (1 - pt) * wq * 16 - 8;
        int Yl = zlo - 1;                                  // row the next load_row call fetches
        int ro = (((Yl >> 1) + (Yl & 1) * pt) * wq) * 16 + (Yl & 1) * 8;      // (consistent for Yl = -1 too; never dereferenced then)
        auto load_row = [&](float4 (&v)[2]) {
            const bool row_ok = (unsigned)Yl < (unsigned)H;
            v[0] = row_ok && ok0 ? __ldcs(src0 + ro) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[1] = row_ok && ok1 ? __ldcs(src1 + ro) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid == 0 && ((Yl + PF) & 1) == pt) prefetch_row(Yl + PF);
            ro += (Yl & 1) ? d_odd : d_even;
            ++Yl;
        };
        int sl = 0, ph = 1;                               // operand slot and its free phase
        auto stage_row = [&](int iy, const float4 (&v)[2]) {
            const bool row_ok = (unsigned)(zlo - 1 + iy) < (unsigned)H;
            mbar_wait_sleep(bar(B_YFREE + sl), ph, 200);
            uint8_t* dst = dst0 + sl * SLOT;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float4 x = v[p];
                x.x += ba.x; x.y += ba.y; x.z += ba.z; x.w += ba.w;
                x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);      // LeakyReLU, 0 <= slope <= 1
                x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
                const bool ok = row_ok && (p ? ok1 : ok0);
                const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
                uint2 o;
                o.x = ok ? h2_bits(h0) : 0u;              // zero outside the image: the convolution pads the ACTIVATED map
                o.y = ok ? h2_bits(h1) : 0u;
                amax2 = __hmax2_nan(amax2, __hmax2_nan(__habs2(bits_h2(o.x)), __habs2(bits_h2(o.y))));
                *reinterpret_cast<uint2*>(dst + p * 64 * 16) = o;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_YFULL + sl));
            if (++sl == NYS) { sl = 0; ph ^= 1; }
        };
        float4 r0[2], r1[2], r2[2];
        load_row(r0);
        if (1 < ny) load_row(r1);
        for (int iy = 0; iy < ny; iy += 3) {
            if (iy + 2 < ny) load_row(r2);
            stage_row(iy, r0);
            if (iy + 1 < ny) {
                if (iy + 3 < ny) load_row(r0);
                stage_row(iy + 1, r1);
            }
            if (iy + 2 < ny) {
                if (iy + 4 < ny) load_row(r1);
                stage_row(iy + 2, r2);
            }
        }
    } else if (warp >= W_EPI1 && warp < W_EPI2) {
        // ===================== z epilogue: TMEM -> bias_b + LeakyReLU + image mask -> fp16 -> z ring =====================
        static_assert(ZR % EPI1_GROUPS == 0 && NZS % EPI1_GROUPS == 0, "a group keeps the parity of its ring slots");
        const int lg = (warp - W_EPI1) & 3, grp = (warp - W_EPI1) >> 2, x = lg * 32 + lane;   // TMEM lane = z position; image column C0 - 1 + x
        const bool colok = (unsigned)(C0 - 1 + x) < (unsigned)W;
        const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16);
        int st = grp, stph = 0, zsl = grp, zph = 1;       // group g takes z rows g, g + EPI1_GROUPS, ...
        for (int rz = grp; rz < nz; rz += EPI1_GROUPS) {
            mbar_wait_sleep(bar(B_AFULL + st), stph, 100);
            tc_fence_after();
            uint32_t r[16];
            TMEM_LD16(trow + 16 * (ZR - 1 - st), r);
            tmem_wait_ld();
            tmem_zero16(trow + 16 * (ZR - 1 - st));       // the next tenant's MMAs all accumulate
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_AFREE + st));
            uint32_t hv[8];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                float v0 = __uint_as_float(r[c]) + P.bias_b[c], v1 = __uint_as_float(r[c + 1]) + P.bias_b[c + 1];
                v0 = fmaxf(v0, v0 * slope); v1 = fmaxf(v1, v1 * slope);
                const __half2 hh = __floats2half2_rn(v0, v1);
                amax2 = __hmax2_nan(amax2, __habs2(hh));
                hv[c >> 1] = colok ? h2_bits(hh) : 0u;                       // the head's zero padding outside the image
            }
            mbar_wait_sleep(bar(B_ZFREE + zsl), zph, 100);
            uint8_t* dst = zs + zsl * SLOT + x * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<uint4*>(dst + PLANE) = make_uint4(hv[4], hv[5], hv[6], hv[7]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_ZFULL + zsl));
            if ((st += EPI1_GROUPS) >= ZR) { st -= ZR; stph ^= 1; }
            if ((zsl += EPI1_GROUPS) >= NZS) { zsl -= NZS; zph ^= 1; }
        }
    } else if (warp >= W_EPI2 && warp < W_EPI2 + 4 * EPI2_GROUPS) {
        // ===================== head epilogue: logits[o] = Dh[o-1][0] + Dh[o][1] + Dh[o+1][2] =====================
        // One group: rows in order, Dh[q-1] is free once row q has been read.  Two groups (alternating rows, running concurrently):
        // an accumulator is free when ALL its reader rows have arrived (3 rows x 4 warps = 12); the readers that do not exist at
        // the top of the strip are arrived for up front.
        const int lg = (warp - W_EPI2) & 3, grp = (warp - W_EPI2) >> 2, x = lg * 32 + lane;   // TMEM lane = output position; image column C0 + x
        const bool ok = x < TC && C0 + x < W;
        const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16) + DH_COL0;
        const int q0 = R0 - zlo, q1 = R1 - zlo;           // q = o - zlo: z row of the output row, 0 <= q < nz
        if (EPI2_GROUPS > 1 && grp == 0 && lane == 0) {
            for (int j = 0; j < 2 && j < nz; ++j)
                for (int r = j - 1; r <= j + 1; ++r)
                    if (r < q0) mbar_arrive(bar(B_HFREE + j % NDH));
        }
        float* op = out + (size_t)n * H * W + (size_t)(R0 + grp) * W + (ok ? C0 + x : 0);
        for (int q = q0 + grp; q < q1; q += EPI2_GROUPS, op += EPI2_GROUPS * W) {
            const int hq = min(q + 1, nz - 1);
            mbar_wait_sleep(bar(B_HFULL + hq % NDH), (hq / NDH) & 1, 100);
            tc_fence_after();
            uint32_t v0 = 0u, v1 = 0u, v2 = 0u;
            if (q >= 1) TMEM_LD1(trow + 16 * ((q - 1) % NDH) + 0, v0);
            TMEM_LD1(trow + 16 * (q % NDH) + 1, v1);
            if (q + 1 < nz) TMEM_LD1(trow + 16 * ((q + 1) % NDH) + 2, v2);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (q >= 1) mbar_arrive(bar(B_HFREE + (q - 1) % NDH));
                if (EPI2_GROUPS > 1) {
                    mbar_arrive(bar(B_HFREE + q % NDH));
                    if (q + 1 < nz) mbar_arrive(bar(B_HFREE + (q + 1) % NDH));
                }
            }
            const float sum = (__uint_as_float(v0) + __uint_as_float(v1)) + __uint_as_float(v2);
            if (ok) __stcs(op, P.sigmoid ? (1.0f / (1.0f + expf(-sum))) : sum);
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: the whole warp runs the (warp-uniform) control flow, one elected lane issues ======
        const uint32_t ya = smem_u32(ys), wa = smem_u32(ws);
        constexpr uint32_t I16 = idesc_f16(16), I32 = idesc_f16(32), I48 = idesc_f16(48);
        // descriptors differ only in their start-address field (units of 16 B): + SLOT / 16 per ring slot, + 1 per pixel of
        // horizontal shift; B: + 96 per tap dx (1536 B), + 32 per kernel row (512 B)
        const uint64_t yd0 = smem_desc_nosw(ya, PLANE, 128), bd0 = smem_desc_nosw(wa, 128, 256);
        int ysl = 0, yph = 0, zr = 0, zrph = 1;           // y operand slot / phase, accumulator slot of z row t / its free phase
        for (int t = 0; t < ny; ++t) {
            {
                // y row t: kernel row dy' of the 3x3 accumulates into z row t - dy' (relative); the accumulator of z row t must
                // have been read (and zeroed) by the epilogue of its previous tenant
                const bool yring = !HALF_IN || (unsigned)(zlo - 1 + t) < (unsigned)H;      // fp16 input: rows outside = the zero slot
                if (yring) mbar_wait(bar(B_YFULL + ysl), yph);
                if (t < nz) mbar_wait(bar(B_AFREE + zr), zrph);
                tc_fence_after();
                const uint64_t ad = yd0 + (uint64_t)((yring ? ysl : NYS) * (SLOT / 16));
                const int zr2 = zr >= 2 ? zr - 2 : zr + ZR - 2;                    // accumulator slot of z row t - 2
                if (t >= 2 && t < nz) {
                    // all three kernel rows have a target: their accumulators (z rows t, t-1, t-2) are adjacent columns unless
                    // the ring wraps between them
                    if (elect_one()) {
                        if (zr >= 2) {
                            const uint32_t c0 = tmem + 16 * (ZR - 1 - zr);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                if (dx == 0 || !(P.dbg & 2)) umma_f16(c0, ad + dx, bd0 + 96 * dx, I48, 1);
                        } else if (zr == 1) {               // rows t, t-1 -> columns 160..191, row t-2 -> column 0
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                umma_f16(tmem + 16 * (ZR - 2), ad + dx, bd0 + 96 * dx, I32, 1);
                                umma_f16(tmem, ad + dx, bd0 + 96 * dx + 64, I16, 1);
                            }
                        } else {                            // row t -> column 176, rows t-1, t-2 -> columns 0..31
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                umma_f16(tmem + 16 * (ZR - 1), ad + dx, bd0 + 96 * dx, I16, 1);
                                umma_f16(tmem, ad + dx, bd0 + 96 * dx + 32, I32, 1);
                            }
                        }
                        if (yring) umma_commit(bar(B_YFREE + ysl));
                        umma_commit(bar(B_AFULL + zr2));                           // z row t - 2 has its three kernel rows
                    }
                } else {
                    if (elect_one()) {
                        for (int b = 0; b < 3; ++b) {
                            const int rzb = t - b;
                            if (rzb < 0 || rzb >= nz) continue;
                            const int zb = zr >= b ? zr - b : zr + ZR - b;
                            const uint32_t cb = tmem + 16 * (ZR - 1 - zb);
                            for (int dx = 0; dx < 3; ++dx) umma_f16(cb, ad + dx, bd0 + (uint64_t)(96 * dx + 32 * b), I16, 1);
                        }
                        if (yring) umma_commit(bar(B_YFREE + ysl));
                        if (t >= 2) umma_commit(bar(B_AFULL + zr2));
                    }
                }
                __syncwarp();
                if (yring && ++ysl == NYS) { ysl = 0; yph ^= 1; }
                if (++zr == ZR) { zr = 0; zrph ^= 1; }
            }
        }
    } else if (warp == W_MMA_HEAD) {
        // ===================== MMA issuer of the head: its own warp.  One warp issuing both convolutions was the slowest
        // stage of the pipeline (~170 instructions of dependent barrier / descriptor arithmetic per row); the two streams
        // are independent (tcgen05.commit tracks the MMAs of the issuing thread) ======
        const uint32_t za = smem_u32(zs), wa = smem_u32(ws);
        constexpr uint32_t I16 = idesc_f16(16);
        const uint64_t zd0 = smem_desc_nosw(za, PLANE, 128), hd0 = smem_desc_nosw(wa + WB_BYTES, 128, 256);
        int hzs = 0, hzph = 0, hds = 0, hdph = 1;         // z operand slot / phase, head accumulator slot / its free phase
        for (int hz = 0; hz < nz; ++hz) {
            mbar_wait(bar(B_ZFULL + hzs), hzph);
            mbar_wait(bar(B_HFREE + hds), hdph);
            tc_fence_after();
            const uint64_t ad = zd0 + (uint64_t)(hzs * (SLOT / 16));
            const uint32_t d = tmem + DH_COL0 + 16 * hds;
            if (elect_one()) {
                umma_f16(d, ad, hd0, I16, 0);
                if (!(P.dbg & 1)) {
                    umma_f16(d, ad + 1, hd0 + 32, I16, 1);
                    umma_f16(d, ad + 2, hd0 + 64, I16, 1);
                }
                umma_commit(bar(B_ZFREE + hzs));
                umma_commit(bar(B_HFULL + hds));
            }
            __syncwarp();
            if (++hzs == NZS) { hzs = 0; hzph ^= 1; }
            if (++hds == NDH) { hds = 0; hdph ^= 1; }
        }
    }
    {
        const float2 m = __half22float2(amax2);
        if (status && !(fmaxf(m.x, m.y) <= 65504.f)) atomicOr(status, 1);         // inf: overflow; also catches NaN
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

static int g_dbg = 0;
}  // namespace tailtc
void set_tail_tc_dbg(int v) { tailtc::g_dbg = v; }
}  // namespace riders

static int decoder_tail_tc_launch(const void* in, bool half_in, int taps, int N, int h, int w, const float* bias_a_host,
                                  const void* w_pack, const float* bias_b_host, float slope, int apply_sigmoid,
                                  float* out, int* status, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && h > 0 && w > 0, "decoder_tail_tc: bad shape");
    RIDERS_REQUIRE(taps == 2 || taps == 3, "decoder_tail_tc: taps must be 2 or 3 (got %d)", taps);
    RIDERS_REQUIRE(slope >= 0.f && slope <= 1.f, "decoder_tail_tc: LeakyReLU slope must be in [0, 1] (got %g)", (double)slope);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(in && w_pack && out, "decoder_tail_tc: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)in) | ((uintptr_t)w_pack)) & 15) == 0, "decoder_tail_tc: input / w_pack must be 16-byte aligned");
    tailtc::Params P;
    for (int c = 0; c < 16; ++c) {
        P.bias_a[c] = bias_a_host ? bias_a_host[c] : 0.f;
        P.bias_b[c] = bias_b_host ? bias_b_host[c] : 0.f;
    }
    P.slope = slope;
    P.sigmoid = apply_sigmoid ? 1 : 0;
    P.t = taps == 2 ? 1 : 0;
    P.h = h; P.w = w;
    P.dbg = tailtc::g_dbg;
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
    count_launches(1);
    if (half_in) {
        static size_t configured[64] = {0};
        if (int rc_smem = ensure_dynamic_smem(tailtc::decoder_tail_tc_kernel<true>, (size_t)(tailtc::Roles<true>::SMEM_BYTES), configured, "decoder_tail_tc: shared memory attribute")) return rc_smem;
        tailtc::decoder_tail_tc_kernel<true><<<(unsigned)blocks, tailtc::Roles<true>::THREADS, tailtc::Roles<true>::SMEM_BYTES, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(in), reinterpret_cast<const uint4*>(w_pack), out, status, P);
    } else {
        static size_t configured[64] = {0};
        if (int rc_smem = ensure_dynamic_smem(tailtc::decoder_tail_tc_kernel<false>, (size_t)(tailtc::Roles<false>::SMEM_BYTES), configured, "decoder_tail_tc: shared memory attribute")) return rc_smem;
        tailtc::decoder_tail_tc_kernel<false><<<(unsigned)blocks, tailtc::Roles<false>::THREADS, tailtc::Roles<false>::SMEM_BYTES, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(in), reinterpret_cast<const uint4*>(w_pack), out, status, P);
    }
    return check_launch("decoder_tail_tc_nhwc");
}

extern "C" int riders_decoder_tail_tc_nhwc(const float* y4, int taps, int N, int h, int w, const float* bias_a_host,
                                           const void* w_pack, const float* bias_b_host, float slope, int apply_sigmoid,
                                           float* out, int* status, void* stream) {
    return decoder_tail_tc_launch(y4, false, taps, N, h, w, bias_a_host, w_pack, bias_b_host, slope, apply_sigmoid, out, status, stream);
}

extern "C" int riders_decoder_tail_tc_half_nhwc(const void* yh, int N, int h, int w, const void* w_pack, const float* bias_b_host,
                                                float slope, int apply_sigmoid, float* out, int* status, void* stream) {
    return decoder_tail_tc_launch(yh, true, 3, N, h, w, nullptr, w_pack, bias_b_host, slope, apply_sigmoid, out, status, stream);
}
