// The last decoder block's up-convolution in ONE kernel, from the RAW output of the convolution in front of it to the fp16
// operand map of the tcgen05 decoder tail (MultiScaleDecoder deconv1.conv -> deconv0.deconv, RCNet/networks.py:678-779,
// utils/net_utils.py:474-569; phase algebra in upsample.cu):
//
//     xa = LeakyReLU(x + bias_p)                       the previous convolution's bias + activation (x: its raw output, 32 ch)
//     y4 = conv2x2(xa, W2, padding 1)                  sub-pixel form of resize x2 + 3x3 convolution: 4 phases x 16 channels
//     y  = LeakyReLU(depth_to_space(y4) + bias_a)      -> fp16, written as yh [N][2h][2 planes][2w][8]
//
// It replaces three passes of the library flow (bias_act over the 32-channel map: 1.6 GB read + 1.6 GB written per 2048 points;
// the phase convolution; the depth-to-space / bias / activation pass) and improves on riders_subpixel_conv_tc_half_nhwc, whose
// TMA-fed TF32 MMAs read the activated fp32 map: that kernel loads every input row four times (one shifted copy per tap) and is
// bound by the shared-memory pipe.  Here the rows go global -> registers (+ bias, LeakyReLU, fp16) -> shared memory ONCE, in the
// unswizzled K-major plane layout of decoder_tail_tc.cu: pixel p of a staged row at byte 16 p of each 8-channel plane, rows 56
// pixels apart.  An unswizzled operand descriptor may start at any 16-byte boundary, so all four taps are shifted views of the
// same staged rows: tap (u, v) of the y4 row pair (I, I + 1) is the M = 128 operand starting at row I - 1 + u, pixel v
// (positions 0..55 of the pair's first row, 56..111 of its second).  Per pair: 8 tcgen05.mma kind::f16 (4 taps x 2 halves of the
// 32 input channels, M 128, N 64, K 16) into one of four 64-column TMEM accumulators; two groups of four warps finish the pairs
// alternately (TMEM -> depth-to-space, + bias_a, LeakyReLU, fp16 -> staging -> one bulk store per finished output row).
//
// A CTA streams down one radar point's map in batches of 8 y4 rows (9 input rows, double-buffered); persistent over the points.
// fp16 operands (11-bit significand, like the TF32 of the flow this replaces), fp32 accumulation; every staged / written value
// passes a running |max| and a value outside the fp16 range raises `status` (the host then answers from the TF32 library flow).
#include "common.cuh"
#include "tcgen05.cuh"
#include <cuda_fp16.h>

namespace riders {
namespace upconv {
using namespace tcg;

constexpr int BROWS = 8;                              // output (y4) rows per batch
constexpr int NBUF = 2, NACC = 4;
constexpr int LOADER_WARPS = 8, W_MMA = 8, W_EPI = 12, EPI_GROUPS = 2, THREADS = (W_EPI + 4 * EPI_GROUPS) * 32;   // warps 9..11 idle
constexpr int B_BFULL = 0, B_BFREE = B_BFULL + NBUF, B_DFULL = B_BFREE + NBUF, B_DFREE = B_DFULL + NACC, NBARS = B_DFREE + NACC;

// Two instances.  <32, 16, 56, true>: the LAST block (deconv0), two y4 rows of <= 55 positions per M = 128 tile, fp16 operand
// map out (the tail's input).  <64, 32, 28, false>: the block before it (deconv1: 60 x 25 -> 120 x 50), four y4 rows of <= 27
// positions per tile, float32 channels-last out, written into the first COUT channels of the block's concat buffer (the skip
// is already in place behind them) - there it replaces the library phase convolution AND the depth-to-space pass.
// <64, 32, 56, false, K3 = true>: the 3x3 convolution between the two (deconv1.conv, 64 -> 32 channels at 120 x 50), the largest
// library convolution of the step: same staging and views with nine taps ((dy, dx) = a row and a pixel shift), one output
// row pair per tile, N = 32, raw float32 out (its consumer - the first instance - applies the bias and the activation).
template <int CIN_, int COUT_, int PITCH_, bool HALF_OUT_, bool K3_ = false> struct Cfg {
    static constexpr int CIN = CIN_, COUT = COUT_, PITCH = PITCH_;
    static constexpr bool HALF_OUT = HALF_OUT_, K3 = K3_;
    static constexpr int NOUT = K3 ? COUT_ : 4 * COUT_;                   // 3x3: plain output channels; 2x2: four phases
    static constexpr int TAPS = K3 ? 9 : 4, TAPW = K3 ? 3 : 2;            // taps, taps per kernel row
    static constexpr int XROWS = BROWS + TAPW - 1;                        // input rows staged per batch of BROWS output (y4) rows
    static constexpr int RPT = 112 / PITCH;                               // output (y4) rows per M = 128 tile
    static constexpr int NPLANES = CIN / 8, KSTEPS = CIN / 16;
    // staged rows + the reach of the last view.  The staging stores are 16 bytes per lane, processed a quarter warp at a time:
    // its 8 lanes are 1 pixel x 8 planes (CIN 64) or 2 adjacent pixels x 4 planes (CIN 32), so the plane stride is made 16 /
    // 32 bytes mod 128 - the 8 lanes then fall into 8 different 16-byte bank groups (a stride of 0 mod 128 was an 8- / 4-way
    // conflict on every store: ncu counted 3x the ideal shared-memory wavefronts)
    static constexpr int PLANE_POS = (((BROWS - RPT + TAPW - 1) * PITCH + (TAPW - 1) + 128) + 7) / 8 * 8 + (CIN / 8 == 8 ? 1 : 2);
    static constexpr int PLANE = PLANE_POS * 16;
    static constexpr int BATCH_BYTES = NPLANES * PLANE;
    static constexpr int W_BYTES = TAPS * KSTEPS * NOUT * 32;             // taps x K steps x (NOUT rows x 16 fp16)
    // epilogue staging per group.  fp16 out: 2 RPT output rows x 2 planes x W pixels x 16 B (bulk-stored row by row);
    // float out: per warp 32 pixels x 128 B (+ 16 B pad) for the transposing store
    static constexpr int STAGE_BYTES = HALF_OUT ? 2 * RPT * 2 * 2 * (PITCH - 1) * 16 : 4 * 32 * 144;
    static constexpr int SMEM_BYTES = NBUF * BATCH_BYTES + W_BYTES + EPI_GROUPS * STAGE_BYTES + NBARS * 8 + 16;
    static constexpr uint32_t TMEM_COLS = NACC * NOUT;
    static_assert(BROWS % RPT == 0 && RPT * PITCH <= 128 && XROWS * PITCH <= PLANE_POS && NPLANES * (PITCH - 2) <= 2 * LOADER_WARPS * 32,
                  "tile geometry");
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
};

struct Params {
    const float4* x;          // (N, h, w, CIN) fp32, raw output of the previous convolution
    const uint4* w_pack;
    __half* yh;               // HALF_OUT: [N][2h][2][2w][8]
    float* out;               // !HALF_OUT: (N, 2h, 2w, ctot) float32, channels [0, COUT) written
    int ctot;
    int* status;
    float bias_p[64];         // the previous convolution's bias (zeros if none)
    float slope_p;            // its LeakyReLU slope (1 = identity: x is already activated)
    float bias_a[32];
    float slope;
    int N, h, w;
};

__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

template <class C>
__global__ void __launch_bounds__(THREADS, 1)
upconv_tc_kernel(const __grid_constant__ Params P) {
    constexpr int CIN = C::CIN, NOUT = C::NOUT, PITCH = C::PITCH, RPT = C::RPT, NPLANES = C::NPLANES, KSTEPS = C::KSTEPS;
    constexpr int PLANE = C::PLANE, BATCH_BYTES = C::BATCH_BYTES, W_BYTES = C::W_BYTES, STAGE_BYTES = C::STAGE_BYTES;
    constexpr int XROWS = C::XROWS, TAPS = C::TAPS, TAPW = C::TAPW;
    constexpr bool K3 = C::K3;
    constexpr uint32_t TMEM_COLS = C::TMEM_COLS;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* xb = smem;                                   // [NBUF][NPLANES][PLANE_POS] x 16 B
    uint8_t* ws = xb + NBUF * BATCH_BYTES;                // packed W2
    uint8_t* stg = ws + W_BYTES;                          // [EPI_GROUPS][4 rows][2 planes][W] x 16 B
    const uint32_t bars = smem_u32(stg + EPI_GROUPS * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg + EPI_GROUPS * STAGE_BYTES + NBARS * 8);
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = P.h, w = P.w, H = 2 * h, W = 2 * w;
    const int orows = K3 ? h : h + 1;                     // output rows of the GEMM: y4 has h + 1 rows, the 3x3 output h
    const int nbatch = (orows + BROWS - 1) / BROWS;

    for (int i = tid; i < NBUF * BATCH_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(xb)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < W_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(ws)[i] = __ldg(P.w_pack + i);
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(bar(B_BFULL + i), LOADER_WARPS); mbar_init(bar(B_BFREE + i), 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(bar(B_DFULL + i), 1); mbar_init(bar(B_DFREE + i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    __half2 amax2 = __float2half2_rn(0.f);

    if (warp < LOADER_WARPS) {
        // ===================== input rows: global -> + bias_p, LeakyReLU -> fp16 -> plane layout =====================
        // thread = (pixel, 8-channel chunk): a warp reads 8 pixels x 128 B (contiguous) and stores 8 pixels x 4 planes
        // (wide 64-channel rows - 54 pixels x 8 chunks - take two passes per thread)
        constexpr int PASSES = (NPLANES * (PITCH - 2) + LOADER_WARPS * 32 - 1) / (LOADER_WARPS * 32), PXSTEP = LOADER_WARPS * 32 / NPLANES;
        const int px = tid / NPLANES, ch = tid % NPLANES;
        float bp[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bp[e] = P.bias_p[ch * 8 + e];
        const float sp = P.slope_p;
        // (point, batch) items of this CTA in order; item it = point blockIdx.x + (it / nbatch) gridDim.x, batch it % nbatch
        const int my_points = (P.N - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int items = my_points * nbatch;
        auto prefetch_item = [&](int it) {                 // the batch's input rows are contiguous in memory: one bulk L2 prefetch
            if (it >= items) return;
            const int n = blockIdx.x + (it / nbatch) * gridDim.x, bi = it % nbatch;
            const int r0 = max(bi * BROWS - 1, 0), r1 = min(bi * BROWS - 1 + XROWS, h);
            if (r1 > r0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                             ::"l"(P.x + ((size_t)n * h + r0) * w * (CIN / 4)), "r"((r1 - r0) * w * CIN * 4) : "memory");
        };
        constexpr int PFD = 3;                             // prefetch distance in batches
        if (tid == 0)
            for (int it = 0; it < PFD; ++it) prefetch_item(it);
        constexpr int DEPTH = PASSES == 1 ? 4 : 3;         // input rows in flight per thread (registers)
        for (int it = 0; it < items; ++it) {
            const int n = blockIdx.x + (it / nbatch) * gridDim.x, bi = it % nbatch;
            const float4* xn = P.x + ((size_t)n * h * w + px) * (CIN / 4) + ch * 2;
            const int buf = it % NBUF;
            const int i0 = bi * BROWS - 1;                 // first input row of the batch (row -1 = the padding)
            if (tid == 0) prefetch_item(it + PFD);
            uint8_t* dst = xb + buf * BATCH_BYTES + ch * PLANE + (1 + px) * 16;
            float4 v[DEPTH][2 * PASSES];
            auto load = [&](int r, float4 (&t)[2 * PASSES]) {
                const int row = i0 + r;
#pragma unroll
                for (int q = 0; q < PASSES; ++q) {
                    if (px + q * PXSTEP < w && (unsigned)row < (unsigned)h) {
                        const float4* g = xn + ((size_t)row * w + q * PXSTEP) * (CIN / 4);
                        t[2 * q] = __ldcs(g); t[2 * q + 1] = __ldcs(g + 1);
                    } else {
                        t[2 * q] = make_float4(0.f, 0.f, 0.f, 0.f); t[2 * q + 1] = t[2 * q];
                    }
                }
            };
            auto stage = [&](int r, const float4 (&t)[2 * PASSES]) {
#pragma unroll
                for (int q = 0; q < PASSES; ++q) {
                    const bool active = px + q * PXSTEP < w;
                    const bool ok = active && (unsigned)(i0 + r) < (unsigned)h;
                    const float a[8] = {t[2 * q].x, t[2 * q].y, t[2 * q].z, t[2 * q].w, t[2 * q + 1].x, t[2 * q + 1].y, t[2 * q + 1].z, t[2 * q + 1].w};
                    uint32_t hv[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        float v0 = a[e] + bp[e], v1 = a[e + 1] + bp[e + 1];
                        v0 = fmaxf(v0, v0 * sp); v1 = fmaxf(v1, v1 * sp);
                        const __half2 hh = __floats2half2_rn(v0, v1);
                        hv[e >> 1] = ok ? h2_bits(hh) : 0u;           // zero rows / columns: the convolution pads the ACTIVATED map
                        if (ok) amax2 = __hmax2_nan(amax2, __habs2(hh));
                    }
                    if (active) *reinterpret_cast<uint4*>(dst + (r * PITCH + q * PXSTEP) * 16) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                }
            };
            // the first rows are requested BEFORE the wait for the buffer (they only need registers)
#pragma unroll
            for (int r = 0; r < DEPTH - 1; ++r) load(r, v[r]);
            mbar_wait_sleep(bar(B_BFREE + buf), ((it / NBUF) & 1) ^ 1, 100);
#pragma unroll
            for (int r = 0; r < XROWS; ++r) {
                if (r + DEPTH - 1 < XROWS) load(r + DEPTH - 1, v[(r + DEPTH - 1) % DEPTH]);
                stage(r, v[r % DEPTH]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_BFULL + buf));
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        const uint32_t xa = smem_u32(xb), wa = smem_u32(ws);
        constexpr uint32_t IDESC = idesc_f16(NOUT);
        const uint64_t bd0 = smem_desc_nosw(wa, 128, 256);
        uint32_t nb = 0, np = 0;                           // batches / pairs issued so far
        for (int n = blockIdx.x; n < P.N; n += gridDim.x) {
            for (int bi = 0; bi < nbatch; ++bi, ++nb) {
                const int buf = nb % NBUF;
                mbar_wait(bar(B_BFULL + buf), (nb / NBUF) & 1);
                const uint64_t ad0 = smem_desc_nosw(xa + buf * BATCH_BYTES, PLANE, 128);
                const int npairs = min(BROWS / RPT, (orows - bi * BROWS + RPT - 1) / RPT);   // tiles of RPT y4 rows in this batch
                for (int k = 0; k < npairs; ++k, ++np) {
                    const int acc = np % NACC;
                    mbar_wait(bar(B_DFREE + acc), ((np / NACC) & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < TAPS; ++t) {
                            const int u = t / TAPW, v = t % TAPW;
#pragma unroll
                            for (int kh = 0; kh < KSTEPS; ++kh)
                                umma_f16(tmem + acc * NOUT,
                                         ad0 + (uint64_t)(kh * 2 * (PLANE / 16) + (RPT * k + u) * PITCH + v),
                                         bd0 + (uint64_t)((t * KSTEPS + kh) * (NOUT * 32 / 16)), IDESC, (t | kh) ? 1u : 0u);
                        }
                        umma_commit(bar(B_DFULL + acc));
                        if (k == npairs - 1) umma_commit(bar(B_BFREE + buf));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= W_EPI) {
        // ===================== finish: TMEM -> depth-to-space + bias_a + LeakyReLU -> fp16 -> staging -> bulk stores =====
        const int quad = warp & 3, grp = (warp - W_EPI) >> 2;
        const int row = quad * 32 + lane;                          // TMEM lane = position in the tile
        const int r2 = row / PITCH, J = row - r2 * PITCH;          // y4 row of the tile, y4 column
        const int leader = (W_EPI + 4 * grp) * 32;
        const uint32_t ROWB = (uint32_t)W * 32;                    // HALF_OUT: bytes of one output row (2 planes x W x 16 B)
        uint8_t* stage = stg + grp * STAGE_BYTES;
        const float slope = P.slope;
        uint32_t np = 0;
        for (int n = blockIdx.x; n < P.N; n += gridDim.x) {
            for (int bi = 0; bi < nbatch; ++bi) {
                const int npairs = min(BROWS / RPT, (orows - bi * BROWS + RPT - 1) / RPT);
                for (int k = 0; k < npairs; ++k, ++np) {
                    if ((int)(np % EPI_GROUPS) != grp) continue;
                    const int acc = np % NACC;
                    const int p = bi * BROWS + RPT * k;            // first y4 row of the tile
                    const int I = p + r2;
                    const bool lane_ok = row < RPT * PITCH && (K3 ? (I < h && J < w) : (I <= h && J <= w));
                    mbar_wait_sleep(bar(B_DFULL + acc), (np / NACC) & 1, 100);
                    tc_fence_after();
                    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + acc * NOUT;
                    if constexpr (C::HALF_OUT) {
                        uint32_t r0[32], r1[32];
                        TMEM_LD32(taddr, r0);
                        TMEM_LD32(taddr + 32, r1);
                        tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(B_DFREE + acc));
                        if (tid == leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the stores that read this buffer are done
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                        // phase (a, b) of y4 position (I, J) is output pixel (2 I - a, 2 J - b); local output row lr = Y - (2 p - 1)
                        if (lane_ok) {
#pragma unroll
                            for (int ab = 0; ab < 4; ++ab) {
                                const int a = ab >> 1, b = ab & 1;
                                const int Y = 2 * I - a, X = 2 * J - b;
                                if ((unsigned)Y < (unsigned)H && (unsigned)X < (unsigned)W) {
                                    const uint32_t* v = ab < 2 ? r0 + 16 * ab : r1 + 16 * (ab - 2);
                                    uint32_t hv[8];
#pragma unroll
                                    for (int c = 0; c < 16; c += 2) {
                                        float v0 = __uint_as_float(v[c]) + P.bias_a[c], v1 = __uint_as_float(v[c + 1]) + P.bias_a[c + 1];
                                        v0 = fmaxf(v0, v0 * slope); v1 = fmaxf(v1, v1 * slope);
                                        const __half2 hh = __floats2half2_rn(v0, v1);
                                        amax2 = __hmax2_nan(amax2, __habs2(hh));
                                        hv[c >> 1] = h2_bits(hh);
                                    }
                                    uint8_t* d = stage + (uint32_t)(2 * r2 - a + 1) * ROWB + (uint32_t)X * 16;
                                    *reinterpret_cast<uint4*>(d) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                                    *reinterpret_cast<uint4*>(d + (uint32_t)W * 16) = make_uint4(hv[4], hv[5], hv[6], hv[7]);
                                }
                            }
                        }
                        fence_async_smem();
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                        if (tid == leader) {
                            for (int lr = 0; lr < 2 * RPT; ++lr) {
                                const int Y = 2 * p - 1 + lr;
                                if ((unsigned)Y >= (unsigned)H) continue;
                                __half* gdst = P.yh + ((size_t)n * H + Y) * (size_t)W * 16;
                                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                             ::"l"(gdst), "r"(smem_u32(stage + lr * ROWB)), "r"(ROWB) : "memory");
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    } else {
                        // float32 channels-last out: phase ab = 32 accumulator columns = the COUT channels of output pixel
                        // (2 I - a, 2 J - b): + bias_a, LeakyReLU, eight 16-byte stores (the pixel's 128 contiguous bytes)
                        static_assert(C::HALF_OUT || C::COUT == 32, "float32 epilogue: one TMEM_LD32 per phase");
                        // The pixels of a warp's lanes are 512 B apart in memory; storing them lane by lane is 32 partial sectors per
                        // instruction.  Each lane parks its pixel in a per-warp staging row, then the warp writes 4 whole pixels per
                        // instruction (8 lanes x 16 B = one pixel's 128 contiguous bytes), addresses passed by shuffle.
                        uint8_t* wst = stage + (warp & 3) * (32 * 144);
                        constexpr int NPH = K3 ? 1 : 4;                    // phases per accumulator (3x3: the tile's pixels themselves)
#pragma unroll
                        for (int ab = 0; ab < NPH; ++ab) {
                            uint32_t r[32];
                            TMEM_LD32(taddr + 32 * ab, r);
                            tmem_wait_ld();
                            if (ab == NPH - 1) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bar(B_DFREE + acc));
                            }
                            const int a = ab >> 1, b = ab & 1;
                            const int Y = K3 ? I : 2 * I - a, X = K3 ? J : 2 * J - b;
                            const int OH = K3 ? h : H, OW = K3 ? w : W;
                            const bool ok = lane_ok && (unsigned)Y < (unsigned)OH && (unsigned)X < (unsigned)OW;
                            const unsigned long long gp = ok ? (unsigned long long)(P.out + (((size_t)n * OH + Y) * OW + X) * (size_t)P.ctot) : 0ull;
                            __syncwarp();                                  // the previous phase has been read out of the staging rows
#pragma unroll
                            for (int c = 0; c < 32; c += 4) {
                                float4 q;
                                q.x = __uint_as_float(r[c]) + P.bias_a[c];         q.y = __uint_as_float(r[c + 1]) + P.bias_a[c + 1];
                                q.z = __uint_as_float(r[c + 2]) + P.bias_a[c + 2]; q.w = __uint_as_float(r[c + 3]) + P.bias_a[c + 3];
                                q.x = fmaxf(q.x, q.x * slope); q.y = fmaxf(q.y, q.y * slope);
                                q.z = fmaxf(q.z, q.z * slope); q.w = fmaxf(q.w, q.w * slope);
                                *reinterpret_cast<float4*>(wst + lane * 144 + c * 4) = q;
                            }
                            __syncwarp();
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int src = it * 4 + (lane >> 3);
                                const unsigned long long g = __shfl_sync(0xffffffffu, gp, src);
                                const float4 q = *reinterpret_cast<const float4*>(wst + src * 144 + (lane & 7) * 16);
                                if (g) reinterpret_cast<float4*>(g)[lane & 7] = q;
                            }
                        }
                    }
                }
            }
        }
        if (C::HALF_OUT && tid == leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    {
        const float2 m = __half22float2(amax2);
        if (P.status && !(fmaxf(m.x, m.y) <= 65504.f)) atomicOr(P.status, 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace upconv
}  // namespace riders

template <class C>
static int upconv_launch(const riders::upconv::Params& P, cudaStream_t st, const char* what) {
    using namespace riders;
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(upconv::upconv_tc_kernel<C>, (size_t)(C::SMEM_BYTES), configured, "upconv_tc: shared memory attribute")) return rc_smem;
    const int grid = P.N < sm_count() ? P.N : sm_count();
    count_launches(1);
    upconv::upconv_tc_kernel<C><<<grid, upconv::THREADS, C::SMEM_BYTES, st>>>(P);
    return check_launch(what);
}

static int upconv_fill(riders::upconv::Params& P, const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                       int N, int h, int w, const float* bias_a_host, float slope, int cin, int cout, int pitch) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && h > 0 && w > 0, "upconv_tc: bad shape");
    RIDERS_REQUIRE(w + 2 <= pitch, "upconv_tc: at most %d columns (got w = %d)", pitch - 2, w);
    RIDERS_REQUIRE(slope >= 0.f && slope <= 1.f && slope_p >= 0.f && slope_p <= 1.f, "upconv_tc: LeakyReLU slopes must be in [0, 1]");
    P.x = reinterpret_cast<const float4*>(x);
    P.w_pack = reinterpret_cast<const uint4*>(w_pack);
    P.yh = nullptr; P.out = nullptr; P.ctot = 0; P.status = nullptr;
    for (int c = 0; c < 64; ++c) P.bias_p[c] = (bias_p_host && c < cin) ? bias_p_host[c] : 0.f;
    P.slope_p = slope_p;
    for (int c = 0; c < 32; ++c) P.bias_a[c] = (bias_a_host && c < cout) ? bias_a_host[c] : 0.f;
    P.slope = slope;
    P.N = N; P.h = h; P.w = w;
    return RIDERS_OK;
}

extern "C" int riders_upconv_tc_half_nhwc(const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                                          int N, int h, int w, const float* bias_a_host, float slope, void* yh,
                                          int* status, void* stream) {
    using namespace riders;
    upconv::Params P;
    if (int rc = upconv_fill(P, x, bias_p_host, slope_p, w_pack, N, h, w, bias_a_host, slope, 32, 16, 56)) return rc;
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && w_pack && yh, "upconv_tc: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)w_pack) | ((uintptr_t)yh)) & 15) == 0, "upconv_tc: 16-byte alignment");
    P.yh = reinterpret_cast<__half*>(yh);
    P.status = status;
    return upconv_launch<upconv::Cfg<32, 16, 56, true>>(P, (cudaStream_t)stream, "upconv_tc_half_nhwc");
}

extern "C" int riders_upconv_tc_f32_nhwc(const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                                         int N, int h, int w, const float* bias_a_host, float slope, float* out, int ctot,
                                         int* status, void* stream) {
    using namespace riders;
    upconv::Params P;
    if (int rc = upconv_fill(P, x, bias_p_host, slope_p, w_pack, N, h, w, bias_a_host, slope, 64, 32, 28)) return rc;
    RIDERS_REQUIRE(ctot >= 32 && ctot % 4 == 0, "upconv_tc_f32: the output pixel stride must be >= 32 channels and a multiple of 4 (got %d)", ctot);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && w_pack && out, "upconv_tc: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)w_pack) | ((uintptr_t)out)) & 15) == 0, "upconv_tc: 16-byte alignment");
    P.out = out;
    P.ctot = ctot;
    P.status = status;
    return upconv_launch<upconv::Cfg<64, 32, 28, false>>(P, (cudaStream_t)stream, "upconv_tc_f32_nhwc");
}

extern "C" int riders_conv3x3_tc_f32_nhwc(const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                                          int N, int h, int w, const float* bias_host, float slope, float* out, int ctot,
                                          int* status, void* stream) {
    using namespace riders;
    upconv::Params P;
    if (int rc = upconv_fill(P, x, bias_p_host, slope_p, w_pack, N, h, w, bias_host, slope, 64, 32, 56)) return rc;
    RIDERS_REQUIRE(ctot >= 32 && ctot % 4 == 0, "conv3x3_tc_f32: the output pixel stride must be >= 32 channels and a multiple of 4 (got %d)", ctot);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && w_pack && out, "conv3x3_tc: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)w_pack) | ((uintptr_t)out)) & 15) == 0, "conv3x3_tc: 16-byte alignment");
    P.out = out;
    P.ctot = ctot;
    P.status = status;
    return upconv_launch<upconv::Cfg<64, 32, 56, false, true>>(P, (cudaStream_t)stream, "conv3x3_tc_f32_nhwc");
}
