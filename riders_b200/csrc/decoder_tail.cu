// Fused tail of RC-Net's decoder, the three passes in front of riders_confidence_reduce
// (MultiScaleDecoder, RCNet/networks.py:678-779; DecoderBlock, utils/net_utils.py:474-569):
//
//     y      = LeakyReLU(depth_to_space(y4) + bias_a)          last block's up-convolution (sub-pixel form, see upsample.cu)
//     z      = LeakyReLU(conv3x3(y; Wb) + bias_b)              last block's 16 -> 16 convolution
//     logits = conv3x3(z; Wh)        [-> sigmoid]              output0, 16 -> 1, no bias
//
// As separate passes (d2s kernel, cuDNN convolution, head kernel) the 16-channel 240 x 100 maps cross HBM five
// times per radar point (15.7 GB per 2048 points); here y4 is read once (64 B per output pixel) and 4 B per pixel are
// written: y and z only ever exist in shared memory / registers.
//
// Numerics: this is the TF32 path.  It is only selected where the reference flow would run the convolution in TF32
// as well (torch.backends.cudnn.allow_tf32, torch's default): y and Wb are rounded to TF32 with cvt.rna (what cuDNN /
// CUTLASS do), products accumulate in fp32; the 16 -> 1 head is plain fp32 FMA.  With allow_tf32 = False the host
// keeps the fp32 passes.
//
// Why mma.sync and not tcgen05 here: N = 16 output channels.  An SS-mode tcgen05.mma pays for re-reading the whole
// 128-row A operand from shared memory per 8-deep K step (~60 cycles for 128 x 16 x 8 MACs measured on this part,
// tools/mma_rate.py: ~43 + N/2), i.e. ~270 MAC/cycle/SM; the warp-level m16n8k8 keeps Wb in registers (72 per thread)
// and re-uses every A fragment for both 8-wide halves of N, so shared memory supplies 256 B per 2 x 1024 MACs.
//
// Measured (B200, 2048 points, profiles/r01_ncu_summary_v6_decoder_tail.txt): 2.9 ms against 3.25 ms for the three passes;
// tensor pipe 38 % busy, L1/shared data pipe 57 %, issue slots 34 %: with only 16 output channels there are ~5 scalar
// instructions (activation, rounding, masks, addressing) per MMA and the kernel is bound by their issue / latency, not by
// a pipe.  A warp-specialised persistent variant (loader + mma warps, double-buffered tiles, vertically shared A
// fragments) and deeper ldmatrix prefetch were built and measured at the same 2.9 - 3.0 ms (git history), so the simple
// form stays.
//
// CTA = 12 x 50 output pixels of one point (6 warps, 2 CTAs / SM):
//   1. y tile (16 x 54, halo 2) <- y4 + bias_a, LeakyReLU, cvt.rna.tf32; 20 floats per pixel so that the 8 rows of an
//      ldmatrix 8 x 16 B tile fall into distinct bank groups;
//   2. z tile (14 x 52, halo 1) as 46 M-tiles of 16 consecutive positions: per tile 18 ldmatrix.x4 (tap-shifted
//      windows of the y tile: implicit GEMM, no im2col) and 36 mma.m16n8k8.tf32;
//   3. in registers: bias_b, LeakyReLU, zero outside the patch (the head's zero padding), then the head's nine
//      per-tap dot products over the 16 channels (4 per lane + quad butterfly) -> u[tap][position] in shared memory;
//   4. logits[p] = sum_tap u[tap][p + tap offset]; optional sigmoid; coalesced store.
#include "common.cuh"
#include "mma_sync.cuh"

namespace riders {
namespace tail {

constexpr int C = 16;
constexpr int TH = 12, TW = 50;
constexpr int YH = TH + 4, YW = TW + 4;
constexpr int ZH = TH + 2, ZW = TW + 2;
constexpr int YP = 20;                               // floats per y-tile pixel
constexpr int NZ = ZH * ZW;                          // 728 z positions
constexpr int MT = (NZ + 15) / 16;                   // 46 M-tiles
constexpr int UP = MT * 16 + 4;                      // u plane stride: == 4 mod 32 -> taps 2t (+1) of the 4 quad lanes hit distinct banks
constexpr int WARPS = 6, THREADS = WARPS * 32;
constexpr int KSTEPS = 18;                           // 9 taps x 2 halves of the 16 input channels
constexpr int SMEM_BYTES = (YH * YW * YP + 9 * UP) * 4;      // 69120 + 26640 = 95760

struct Params {
    float bias_a[C];
    float bias_b[C];
    float wh[9][C];          // [tap = dy * 3 + dx][c]
    float slope;
    int sigmoid;
    int t;                   // 0: y4 from the 3x3 phase kernels (h, w); 1: from the 2x2 phase kernels, (h + 1, w + 1)
};

__global__ void __launch_bounds__(THREADS, 2)
decoder_tail_kernel(const float4* __restrict__ y4, const float4* __restrict__ wb_frag, float* __restrict__ out,
                    const __grid_constant__ Params P, int h, int w, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float smem[];
    float* ys = smem;                                 // [YH][YW][YP]
    float* us = smem + YH * YW * YP;                  // [9][UP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int H = 2 * h, W = 2 * w;
    int tile = blockIdx.x;
    const int tx = tile % tiles_x; tile /= tiles_x;
    const int ty = tile % tiles_y;
    const long long n = tile / tiles_y;
    const int R0 = ty * TH, C0 = tx * TW;
    const float slope = P.slope;

    // ---- 1. y tile: depth-to-space + bias_a + LeakyReLU + TF32 rounding, zero outside the patch (the convolution's zero
    //         padding applies to the ACTIVATED map).  Warp w owns tile rows w, w + 6, w + 12; a pass covers 8 pixels x 4
    //         quarters of a row, lane = (quarter, pixel): the 8 lanes of a 128-bit store phase hit 8 pixels (80 B apart:
    //         distinct bank groups).  Row part of the address once per row, column part once per thread (C0 and the 8-pixel
    //         pass are even, so the column phase is fixed per lane and a pass advances 4 low-resolution pixels): every
    //         load / store is base + immediate.  All loads are issued before the first store.
    {
        constexpr int PASSES = (YW + 7) / 8;                          // 7 (the last one covers 6 pixels)
        constexpr int ROWS_PER_WARP = (YH + WARPS - 1) / WARPS;       // 3
        static_assert((TW & 1) == 0, "column phase per lane needs an even tile origin");
        const int q = lane >> 3, l8 = lane & 7;
        const float4 ba = make_float4(P.bias_a[4 * q + 0], P.bias_a[4 * q + 1], P.bias_a[4 * q + 2], P.bias_a[4 * q + 3]);
        const int hq = h + P.t, wq = w + P.t, pt = P.t;
        const int X0 = C0 - 2 + l8;                                   // column of pass 0; X0 + 8 p for pass p
        const int bcol = X0 & 1;
        // float4 offset of (Y, X) = [((Y >> 1) + a pt) wq + (X >> 1) + b pt] * 16 + (2 a + b) * 4 + q
        const float4* src = y4 + (size_t)n * hq * wq * (4 * C / 4) + q + ((X0 >> 1) + bcol * pt) * (4 * C / 4) + bcol * (C / 4);
        float4 v[ROWS_PER_WARP][PASSES];
        bool in[ROWS_PER_WARP][PASSES];
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
            const int Y = R0 - 2 + warp + r * WARPS, a = Y & 1;
            const bool row_ok = warp + r * WARPS < YH && (unsigned)Y < (unsigned)H;
            const float4* rp = src + (row_ok ? (((Y >> 1) + a * pt) * wq) * (4 * C / 4) + a * (2 * C / 4) : 0);
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                in[r][p] = row_ok && (unsigned)(X0 + 8 * p) < (unsigned)W && l8 + 8 * p < YW;
                v[r][p] = __ldcs(in[r][p] ? rp + p * 4 * (4 * C / 4) : y4);       // always a valid address: no branch
            }
        }
        float* dstp = ys + ((warp * YW + l8) * YP + 4 * q);
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                float4 x = v[r][p];
                x.x += ba.x; x.y += ba.y; x.z += ba.z; x.w += ba.w;
                x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);          // LeakyReLU, 0 <= slope <= 1
                x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
                uint4 o;
                o.x = in[r][p] ? to_tf32(x.x) : 0u; o.y = in[r][p] ? to_tf32(x.y) : 0u;
                o.z = in[r][p] ? to_tf32(x.z) : 0u; o.w = in[r][p] ? to_tf32(x.w) : 0u;
                if (warp + r * WARPS < YH && l8 + 8 * p < YW)
                    *reinterpret_cast<uint4*>(dstp + (r * WARPS * YW + 8 * p) * YP) = o;
            }
        }
    }

    // Wb fragments (constant over the whole kernel, TF32-rounded by the host): per K step s one float4 per lane =
    // {b0, b1 of N half 0, b0, b1 of N half 1}, b0 = B[k = t][n = g], b1 = B[k = t + 4][n = g]
    uint32_t bf[KSTEPS][2][2];
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
        const float4 v = __ldg(wb_frag + s * 32 + lane);
        bf[s][0][0] = __float_as_uint(v.x); bf[s][0][1] = __float_as_uint(v.y);
        bf[s][1][0] = __float_as_uint(v.z); bf[s][1][1] = __float_as_uint(v.w);
    }
    // head: u[pos][tap] = sum_c z[pos][c] Wh[tap][c] as a second MMA whose A operand IS the accumulator fragment of the
    // first one (K slot t <-> channel 2t, slot t + 4 <-> channel 2t + 1 of each 8-channel half: the K order is free as
    // long as B uses the same one), N = taps 0..7 | 8 (+ 7 zero columns)
    uint32_t hb[2][2][2];                                     // [channel half][tap half][b0, b1]
    float bbr[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            bbr[nt][e] = P.bias_b[nt * 8 + 2 * t + e];
            hb[nt][0][e] = to_tf32(P.wh[g][nt * 8 + 2 * t + e]);
            hb[nt][1][e] = g == 0 ? to_tf32(P.wh[8][nt * 8 + 2 * t + e]) : 0u;
        }
    }
    __syncthreads();

    // ---- 2 + 3. z = conv3x3(y) on the tensor cores; bias_b, LeakyReLU, zero outside the patch (the head's zero padding),
    //             TF32; head per-tap products by MMA straight from the accumulator registers -> u[tap][position] ----
    const uint32_t ys_addr = (uint32_t)__cvta_generic_to_shared(ys);
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;     // ldmatrix: this lane supplies row lrow of the 16-row tile ...
    const int lk = (lane >> 4) * 4;                           // ... at channel offset 0 (matrices 0, 1) or 4 (matrices 2, 3)
    // two M-tiles per iteration: four independent accumulator chains per warp keep the tensor pipe fed
    for (int mp = warp; mp < (MT + 1) / 2; mp += WARPS) {
        uint32_t base[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int l = min(16 * (2 * mp + i) + lrow, NZ - 1);
            const int zy = l / ZW, zx = l - zy * ZW;
            base[i] = ys_addr + (uint32_t)(((zy * YW + zx) * YP + lk) * 4);
        }
        float acc[2][2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][nt][e] = bbr[nt][e & 1];       // bias_b rides in the accumulator
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t a0[4], a1[4];
                ldmatrix_x4(base[0] + (uint32_t)(((dy * YW + dx) * YP + 8 * j) * 4), a0);
                ldmatrix_x4(base[1] + (uint32_t)(((dy * YW + dx) * YP + 8 * j) * 4), a1);
                mma_tf32(acc[0][0], a0, bf[tap * 2 + j][0][0], bf[tap * 2 + j][0][1]);
                mma_tf32(acc[1][0], a1, bf[tap * 2 + j][0][0], bf[tap * 2 + j][0][1]);
                mma_tf32(acc[0][1], a0, bf[tap * 2 + j][1][0], bf[tap * 2 + j][1][1]);
                mma_tf32(acc[1][1], a1, bf[tap * 2 + j][1][0], bf[tap * 2 + j][1][1]);
            }
        }
        // accumulator layout: acc[.][nt][0..1] -> position 16 m + g, channels nt * 8 + 2 t (+1); acc[.][nt][2..3] -> position + 8
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int pos0 = 16 * (2 * mp + i) + g;
            bool inside[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int pos = pos0 + 8 * half;
                const int py = pos / ZW, px = pos - py * ZW;
                inside[half] = pos < NZ && (unsigned)(R0 - 1 + py) < (unsigned)H && (unsigned)(C0 - 1 + px) < (unsigned)W;
            }
            float u[2][4];
#pragma unroll
            for (int nh = 0; nh < 2; ++nh)
#pragma unroll
                for (int e = 0; e < 4; ++e) u[nh][e] = 0.f;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                uint32_t za[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float v = fmaxf(acc[i][nt][e], acc[i][nt][e] * slope);
                    za[e] = inside[e >> 1] ? to_tf32(v) : 0u;
                }
                const uint32_t afrag[4] = {za[0], za[2], za[1], za[3]};      // (g, k=t) (g+8, k=t) (g, k=t+4) (g+8, k=t+4)
                mma_tf32(u[0], afrag, hb[nt][0][0], hb[nt][0][1]);
                mma_tf32(u[1], afrag, hb[nt][1][0], hb[nt][1][1]);
            }
            // u[0][0..1]: position pos0, taps 2t, 2t+1; u[0][2..3]: position pos0 + 8; u[1][0], u[1][2]: tap 8 on the t == 0 lanes
            us[(2 * t) * UP + pos0] = u[0][0];
            us[(2 * t + 1) * UP + pos0] = u[0][1];
            us[(2 * t) * UP + pos0 + 8] = u[0][2];
            us[(2 * t + 1) * UP + pos0 + 8] = u[0][3];
            if (t == 0) { us[8 * UP + pos0] = u[1][0]; us[8 * UP + pos0 + 8] = u[1][2]; }
        }
    }
    __syncthreads();

    // ---- 4. logits = sum over the head's taps of the per-tap products of the neighbouring z positions -----------
    float* o = out + (size_t)n * H * W;
    {
        int oy = tid / TW, ox = tid - oy * TW;
#pragma unroll
        for (int k = 0; k < (TH * TW + THREADS - 1) / THREADS; ++k) {
            if (oy < TH && R0 + oy < H && C0 + ox < W) {
                const float* up = us + oy * ZW + ox;
                float s = 0.f;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) s += up[tap * UP + (tap / 3) * ZW + (tap % 3)];
                __stcs(o + (size_t)(R0 + oy) * W + C0 + ox, P.sigmoid ? (1.0f / (1.0f + expf(-s))) : s);
            }
            ox += THREADS % TW; oy += THREADS / TW;
            if (ox >= TW) { ox -= TW; ++oy; }
        }
    }
}

// issue-rate probe for the warp-level TF32 MMA (riders_debug_mma_sync_rate): `warps` warps per CTA, one CTA per SM,
// 8 independent accumulators per warp, registers only
// MODE 0: all MMAs share one A and one B fragment (operand registers re-used back to back); MODE 1: every MMA of an
// iteration has its own A and B registers, as in a real kernel.  CHAINS independent accumulators per warp.
template <int CHAINS, int MODE>
__global__ void __launch_bounds__(1024, 1)
mma_sync_rate_kernel(int iters, long long* cycles, float* sink) {
    float acc[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    uint32_t a[8][4], b[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) a[i][e] = 0x3f800000u + threadIdx.x * 8 + (MODE ? i * 4 + e : e) + iters;
        b[i][0] = 0x3f000000u + threadIdx.x + (MODE ? i : 0) + iters;
        b[i][1] = 0x3f400000u + threadIdx.x * 3 + (MODE ? i : 0) + iters;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8 / CHAINS; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                const int o = MODE ? r * CHAINS + i : 0;
                mma_tf32(acc[i], a[o], b[o][0], b[o][1]);
            }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace tail

}  // namespace riders

extern "C" int riders_decoder_tail_nhwc(const float* y4, int taps, int N, int h, int w,
                                        const float* bias_a_host, const float* wb_frag, const float* bias_b_host,
                                        const float* wh_host, float slope, int apply_sigmoid,
                                        float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && h > 0 && w > 0, "decoder_tail: bad shape");
    RIDERS_REQUIRE(taps == 2 || taps == 3, "decoder_tail: taps must be 2 or 3 (got %d)", taps);
    RIDERS_REQUIRE(slope >= 0.f && slope <= 1.f, "decoder_tail: LeakyReLU slope must be in [0, 1] (got %g)", (double)slope);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(y4 && wb_frag && wh_host && out, "decoder_tail: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)y4) | ((uintptr_t)wb_frag)) & 15) == 0, "decoder_tail: y4 / wb_frag must be 16-byte aligned");
    tail::Params P;
    for (int c = 0; c < tail::C; ++c) {
        P.bias_a[c] = bias_a_host ? bias_a_host[c] : 0.f;
        P.bias_b[c] = bias_b_host ? bias_b_host[c] : 0.f;
        for (int tap = 0; tap < 9; ++tap) P.wh[tap][c] = wh_host[c * 9 + tap];        // torch (1, C, 3, 3)
    }
    P.slope = slope;
    P.sigmoid = apply_sigmoid ? 1 : 0;
    P.t = taps == 2 ? 1 : 0;
    const int tiles_x = ceil_div(2 * w, tail::TW), tiles_y = ceil_div(2 * h, tail::TH);
    const long long blocks = (long long)N * tiles_x * tiles_y;
    RIDERS_REQUIRE(blocks < (1ll << 31), "decoder_tail: too many tiles for one launch");
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(tail::decoder_tail_kernel, (size_t)(tail::SMEM_BYTES), configured, "decoder_tail: shared memory attribute")) return rc_smem;
    count_launches(1);
    tail::decoder_tail_kernel<<<(unsigned)blocks, tail::THREADS, tail::SMEM_BYTES, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(y4), reinterpret_cast<const float4*>(wb_frag), out, P, h, w, tiles_x, tiles_y);
    return check_launch("decoder_tail_nhwc");
}

extern "C" int riders_debug_mma_sync_rate(int warps, int iters, int ctas, long long* device_cycles, float* device_sink, void* stream) {
    using namespace riders;
    // warps: bits 0..7 warps per CTA; bits 8..11 independent accumulators per warp (1, 2, 4, 8; 0 = 8); bit 12: distinct
    // operand registers per MMA
    const int chains = (warps >> 8) & 15 ? (warps >> 8) & 15 : 8, mode = (warps >> 12) & 1;
    warps &= 255;
    RIDERS_REQUIRE(warps >= 1 && warps <= 32 && iters > 0 && ctas > 0 && device_cycles && device_sink, "debug_mma_sync_rate: bad arguments");
    RIDERS_REQUIRE(chains == 1 || chains == 2 || chains == 4 || chains == 8, "debug_mma_sync_rate: chains must be 1, 2, 4 or 8");
    count_launches(1);
    cudaStream_t st = (cudaStream_t)stream;
#define RIDERS_RATE(C_, M_) tail::mma_sync_rate_kernel<C_, M_><<<ctas, warps * 32, 0, st>>>(iters, device_cycles, device_sink)
    if (mode) { if (chains == 1) RIDERS_RATE(1, 1); else if (chains == 2) RIDERS_RATE(2, 1); else if (chains == 4) RIDERS_RATE(4, 1); else RIDERS_RATE(8, 1); }
    else      { if (chains == 1) RIDERS_RATE(1, 0); else if (chains == 2) RIDERS_RATE(2, 0); else if (chains == 4) RIDERS_RATE(4, 0); else RIDERS_RATE(8, 0); }
#undef RIDERS_RATE
    return check_launch("debug_mma_sync_rate");
}
