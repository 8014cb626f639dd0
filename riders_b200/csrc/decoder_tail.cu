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
// CTA = 12 x 50 output pixels of one point (6 warps, 2 CTAs / SM):
//   1. y tile (16 x 54, halo 2) <- y4 + bias_a, LeakyReLU, cvt.rna.tf32; 20 floats per pixel so that the 8 rows of an
//      ldmatrix 8 x 16 B tile fall into distinct bank groups;
//   2. z tile (14 x 52, halo 1) as 46 M-tiles of 16 consecutive positions: per tile 18 ldmatrix.x4 (tap-shifted
//      windows of the y tile: implicit GEMM, no im2col) and 36 mma.m16n8k8.tf32;
//   3. in registers: bias_b, LeakyReLU, zero outside the patch (the head's zero padding), then the head's nine
//      per-tap dot products over the 16 channels (4 per lane + quad butterfly) -> u[tap][position] in shared memory;
//   4. logits[p] = sum_tap u[tap][p + tap offset]; optional sigmoid; coalesced store.
#include "common.cuh"

namespace riders {
namespace tail {

constexpr int C = 16;
constexpr int TH = 12, TW = 50;
constexpr int YH = TH + 4, YW = TW + 4;
constexpr int ZH = TH + 2, ZW = TW + 2;
constexpr int YP = 20;                               // floats per y-tile pixel
constexpr int NZ = ZH * ZW;                          // 728 z positions
constexpr int MT = (NZ + 15) / 16;                   // 46 M-tiles
constexpr int UP = MT * 16 + 8;                      // u plane stride: == 8 mod 32 -> the quad's 4 taps hit distinct banks
constexpr int WARPS = 6, THREADS = WARPS * 32;
constexpr int KSTEPS = 18;                           // 9 taps x 2 halves of the 16 input channels
constexpr int SMEM_BYTES = (YH * YW * YP + 9 * UP) * 4;      // 69120 + 26784 = 95904

struct Params {
    float bias_a[C];
    float bias_b[C];
    float wh[9][C];          // [tap = dy * 3 + dx][c]
    float slope;
    int sigmoid;
    int t;                   // 0: y4 from the 3x3 phase kernels (h, w); 1: from the 2x2 phase kernels, (h + 1, w + 1)
};

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&a)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(THREADS, 2)
decoder_tail_kernel(const float4* __restrict__ y4, const float2* __restrict__ wb_frag, float* __restrict__ out,
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

    // Wb fragments (constant over the whole kernel): b0 = B[k = t][n = g], b1 = B[k = t + 4][n = g] per (K step, N half)
    uint32_t bf[KSTEPS][2][2];
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const float2 v = __ldg(wb_frag + (s * 2 + nt) * 32 + lane);
            bf[s][nt][0] = to_tf32(v.x);
            bf[s][nt][1] = to_tf32(v.y);
        }

    // ---- 1. y tile: depth-to-space + bias + LeakyReLU + TF32 rounding, zero outside the patch --------------------
    {
        const int hq = h + P.t, wq = w + P.t;
        const float4* src = y4 + (size_t)n * hq * wq * (4 * C / 4);
#pragma unroll 6
        for (int idx = tid; idx < YH * YW * 4; idx += THREADS) {
            const int pix = idx >> 2, q = idx & 3;
            const int yy = pix / YW, yx = pix - yy * YW;
            const int Y = R0 - 2 + yy, X = C0 - 2 + yx;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((unsigned)Y < (unsigned)H && (unsigned)X < (unsigned)W) {
                const int a = Y & 1, b = X & 1;
                v = __ldcs(src + ((size_t)((Y >> 1) + a * P.t) * wq + (X >> 1) + b * P.t) * (4 * C / 4) + (2 * a + b) * (C / 4) + q);
                v.x += P.bias_a[4 * q + 0]; v.y += P.bias_a[4 * q + 1]; v.z += P.bias_a[4 * q + 2]; v.w += P.bias_a[4 * q + 3];
                v.x = v.x > 0.f ? v.x : v.x * P.slope; v.y = v.y > 0.f ? v.y : v.y * P.slope;
                v.z = v.z > 0.f ? v.z : v.z * P.slope; v.w = v.w > 0.f ? v.w : v.w * P.slope;
                v.x = __uint_as_float(to_tf32(v.x)); v.y = __uint_as_float(to_tf32(v.y));
                v.z = __uint_as_float(to_tf32(v.z)); v.w = __uint_as_float(to_tf32(v.w));
            }
            *reinterpret_cast<float4*>(ys + pix * YP + 4 * q) = v;
        }
    }
    __syncthreads();

    // ---- 2 + 3. z = conv3x3(y) on the tensor cores, then bias / activation / head partial products in registers ----
    const uint32_t ys_addr = (uint32_t)__cvta_generic_to_shared(ys);
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;     // ldmatrix: this lane supplies row lrow of the 16-row tile ...
    const int lk = (lane >> 4) * 4;                           // ... at channel offset 0 (matrices 0, 1) or 4 (matrices 2, 3)
    float whr[9][4];                                          // the head weights of this lane's four channels
    float bbr[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int co = (e >> 1) * 8 + 2 * t + (e & 1);
        bbr[e] = P.bias_b[co];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) whr[tap][e] = P.wh[tap][co];
    }
    for (int m = warp; m < MT; m += WARPS) {
        const int l = min(16 * m + lrow, NZ - 1);
        const int zy = l / ZW, zx = l - zy * ZW;
        const uint32_t base = ys_addr + (uint32_t)(((zy * YW + zx) * YP + lk) * 4);
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t a[4];
                ldmatrix_x4(base + (uint32_t)(((dy * YW + dx) * YP + 8 * j) * 4), a);
                mma_tf32(acc[0], a, bf[tap * 2 + j][0][0], bf[tap * 2 + j][0][1]);
                mma_tf32(acc[1], a, bf[tap * 2 + j][1][0], bf[tap * 2 + j][1][1]);
            }
        }
        // accumulator layout: acc[nt][0..1] -> position 16 m + g, channels nt * 8 + 2 t (+1); acc[nt][2..3] -> position + 8
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int pos = 16 * m + g + 8 * half;
            const int py = pos / ZW, px = pos - py * ZW;
            const bool inside = pos < NZ && (unsigned)(R0 - 1 + py) < (unsigned)H && (unsigned)(C0 - 1 + px) < (unsigned)W;
            float z[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = acc[e >> 1][2 * half + (e & 1)] + bbr[e];
                v = v > 0.f ? v : v * P.slope;
                z[e] = inside ? v : 0.f;                      // zero padding of the ACTIVATED map for the head
            }
            float u[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                float s = whr[tap][0] * z[0];
                s = fmaf(whr[tap][1], z[1], s); s = fmaf(whr[tap][2], z[2], s); s = fmaf(whr[tap][3], z[3], s);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                u[tap] = s;
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
                if ((tap & 3) == t) us[tap * UP + pos] = u[tap];          // pos < MT * 16 <= UP
        }
    }
    __syncthreads();

    // ---- 4. logits = sum over the head's taps of the per-tap products of the neighbouring z positions -----------
    float* o = out + (size_t)n * H * W;
    for (int idx = tid; idx < TH * TW; idx += THREADS) {
        const int oy = idx / TW, ox = idx - oy * TW;
        if (R0 + oy < H && C0 + ox < W) {
            float s = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap - dy * 3;
                s += us[tap * UP + (oy + dy) * ZW + ox + dx];
            }
            __stcs(o + (size_t)(R0 + oy) * W + C0 + ox, P.sigmoid ? __fdividef(1.0f, 1.0f + __expf(-s)) : s);
        }
    }
}

// issue-rate probe for the warp-level TF32 MMA (riders_debug_mma_sync_rate): `warps` warps per CTA, one CTA per SM,
// 8 independent accumulators per warp, registers only
__global__ void __launch_bounds__(1024, 1)
mma_sync_rate_kernel(int iters, long long* cycles, float* sink) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u};
    const uint32_t b0 = 0x3f800000u, b1 = 0x3f000000u + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_tf32(acc[i], a, b0, b1);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
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
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tail::decoder_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tail::SMEM_BYTES) != cudaSuccess)
            return check_launch("decoder_tail: shared memory attribute");
        configured = true;
    }
    count_launches(1);
    tail::decoder_tail_kernel<<<(unsigned)blocks, tail::THREADS, tail::SMEM_BYTES, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(y4), reinterpret_cast<const float2*>(wb_frag), out, P, h, w, tiles_x, tiles_y);
    return check_launch("decoder_tail_nhwc");
}

extern "C" int riders_debug_mma_sync_rate(int warps, int iters, int ctas, long long* device_cycles, float* device_sink, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(warps >= 1 && warps <= 32 && iters > 0 && ctas > 0 && device_cycles && device_sink, "debug_mma_sync_rate: bad arguments");
    count_launches(1);
    tail::mma_sync_rate_kernel<<<ctas, warps * 32, 0, (cudaStream_t)stream>>>(iters, device_cycles, device_sink);
    return check_launch("debug_mma_sync_rate");
}
