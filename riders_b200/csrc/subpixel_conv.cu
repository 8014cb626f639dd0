// Sub-pixel phase convolution of the LAST decoder block (32 -> 4 x 16 channels, 120 x 50 -> 121 x 51 positions; the
// 2x2-tap form of "nearest 2x resize -> 3x3 convolution", see upsample.cu) with the bias + LeakyReLU of the convolution
// that produced its input fused into the operand load (MultiScaleDecoder deconv1.conv -> deconv0.deconv,
// RCNet/networks.py:678-779, utils/net_utils.py:474-569).  It feeds riders_decoder_tail_nhwc.
//
// In the library flow this is two passes: an in-place bias + activation pass over the 32-channel 120 x 50 maps and cuDNN's
// generic implicit GEMM for the 2x2 kernel (1.7 ms per 2048 points, 118 TFLOP/s effective: small K, N = 64).  Here one
// kernel reads the raw 32-channel map once and writes the 64-channel phase map once.  TF32 flow only (the host selects it
// where torch.backends.cudnn.allow_tf32, like the tail kernel): operands rounded with cvt.rna, fp32 accumulation.
//
// GEMM view: M = output positions (p, q) of the (h + 1) x (w + 1) grid, N = 64 = 4 phases x 16 channels,
// K = 128 = 2 x 2 taps x 32 channels:  y4[p][q][n] = sum_{u, v, c} W2[n][c][u][v] * act(x + bias)[p - 1 + u][q - 1 + v][c]
// with zeros outside the map.  CTA = 8 position rows x (w + 1) columns of one point:
//   1. x window (9 rows x (w + 2) columns) -> + bias, LeakyReLU, TF32 -> shared memory, 36 floats per pixel (ldmatrix rows
//      of 16 B at a 144 B stride fall into distinct bank groups); the 32 KB of B fragments next to it;
//   2. M-tiles of 16 consecutive positions, two per warp iteration: per 8-deep K step 2 ldmatrix.x4 (tap-shifted windows:
//      implicit GEMM) + 4 LDS.128 (B fragments of the 8 N tiles) feed 16 mma.m16n8k8.tf32;
//   3. accumulators -> y4 (float2 per lane and N tile; a quad writes one 32-byte sector).
#include "common.cuh"
#include "mma_sync.cuh"

namespace riders {
namespace spc {

using tail::to_tf32;
using tail::ldmatrix_x4;
using tail::mma_tf32;

constexpr int CIN = 32, NOUT = 64;
constexpr int TP = 8;                                  // position rows per CTA
constexpr int XP = 36;                                 // floats per staged pixel (32 + 4 pad)
constexpr int KSTEPS = 16, NTILES = 8;
constexpr int WARPS = 8, THREADS = WARPS * 32;
constexpr int BS_FLOATS = KSTEPS * (NTILES / 2) * 32 * 4;          // [K step][N tile pair][lane] float4 = 32 KB

struct Params {
    float bias[CIN];
    float slope;
};

__global__ void __launch_bounds__(THREADS, 2)
subpixel_conv_kernel(const float4* __restrict__ x, const float4* __restrict__ w_frag, float* __restrict__ y4,
                     const __grid_constant__ Params P, int h, int w, int row_tiles) {
    extern __shared__ __align__(16) float smem[];
    const int WQ = w + 1, XW = w + 2;                  // positions per row, staged pixels per row
    float* bs = smem;                                  // B fragments
    float* xs = smem + BS_FLOATS;                      // [(TP + 1)][XW][XP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rt = blockIdx.x % row_tiles;
    const long long n = blockIdx.x / row_tiles;
    const int P0 = rt * TP;                            // first position row of this CTA
    const int rows = min(TP, h + 1 - P0);              // position rows actually present
    const float slope = P.slope;

    // ---- 1. stage the activated x window and the B fragments --------------------------------------------------------
    {
        const int q = lane & 7, l4 = lane >> 3;         // lane = (pixel of a group of 4, 16-byte quarter of its 32 channels)
        const float4 bq = make_float4(P.bias[4 * q + 0], P.bias[4 * q + 1], P.bias[4 * q + 2], P.bias[4 * q + 3]);
        const int passes = (XW + 3) / 4;
        const int units = (rows + 1) * passes;          // (staged row, group of 4 pixels)
        const float4* src = x + (size_t)n * h * w * (CIN / 4) + q;
        for (int u0 = warp; u0 < units; u0 += 4 * WARPS) {
            float4 v[4];
            int dst[4];
            bool in[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {               // four units in flight per warp
                const int u = u0 + i * WARPS;
                const int ry = u / passes, rx = (u - ry * passes) * 4 + l4;
                const int Y = P0 - 1 + ry, X = rx - 1;
                const bool ok = u < units && rx < XW;
                in[i] = ok && (unsigned)Y < (unsigned)h && (unsigned)X < (unsigned)w;
                dst[i] = ok ? (ry * XW + rx) * XP + 4 * q : -1;
                v[i] = __ldcs(in[i] ? src + ((size_t)Y * w + X) * (CIN / 4) : x);       // always a valid address
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (dst[i] < 0) continue;
                float4 r = v[i];
                r.x += bq.x; r.y += bq.y; r.z += bq.z; r.w += bq.w;
                r.x = fmaxf(r.x, r.x * slope); r.y = fmaxf(r.y, r.y * slope);            // LeakyReLU, 0 <= slope <= 1
                r.z = fmaxf(r.z, r.z * slope); r.w = fmaxf(r.w, r.w * slope);
                uint4 o;
                o.x = in[i] ? to_tf32(r.x) : 0u; o.y = in[i] ? to_tf32(r.y) : 0u;
                o.z = in[i] ? to_tf32(r.z) : 0u; o.w = in[i] ? to_tf32(r.w) : 0u;
                *reinterpret_cast<uint4*>(xs + dst[i]) = o;
            }
        }
        for (int i = tid; i < BS_FLOATS / 4; i += THREADS) reinterpret_cast<float4*>(bs)[i] = __ldg(w_frag + i);
    }
    __syncthreads();

    // ---- 2 + 3. implicit GEMM over tap-shifted windows, two 16-position M-tiles per warp iteration -------------------
    const int npos = rows * WQ;
    const int mtiles = (npos + 15) / 16;
    const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(xs);
    const uint32_t bs_addr = (uint32_t)__cvta_generic_to_shared(bs) + lane * 16;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;     // ldmatrix: this lane supplies row lrow of the 16-row tile ...
    const int lk = (lane >> 4) * 4;                           // ... at channel offset 0 (matrices 0, 1) or 4 (matrices 2, 3)
    float* yout = y4 + (size_t)n * (h + 1) * WQ * NOUT;
    for (int mp = warp; 2 * mp < mtiles; mp += WARPS) {
        uint32_t base[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int l = min(16 * (2 * mp + i) + lrow, npos - 1);
            const int py = l / WQ, px = l - py * WQ;
            base[i] = xs_addr + (uint32_t)(((py * XW + px) * XP + lk) * 4);           // window origin: staged pixel (py, px)
        }
        float acc[2][NTILES][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int tap = ks >> 2, u = tap >> 1, v = tap & 1, c8 = ks & 3;
            const uint32_t off = (uint32_t)(((u * XW + v) * XP + 8 * c8) * 4);
            uint32_t a0[4], a1[4];
            ldmatrix_x4(base[0] + off, a0);
            ldmatrix_x4(base[1] + off, a1);
#pragma unroll
            for (int np = 0; np < NTILES / 2; ++np) {
                float4 b;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                             : "r"(bs_addr + (uint32_t)(((ks * (NTILES / 2) + np) * 32) * 16)));
                mma_tf32(acc[0][2 * np], a0, __float_as_uint(b.x), __float_as_uint(b.y));
                mma_tf32(acc[1][2 * np], a1, __float_as_uint(b.x), __float_as_uint(b.y));
                mma_tf32(acc[0][2 * np + 1], a0, __float_as_uint(b.z), __float_as_uint(b.w));
                mma_tf32(acc[1][2 * np + 1], a1, __float_as_uint(b.z), __float_as_uint(b.w));
            }
        }
        // accumulator layout: acc[i][nt][0..1] -> position 16 m + g, channels 8 nt + 2 t (+1); [2..3] -> position + 8
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int l = 16 * (2 * mp + i) + g + 8 * half;
                if (l < npos) {
                    float* o = yout + ((size_t)P0 * WQ + l) * NOUT + 2 * t;       // positions of a CTA are contiguous in y4
#pragma unroll
                    for (int nt = 0; nt < NTILES; ++nt)
                        *reinterpret_cast<float2*>(o + 8 * nt) = make_float2(acc[i][nt][2 * half], acc[i][nt][2 * half + 1]);
                }
            }
    }
}

}  // namespace spc
}  // namespace riders

extern "C" size_t riders_subpixel_conv_smem_bytes(int w) {
    return (size_t)(riders::spc::BS_FLOATS + (riders::spc::TP + 1) * (w + 2) * riders::spc::XP) * 4;
}

extern "C" int riders_subpixel_conv_act_nhwc(const float* x, const float* bias_host, float slope, const float* w_frag,
                                             int N, int h, int w, float* y4, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && h > 0 && w > 0, "subpixel_conv: bad shape");
    RIDERS_REQUIRE(slope >= 0.f && slope <= 1.f, "subpixel_conv: LeakyReLU slope must be in [0, 1] (got %g)", (double)slope);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && w_frag && y4, "subpixel_conv: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)w_frag) | ((uintptr_t)y4)) & 15) == 0, "subpixel_conv: 16-byte alignment");
    const size_t smem = riders_subpixel_conv_smem_bytes(w);
    RIDERS_REQUIRE(smem <= 113 * 1024, "subpixel_conv: map too wide (w = %d) for two CTAs per SM", w);
    spc::Params P;
    for (int c = 0; c < spc::CIN; ++c) P.bias[c] = bias_host ? bias_host[c] : 0.f;
    P.slope = slope;
    const int row_tiles = ceil_div(h + 1, spc::TP);
    const long long blocks = (long long)N * row_tiles;
    RIDERS_REQUIRE(blocks < (1ll << 31), "subpixel_conv: too many tiles for one launch");
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(spc::subpixel_conv_kernel, (size_t)(smem), configured, "subpixel_conv: shared memory attribute")) return rc_smem;
    count_launches(1);
    spc::subpixel_conv_kernel<<<(unsigned)blocks, spc::THREADS, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(w_frag), y4, P, h, w, row_tiles);
    return check_launch("subpixel_conv_act_nhwc");
}
