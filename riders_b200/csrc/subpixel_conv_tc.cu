// tcgen05 / TMEM / TMA form of the last decoder block's sub-pixel phase convolution:
//     y4 (N, 64, h+1, w+1) = conv2x2(x, W2, padding 1),  x (N, 32, h, w) the ACTIVATED output of the previous block,
// both channels-last (MultiScaleDecoder deconv0.deconv in its 2x2-tap sub-pixel form, RCNet/networks.py:678-779,
// utils/net_utils.py:474-569; see upsample.cu for the phase algebra).  It feeds riders_decoder_tail_nhwc.
//
// cuDNN runs this shape (K = 128, N = 64, 12.6 M output positions per 2048 points) through a generic implicit GEMM at
// 121 TFLOP/s, 1.7 ms, against an HBM floor of 0.75 ms (1.6 GB in, 3.2 GB out).  Here it is the canonical Blackwell
// pipeline: M = output positions, N = 64, K = 4 taps x 32 channels; the K block of one tap is exactly one 128-byte row of
// a K-major SWIZZLE_128B slab, so a tap's A operand is the input map itself shifted by (u - 1, v - 1): TMA loads it
// straight from global memory (box = 32 channels x (w + 1) pixels of one input row, SWIZZLE_128B, out-of-map parts
// zero-filled = the convolution's padding), no staging code at all.
//   tile  = two position rows (p, p + 1) of one point: slab rows [0, w] and [56, 56 + w] (8-row aligned), M = 128
//   warp 0: TMA producer (8 box loads per tile into a 2-stage ring of 4 x 16 KB slabs, mbarrier complete_tx)
//   warp 1: 16 tcgen05.mma.kind::tf32 (M 128, N 64, K 8) per tile into one of two 64-column TMEM accumulators;
//           tcgen05.commit frees the A stage and publishes the accumulator
//   warps 2..5: accumulator read-back (tcgen05.ld, one TMEM lane quadrant each) into a staging buffer that the bulk-copy
//           engine writes out (the two position rows of a tile are contiguous runs of y4)
// Persistent: one CTA per SM, tiles round-robin.  TF32 flow only (selected where torch.backends.cudnn.allow_tf32): the
// tensor core reads the fp32 activations and truncates them to TF32 like the library kernels do; W2 is rounded (rna) when packed.
#include "common.cuh"
#include "tcgen05.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace riders {
namespace spctc {

using namespace tcg;

constexpr int CIN = 32, NOUT = 64, TILE_M = 128, ROW_SLOT = 56, THREADS = 192;
constexpr int A_SLAB = TILE_M * 128, B_SLAB = NOUT * 128;            // bytes per tap
constexpr int A_STAGE = 4 * A_SLAB;                                   // four tap slabs per tile
constexpr int NSTAGE = 2;
constexpr int OUT_STAGE = TILE_M * NOUT * 4;                          // read-back staging: 128 rows x 256 B, one per accumulator
constexpr int SMEM_BYTES = 1024 + NSTAGE * A_STAGE + 4 * B_SLAB + 2 * OUT_STAGE + 128;       // 230528 of 232448
constexpr uint32_t TMEM_COLS = 128;                                   // two 64-column accumulators

struct Params {
    CUtensorMap map;         // x as a 4-D tensor (32, w, h, N), box (32, w + 1, 1, 1), SWIZZLE_128B, zero OOB fill
    const uint4* w_pack;
    float* y4;
    int h, w, pairs, n_tiles;
    // HALF variant: the epilogue finishes the up-convolution - depth-to-space, + bias, LeakyReLU, fp16 - and writes the map the
    // tcgen05 decoder tail stages as its MMA operand: yh [N][2h][2 planes of 8 channels][2w][8] fp16 (a row = two contiguous
    // planes = what one bulk copy moves into an operand ring slot)
    __half* yh;
    int* status;             // bit 0 raised when a value leaves the fp16 range
    float bias[16];
    float slope;
};
constexpr int THREADS_HALF = 320;                     // + a second group of read-back warps (one per accumulator)

__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

template <bool HALF>
__global__ void __launch_bounds__(HALF ? THREADS_HALF : THREADS, 1)
subpixel_conv_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_base = smem;                                        // NSTAGE x 4 x 16 KB
    uint8_t* b_base = smem + NSTAGE * A_STAGE;                     // 4 x 8 KB
    uint8_t* o_base = b_base + 4 * B_SLAB;                         // 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(o_base + 2 * OUT_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const uint32_t full = smem_u32(bars), empty = smem_u32(bars + NSTAGE), dfull = smem_u32(bars + 2 * NSTAGE), dempty = smem_u32(bars + 2 * NSTAGE + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int WQ = P.w + 1;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(dfull + 8 * s, 1); mbar_init(dempty + 8 * s, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 4 * B_SLAB / 16; i += (HALF ? THREADS_HALF : THREADS)) reinterpret_cast<uint4*>(b_base)[i] = __ldg(P.w_pack + i);
    fence_async_smem();                                            // B was written through the generic proxy, UMMA reads it through the async one
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t k = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++k) {
                const uint32_t s = k % NSTAGE, ph = (k / NSTAGE) & 1;
                const int n = tile / P.pairs, p = (tile - n * P.pairs) * 2;
                const int nrows = min(2, P.h + 1 - p);
                mbar_wait(empty + 8 * s, ph ^ 1);
                mbar_expect_tx(full + 8 * s, (uint32_t)(nrows * 4 * WQ * 128));
#pragma unroll
                for (int tap = 0; tap < 4; ++tap)
                    for (int rr = 0; rr < nrows; ++rr) {
                        const uint32_t dst = smem_u32(a_base + s * A_STAGE + tap * A_SLAB + rr * ROW_SLOT * 128);
                        asm volatile(
                            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                            " [%0], [%1, {%2, %3, %4, %5}], [%6];"
                            ::"r"(dst), "l"(&P.map), "r"(0), "r"((tap & 1) - 1), "r"(p + rr + (tap >> 1) - 1), "r"(n), "r"(full + 8 * s)
                            : "memory");
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t k = 0;
        constexpr uint32_t IDESC = idesc_tf32(NOUT);
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++k) {
            const uint32_t s = k % NSTAGE, ph = (k / NSTAGE) & 1, d = k & 1, dph = (k >> 1) & 1;
            mbar_wait(full + 8 * s, ph);
            mbar_wait(dempty + 8 * d, dph ^ 1);                    // accumulator d drained by the read-back warps
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    const uint64_t ad = smem_desc(smem_u32(a_base + s * A_STAGE + tap * A_SLAB)), bd = smem_desc(smem_u32(b_base + tap * B_SLAB));
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8)
                        umma_tf32(tmem + d * NOUT, ad + (uint64_t)(2 * k8), bd + (uint64_t)(2 * k8), IDESC, (tap | k8) ? 1u : 0u);
                }
                umma_commit(empty + 8 * s);                        // A stage reusable once these MMAs retire
                umma_commit(dfull + 8 * d);                        // ... and the accumulator is complete
            }
            __syncwarp();
        }
    } else if (HALF) {
        // ===================== read-back + finish (HALF): two groups of four warps, group g drains accumulator g ==============
        const int quad = warp & 3, grp = (warp - 2) >> 2;
        const int row = quad * 32 + lane;                          // slab row = TMEM lane
        const int r2 = row >= ROW_SLOT ? 1 : 0, J = row - r2 * ROW_SLOT;          // y4 row p + r2, y4 column J
        const int H = 2 * P.h, W = 2 * P.w;
        const int leader = 64 + grp * 128;                         // first thread of the group
        const uint32_t ROWB = (uint32_t)W * 32;                    // bytes of one output row (2 planes x W x 16 B)
        uint8_t* stage = o_base + grp * OUT_STAGE;                 // [4 output rows][2 planes][W][16 B] <= 14 KB
        __half2 amax2 = __float2half2_rn(0.f);
        const float slope = P.slope;
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++k) {
            if ((int)(k & 1) != grp) continue;
            const uint32_t ph = (k >> 1) & 1;
            const int n = tile / P.pairs, p = (tile - n * P.pairs) * 2;
            mbar_wait(dfull + 8 * grp, ph);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + grp * NOUT;
            TMEM_LD32(taddr, r0);
            TMEM_LD32(taddr + 32, r1);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(dempty + 8 * grp);
            if (tid == leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the stores that read this buffer are done
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            // phase (a, b) of y4 position (I, J) is output pixel (2 I - a, 2 J - b); local output row lr = Y - (2 p - 1)
            const int I = p + r2;
            if (I <= P.h && J <= P.w) {
#pragma unroll
                for (int ab = 0; ab < 4; ++ab) {
                    const int a = ab >> 1, b = ab & 1;
                    const int Y = 2 * I - a, X = 2 * J - b;
                    if ((unsigned)Y < (unsigned)H && (unsigned)X < (unsigned)W) {
                        const uint32_t* v = ab < 2 ? r0 + 16 * ab : r1 + 16 * (ab - 2);
                        uint32_t hv[8];
#pragma unroll
                        for (int c = 0; c < 16; c += 2) {
                            float v0 = __uint_as_float(v[c]) + P.bias[c], v1 = __uint_as_float(v[c + 1]) + P.bias[c + 1];
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
                for (int lr = 0; lr < 4; ++lr) {
                    const int Y = 2 * p - 1 + lr;
                    if ((unsigned)Y >= (unsigned)H) continue;
                    __half* gdst = P.yh + ((size_t)n * H + Y) * (size_t)W * 16;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(gdst), "r"(smem_u32(stage + lr * ROWB)), "r"(ROWB) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (tid == leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        {
            const float2 m = __half22float2(amax2);
            if (P.status && !(fmaxf(m.x, m.y) <= 65504.f)) atomicOr(P.status, 1);
        }
    } else {
        // ===================== read-back warps: TMEM lane quadrant = warp % 4 =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                          // slab row = TMEM lane
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++k) {
            const uint32_t s = k & 1, ph = (k >> 1) & 1;              // accumulator index / phase
            const int n = tile / P.pairs, p = (tile - n * P.pairs) * 2;
            mbar_wait(dfull + 8 * s, ph);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + s * NOUT;
            TMEM_LD32(taddr, r0);
            TMEM_LD32(taddr + 32, r1);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(dempty + 8 * s);
            // The two position rows of a tile are one contiguous run of y4 each: stage the accumulator rows in shared memory
            // and let the bulk-copy engine write them (two 13 KB stores per tile) instead of 16 strided 16-byte stores per thread.
            if (tid == 64) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the store that last read this buffer is done
            asm volatile("bar.sync 1, 128;" ::: "memory");
            uint4* orow = reinterpret_cast<uint4*>(o_base + s * OUT_STAGE + row * (NOUT * 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) orow[j] = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) orow[8 + j] = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
            fence_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 64) {
                const int nrows = min(2, P.h + 1 - p);
                for (int r2 = 0; r2 < nrows; ++r2) {
                    float* gdst = P.y4 + (((size_t)n * (P.h + 1) + p + r2) * WQ) * NOUT;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(gdst), "r"(smem_u32(o_base + s * OUT_STAGE + r2 * ROW_SLOT * NOUT * 4)), "r"(WQ * NOUT * 4) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (tid == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

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
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace spctc
}  // namespace riders

static int subpixel_conv_tc_launch(const float* x, const float* w_pack, int N, int h, int w, float* y4, __half* yh,
                                   const float* bias_host, float slope, int* status, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && h > 0 && w > 0, "subpixel_conv_tc: bad shape");
    RIDERS_REQUIRE(w + 1 <= spctc::ROW_SLOT, "subpixel_conv_tc: at most %d columns (got w = %d)", spctc::ROW_SLOT - 1, w);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && w_pack && (y4 || yh), "subpixel_conv_tc: null pointer");
    RIDERS_REQUIRE(((((uintptr_t)x) | ((uintptr_t)w_pack) | ((uintptr_t)y4) | ((uintptr_t)yh)) & 15) == 0, "subpixel_conv_tc: 16-byte alignment");
    spctc::EncodeTiledFn enc = spctc::get_encode_fn();
    if (!enc) { set_error("subpixel_conv_tc: cuTensorMapEncodeTiled unavailable"); return RIDERS_ECUDA; }
    spctc::Params P;
    cuuint64_t dims[4] = {(cuuint64_t)spctc::CIN, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)spctc::CIN * 4, (cuuint64_t)w * spctc::CIN * 4, (cuuint64_t)h * w * spctc::CIN * 4};
    cuuint32_t box[4] = {(cuuint32_t)spctc::CIN, (cuuint32_t)(w + 1), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&P.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("subpixel_conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr); return RIDERS_ECUDA; }
    P.w_pack = reinterpret_cast<const uint4*>(w_pack);
    P.y4 = y4;
    P.yh = yh;
    P.status = status;
    for (int c = 0; c < 16; ++c) P.bias[c] = bias_host ? bias_host[c] : 0.f;
    P.slope = slope;
    P.h = h; P.w = w;
    P.pairs = (h + 2) / 2;
    const long long tiles = (long long)N * P.pairs;
    RIDERS_REQUIRE(tiles < (1ll << 31), "subpixel_conv_tc: too many tiles for one launch");
    P.n_tiles = (int)tiles;
    const int grid = tiles < sm_count() ? (int)tiles : sm_count();
    count_launches(1);
    if (yh) {
        static size_t configured[64] = {0};
        if (int rc_smem = ensure_dynamic_smem(spctc::subpixel_conv_tc_kernel<true>, (size_t)(spctc::SMEM_BYTES), configured, "subpixel_conv_tc: shared memory attribute")) return rc_smem;
        spctc::subpixel_conv_tc_kernel<true><<<grid, spctc::THREADS_HALF, spctc::SMEM_BYTES, (cudaStream_t)stream>>>(P);
    } else {
        static size_t configured[64] = {0};
        if (int rc_smem = ensure_dynamic_smem(spctc::subpixel_conv_tc_kernel<false>, (size_t)(spctc::SMEM_BYTES), configured, "subpixel_conv_tc: shared memory attribute")) return rc_smem;
        spctc::subpixel_conv_tc_kernel<false><<<grid, spctc::THREADS, spctc::SMEM_BYTES, (cudaStream_t)stream>>>(P);
    }
    return check_launch("subpixel_conv_tc_nhwc");
}

extern "C" int riders_subpixel_conv_tc_nhwc(const float* x, const float* w_pack, int N, int h, int w, float* y4, void* stream) {
    RIDERS_REQUIRE(y4 || N == 0, "subpixel_conv_tc: null pointer");
    return subpixel_conv_tc_launch(x, w_pack, N, h, w, y4, nullptr, nullptr, 1.f, nullptr, stream);
}

extern "C" int riders_subpixel_conv_tc_half_nhwc(const float* x, const float* w_pack, int N, int h, int w, const float* bias_host,
                                                 float slope, void* yh, int* status, void* stream) {
    RIDERS_REQUIRE(yh || N == 0, "subpixel_conv_tc_half: null pointer");
    RIDERS_REQUIRE(slope >= 0.f && slope <= 1.f, "subpixel_conv_tc_half: LeakyReLU slope must be in [0, 1] (got %g)", (double)slope);
    return subpixel_conv_tc_launch(x, w_pack, N, h, w, nullptr, reinterpret_cast<__half*>(yh), bias_host, slope, status, stream);
}
