// A3 fused LoFTREncoderLayer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
// Reference semantics: RCNet/linear_attention.py:112-135 (layer) and :18-45 (linear attention).
//
// Specialised for RC-Net's transformer: d_model 128, 8 heads of 16, 21 tokens per radar point
// (networks.py:378, :441).  One CTA per SM, persistent over tiles of 6 points (126 token rows
// padded to the UMMA M of 128).  Every dense contraction of the layer - q/k/v projections,
// merge, the 256->256->128 MLP - is a tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate)
// with both operands in shared memory (K-major, 128-byte swizzle) and the accumulator in tensor
// memory.  fp32 accuracy comes from the 2-term operand split  a = a_hi + a_lo  with
// a_hi = fp16(a), a_lo = fp16(a - a_hi)  (22-23 significant bits; products of fp16 values are
// exact in the fp32 accumulator) and the 3 MMAs  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi.  Compared with
// a 3xTF32 split this halves the tensor-pipe time (K = 16 per instruction instead of 8) and the
// operand bytes, which is what lets x stay resident next to the working operand.  Weights are
// pre-scaled by 2^8 when packed (exact; keeps the lo halves of small weights out of the fp16
// subnormal range) and every accumulator read-back multiplies by 2^-8.  Activations must stay
// inside the fp16 range (|a| < 65504), the same constraint the reference's V/S scaling exists for.
//
// One launch runs the WHOLE transformer (RCNet/linear_attention.py:159-184): tokens of a radar point never interact with other
// points, so a tile of 6 points walks through all 16 layer calls back to back inside one persistent CTA (tile-major order); its
// tokens ping through the L2-resident output buffers, the 16 per-layer launches, their wave tails and 15/16 of the HBM token
// traffic are gone.
//
// The linear attention itself (KV = K'^T V, normaliser, Q' KV: 2.5 % of the FLOPs but 40 % of the round-1 tile time on CUDA
// cores) runs on the tensor cores too: q / k / v leave TMEM as fp16 hi/lo pairs, and each (point, head) pair is a handful of
// warp-level mma.sync.m16n8k16 with the same 3-term split (KV^T = V^T K' accumulates in registers and IS the B fragment of
// the Q' KV product; the normaliser rides along as an all-ones A row).  tcgen05 cannot take these: M = 21 tokens / 16 x 16
// blocks per (point, head) against its M = 128 tile.
//
// Roles (576 threads): warp 0 streams the pre-split, pre-swizzled weight chunks (32 KB: hi|lo of
// a 128 x 64 block) from L2 with cp.async.bulk through a 3-stage mbarrier ring; one elected lane of
// warp 1 issues the MMAs (N = 128: an SS-mode MMA costs ~43 + N/2 cycles on this part, tools/mma_rate.py,
// so wide instructions matter) and commits their completion to mbarriers; warps 2..17 are the epilogue: four
// warps per TMEM lane quadrant, thread = (token row, 32 of its 128 columns): elu+1, the
// per-(point, head) KV / normaliser contraction (2.5 % of the layer's FLOPs, CUDA cores),
// LayerNorms, ReLU, residual, and the hi/lo re-split of the next A operand straight into the
// swizzled layout.
//
// TMEM columns (512): D0 [0,128) q -> merge -> mlp.2 | D1 [128,256) k | D2 [256,384) v -> h[:,128:]
//                     D3 [384,512) h[:, :128]
// Shared memory: X operand (hi[2] | lo[2] slabs of 128 rows x 128 B) 64 KB: x for q/k/v, then the attention staging (one
// quarter = 2 heads at a time), then x again (re-read from L2 while the merge projection runs) for its half of mlp.0, then
// relu(h[:,128:]) | A operand 64 KB: src (cross), then the message - written by the attention warps quarter by quarter, never
// held in registers -, m, relu(h[:, :128]) and the output staging | 3 weight stages of 32 KB | LayerNorm partials 2 KB.
// The x half of mlp.0 (cat([x, message]) @ W1^T, linear_attention.py:132) does not depend on the attention: it is issued as
// soon as x is back in X and runs under the merge read-back and LayerNorm 1.
#include "common.cuh"
#include <cuda_fp16.h>

namespace riders {
namespace tc {

constexpr int TOK = 21, CH = 128, NHEAD = 8, HD = 16;
constexpr int TILE_M = 128, PTS = 6, ROWS = PTS * TOK;        // 126 valid rows per tile
constexpr int SLAB = 128 * 128;                                // bytes: 128 rows x 64 fp16
constexpr int OPER_BYTES = 4 * SLAB;                           // hi[2] | lo[2] of a 128 x 128 operand
constexpr int CHUNK = 2 * SLAB;                                // hi (16 KB) | lo (16 KB) of one 128 x 64 weight block
constexpr int NSTAGE = 3;
constexpr int NBLOCK = 10;                                     // 128 x 128 weight blocks per layer, in issue order
constexpr int NCHUNK = NBLOCK * 2;
constexpr int LAYER_PACK_BYTES = NCHUNK * CHUNK;               // 640 KB per layer
constexpr int MAX_LAYERS = 16;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + EPI_WARPS * 32;                   // warp 0 producer, warp 1 MMA, warps 2..17 epilogue
constexpr int TAIL_BYTES = 128;                                // barriers + TMEM slot
constexpr int SMEM_BYTES = 2 * OPER_BYTES + NSTAGE * CHUNK + TAIL_BYTES;   // 229504 of 232448; base must be 1024-aligned
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t D0 = 0, D1 = 128, D2 = 256, D3 = 384;
constexpr float W_SCALE = 256.0f, W_UNSCALE = 1.0f / 256.0f;
// instruction descriptor: D fp32, A/B fp16 (format 0), both K-major, M = 128; N in bits 17..22
constexpr uint32_t idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one lane of a converged warp (the pattern ptxas recognises: no per-MMA election loop in the SASS)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte-swizzled operand slab (rows x 64 fp16): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

#define TMEM_LD32(taddr, r)                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                      \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                     \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),     \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),   \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),   \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                        \
                 : "r"(taddr))
#define TMEM_LD16(taddr, r)                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                      \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                              \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),     \
                   "=r"(r[15])                                                                                  \
                 : "r"(taddr))
#define TMEM_LD8(taddr, r)                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])    \
                 : "r"(taddr))
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the 16 epilogue warps

// ------------------------------------------------------------------------------------ operand staging
// byte offset of 16-byte segment `seg` (0..7) of row `row` inside one swizzled slab
__device__ __forceinline__ uint32_t sw128(int row, int seg) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((seg ^ (row & 7)) << 4));
}
// fp16-range guard: the hi/lo split needs |a| <= 65504 (the fp32 reference has no such limit).  Every operand value passes
// through split2, which records a violation (or a NaN) in `g_bad`; the kernel raises the caller's overflow flag and the host
// re-runs the call on the fp32 SIMT kernels (rcnet/linear_attention.py).  One FSETP per value, folded into the compare chain.
#define RANGE_CHECK(a, b) bad = fmaxf(bad, fmaxf(fabsf(a), fabsf(b)))            /* running max |value|: one FMNMX each */
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// split 8 consecutive K elements (col % 8 == 0) of row `row` into the hi / lo slabs of the operand at `base`
__device__ __forceinline__ void split_store8(uint8_t* base, int row, int col, float4 a, float4 b, float& bad) {
    const uint32_t off = (uint32_t)(col >> 6) * SLAB + sw128(row, (col & 63) >> 3);
    RANGE_CHECK(a.x, a.y); RANGE_CHECK(a.z, a.w); RANGE_CHECK(b.x, b.y); RANGE_CHECK(b.z, b.w);
    uint4 hi, lo;
    split2(a.x, a.y, hi.x, lo.x); split2(a.z, a.w, hi.y, lo.y);
    split2(b.x, b.y, hi.z, lo.z); split2(b.z, b.w, hi.w, lo.w);
    *reinterpret_cast<uint4*>(base + off) = hi;
    *reinterpret_cast<uint4*>(base + 2 * SLAB + off) = lo;
}

// coalesced: epilogue warp `ew` (0..15) loads rows ew*8.. of a (rows x 128) fp32 tile and splits them into the operand
__device__ __forceinline__ void load_tile_regs(float4 (&v)[8], const float* g, int valid_rows, int ew, int lane) {
    const int c8 = (lane & 15) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = ew * 8 + i * 2 + (lane >> 4);
        if (row < valid_rows) {
            const float4* src = reinterpret_cast<const float4*>(g + (size_t)row * CH + c8);
            v[2 * i] = __ldcg(src); v[2 * i + 1] = __ldcg(src + 1);     // L2: the tile itself may have written these rows in an earlier layer call
        } else {
            v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f); v[2 * i + 1] = v[2 * i];
        }
    }
}
__device__ __forceinline__ void split_tile_regs(uint8_t* base, const float4 (&v)[8], int ew, int lane, float& bad) {
    const int c8 = (lane & 15) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) split_store8(base, ew * 8 + i * 2 + (lane >> 4), c8, v[2 * i], v[2 * i + 1], bad);
}
// The same rows, asynchronously (LDGSTS) into a raw fp32 image of the tile (row stride 512 bytes; the whole X operand region):
// issued while LayerNorm 2 of the previous call runs, consumed by the SAME thread (wait_group, then its own 16-byte pieces).
__device__ __forceinline__ void prefetch_tile_async(uint8_t* raw, const float* g, int valid_rows, int ew, int lane) {
    const int c8 = (lane & 15) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = ew * 8 + i * 2 + (lane >> 4);
        if (row < valid_rows) {
            const uint32_t dst = smem_u32(raw + ((size_t)row * CH + c8) * 4);
            const float* src = g + (size_t)row * CH + c8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 4) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void read_prefetched_tile(float4 (&v)[8], const uint8_t* raw, int valid_rows, int ew, int lane) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int c8 = (lane & 15) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = ew * 8 + i * 2 + (lane >> 4);
        if (row < valid_rows) {
            const float4* src = reinterpret_cast<const float4*>(raw + ((size_t)row * CH + c8) * 4);
            v[2 * i] = src[0]; v[2 * i + 1] = src[1];
        } else {
            v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f); v[2 * i + 1] = v[2 * i];
        }
    }
}
__device__ __forceinline__ void load_tile_to_oper(uint8_t* base, const float* g, int valid_rows, int ew, int lane, float& bad) {
    float4 v[8];
    load_tile_regs(v, g, valid_rows, ew, lane);
    split_tile_regs(base, v, ew, lane, bad);
}

// ------------------------------------------------------------------------------------ warp-level MMA (attention blocks)
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D (16 x 8, fp32) += A (16 x 16, fp16, row) * B (16 x 8, fp16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// attention staging in the idle A operand: six arrays (Q' hi, Q' lo, K' hi, K' lo, V hi, V lo) of 128 token rows x 32 fp16 (one
// quarter = 2 heads), rows 80 bytes apart so that the 8 rows of an ldmatrix tile (and the 8 lanes of a 16-byte store phase) fall
// into distinct 16-byte bank groups
constexpr int SROW = 80, SARR = 128 * SROW;
static_assert(6 * SARR <= OPER_BYTES, "attention staging must fit the X operand");

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.0f : __expf(x); }   // elu(x) + 1 (SFU exp2, ~2 ulp)
__device__ __forceinline__ float acc_f(uint32_t r) { return __uint_as_float(r) * W_UNSCALE; }  // accumulator read-back

// LayerNorm statistics of a 128-wide row held as 4 x 32 values by the four threads (column quarters) of the row, with ONE
// exchange: every thread reduces its quarter to (mean_i, M2_i) locally, the four pairs meet in shared memory and are merged with
// Chan's parallel formula  M2 = sum M2_i + 32 sum (mean_i - mean)^2  in a fixed order, so all four threads get identical bits and
// there is no E[x^2] - mean^2 cancellation.  (Round 1 exchanged the sums and the squared deviations separately: two barriers more.)
// `buf`: 128 x 4 float2 = 4 KB of the idle A operand.
__device__ __forceinline__ void row_stats(const float (&v)[32], int row, int colq, float2* buf, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    const float mi = s * (1.0f / 32.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = v[i] - mi; q = fmaf(d, d, q); }
    buf[colq * 128 + row] = make_float2(mi, q);
    epi_bar();
    const float2 p0 = buf[row], p1 = buf[128 + row], p2 = buf[256 + row], p3 = buf[384 + row];
    mean = ((p0.x + p1.x) + (p2.x + p3.x)) * 0.25f;
    const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
    const float m2 = ((p0.y + p1.y) + (p2.y + p3.y)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    rstd = 1.0f / sqrtf(m2 * (1.0f / CH) + eps);
}

// signal "operand written (and my TMEM reads done)" to the MMA warp
// (every thread fences its own writes, the warp converges, one lane arrives: 16 arrivals per phase instead of 512 - the
// arrivals are shared-memory atomics that serialise)
__device__ __forceinline__ void publish(uint32_t bar) {
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

constexpr int MAX_CALLS = 16;                   // layer calls per launch (RC-Net: 8 layers x 2 calls)
struct CallArgs {
    const float* x; const float* src; float* out;   // out may alias x (in place): every element is read and written by one thread
    const uint8_t* wpack;                       // NCHUNK x 16 KB, pre-split + pre-swizzled
    const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;
    int is_self;                                // src == x
};
struct LayerArgs {
    CallArgs call[MAX_CALLS];
    int n_calls;
    int N;                                      // radar points
    int n_tiles;
    float attn_eps, ln_eps;
    long long* trace;                           // optional phase timestamps (riders_debug_loftr_trace), else null
    int* overflow;                              // set to 1 when an operand left the fp16 range (see RANGE_CHECK)
};

// trace layout per traced CTA (blockIdx.x == 0 or gridDim.x / 2): 128 slots; [0..63] epilogue thread, [64..127] MMA thread
#define TRACE_E(k) do { if (P.trace && tr_on && lane == 0 && warp == 4 && ci == 0 && tile == tr_tile) P.trace[tr_base + (k)] = clock64(); } while (0)
#define TRACE_M(k) do { if (P.trace && tr_on && lane == 0 && ci == 0 && tile == tr_tile) P.trace[tr_base + 64 + (k)] = clock64(); } while (0)
#define TRACE_W(k) do { if (P.trace && tr_on && lane == 0 && ci == 0 && tile == tr_tile) P.trace[tr_base + 64 + (k)] = wsum; } while (0)

// ------------------------------------------------------------------------------------ the fused layer
__global__ void __launch_bounds__(THREADS, 1)
loftr_layer_tc_kernel(const LayerArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;                                 // SWIZZLE_128B slabs need a 1024-byte aligned base
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* x_base = smem;                                   // X operand: hi[2] | lo[2]
    uint8_t* a_base = smem + OPER_BYTES;                      // A operand: hi[2] | lo[2]
    uint8_t* b_base = smem + 2 * OPER_BYTES;                  // NSTAGE x (hi | lo)
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + NSTAGE * CHUNK);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const uint32_t bfull = smem_u32(bars), bempty = smem_u32(bars + NSTAGE);
    const uint32_t a_ready = smem_u32(bars + 2 * NSTAGE), d_ready = smem_u32(bars + 2 * NSTAGE + 1);
    const uint32_t x_ready = smem_u32(bars + 2 * NSTAGE + 2), h_ready = smem_u32(bars + 2 * NSTAGE + 3);
    // one barrier per attention quarter (a phase per call each: the epilogue can never run two phases ahead of the MMA warp)
    const uint32_t m_ready0 = smem_u32(bars + 10);
    auto m_ready = [&](int q) { return m_ready0 + 8u * (uint32_t)(q + (q >> 1)); };       // slots 10, 11, 13, 14 (12 = TMEM slot)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr_on = blockIdx.x == 0 || blockIdx.x == gridDim.x / 2;
    const int tr_base = (blockIdx.x == 0 ? 0 : 128);
    const int tr_tile = (int)(blockIdx.x + gridDim.x) < P.n_tiles ? (int)(blockIdx.x + gridDim.x) : (int)blockIdx.x;   // steady state if possible

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
        mbar_init(a_ready, EPI_WARPS);
        mbar_init(d_ready, 1);
        mbar_init(x_ready, EPI_WARPS);
        mbar_init(h_ready, 1);
        for (int q = 0; q < 4; ++q) mbar_init(m_ready(q), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== weight producer =====================
        if (lane == 0) {
            uint32_t n = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
                for (int ci = 0; ci < P.n_calls; ++ci)
                    for (int c = 0; c < NCHUNK; ++c, ++n) {
                        const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1;
                        mbar_wait(bempty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bfull + 8 * s, CHUNK);
                        bulk_g2s(smem_u32(b_base + s * CHUNK), P.call[ci].wpack + (size_t)c * CHUNK, CHUNK, bfull + 8 * s);
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        {
            uint32_t n = 0, aph = 0, vph = 0;
            const uint32_t xo = smem_u32(x_base), ao = smem_u32(a_base), b0 = smem_u32(b_base);
            constexpr uint32_t IDESC = idesc_f16(128);
            long long wsum = 0;
            // one 128 x 128 weight block = 2 chunks (K halves): D[:, dcol .. dcol + 128) (+)= oper W^T
            auto issue_block = [&](uint32_t oper, uint32_t dcol, uint32_t acc_in, uint32_t commit_bar) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb, ++n) {
                    const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1;
                    const long long tw = P.trace ? clock64() : 0;
                    mbar_wait(bfull + 8 * s, ph);
                    if (P.trace) wsum += clock64() - tw;
                    tc_fence_after();
                    const uint32_t d = tmem + dcol;
                    const uint64_t ah = smem_desc(oper + kb * SLAB), al = smem_desc(oper + (2 + kb) * SLAB);
                    const uint64_t bh = smem_desc(b0 + s * CHUNK), bl = smem_desc(b0 + s * CHUNK + SLAB);
                    if (elect_one()) {
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {          // 16 fp16 = 32 bytes per MMA along K
                            const uint64_t adv = (uint64_t)(k16 * 2);
                            umma_f16(d, al + adv, bh + adv, IDESC, acc_in | (uint32_t)kb | (k16 > 0));
                            umma_f16(d, ah + adv, bl + adv, IDESC, 1);
                            umma_f16(d, ah + adv, bh + adv, IDESC, 1);
                        }
                        umma_commit(bempty + 8 * s);                 // stage reusable once these MMAs retire
                        if (kb == 1 && commit_bar) umma_commit(commit_bar);
                    }
                    __syncwarp();
                }
            };
            // merge(msg), one attention quarter (2 heads = 32 message columns = 2 of the 8 k steps) at a time -> D3
            uint32_t mph = 0;
            auto issue_merge_quarter = [&](int hq) {
                const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1;
                mbar_wait(m_ready(hq), mph & 1);
                if ((hq & 1) == 0) {
                    const long long tw = P.trace ? clock64() : 0;
                    mbar_wait(bfull + 8 * s, ph);
                    if (P.trace) wsum += clock64() - tw;
                }
                tc_fence_after();
                const int kb = hq >> 1;
                const uint64_t ah = smem_desc(ao + kb * SLAB), al = smem_desc(ao + (2 + kb) * SLAB);
                const uint64_t bh = smem_desc(b0 + s * CHUNK), bl = smem_desc(b0 + s * CHUNK + SLAB);
                if (elect_one()) {
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        const uint64_t adv = (uint64_t)((((hq & 1) << 1) + k2) * 2);
                        umma_f16(tmem + D3, al + adv, bh + adv, IDESC, (uint32_t)(hq > 0 || k2 > 0));
                        umma_f16(tmem + D3, ah + adv, bl + adv, IDESC, 1);
                        umma_f16(tmem + D3, ah + adv, bh + adv, IDESC, 1);
                    }
                    if (hq & 1) umma_commit(bempty + 8 * s);
                    if (hq == 3) umma_commit(d_ready);
                }
                __syncwarp();
                if (hq & 1) ++n;
            };
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
            for (int ci = 0; ci < P.n_calls; ++ci) {
                // a cross call whose source is the map the previous call of this tile just wrote finds it already split in X
                // (written by that call's LayerNorm 2): then x goes to A and the operand roles of seg 0 are swapped
                const bool handoff = ci > 0 && !P.call[ci].is_self && P.call[ci].src == P.call[ci - 1].out;
                const uint32_t q_src = handoff ? ao : xo;
                const uint32_t kv_src = (P.call[ci].is_self || handoff) ? xo : ao;
                // seg 0: k, v, q (-> the attention can start).  After a hand-off the source is in X before this call begins: k and
                // v go at once, q when x has arrived in A.
                TRACE_M(0);
                if (handoff) {
                    mbar_wait(x_ready, vph & 1); ++vph;
                    tc_fence_after();
                    TRACE_M(1); wsum = 0;
                    issue_block(kv_src, D1, 0, 0);
                    issue_block(kv_src, D2, 0, 0);
                    mbar_wait(a_ready, aph & 1); ++aph;
                    tc_fence_after();
                } else {
                    mbar_wait(a_ready, aph & 1); ++aph;
                    tc_fence_after();
                    TRACE_M(1); wsum = 0;
                    issue_block(kv_src, D1, 0, 0);
                    issue_block(kv_src, D2, 0, 0);
                }
                issue_block(q_src, D0, 0, d_ready);
                TRACE_M(2);
                TRACE_W(3);
                // seg 1: merge(msg) -> D3, two k steps per finished attention quarter: only the last quarter's 6 MMAs are exposed
                TRACE_M(4);
                wsum = 0;
                issue_merge_quarter(0); issue_merge_quarter(1); issue_merge_quarter(2);
                TRACE_M(5);
                issue_merge_quarter(3);
                ++mph;
                TRACE_M(6);
                TRACE_W(7);
                // seg 2: x is back in X -> its half of mlp.0 (runs under the merge read-back and LayerNorm 1); k and v are dead
                TRACE_M(8);
                mbar_wait(x_ready, vph & 1); ++vph;
                tc_fence_after();
                TRACE_M(9); wsum = 0;
                issue_block(xo, D1, 0, 0);
                issue_block(xo, D2, 0, 0);
                TRACE_M(10);
                TRACE_W(11);
                // seg 3: mlp.0[:, C:] on m, accumulated onto the x halves
                TRACE_M(12);
                mbar_wait(a_ready, aph & 1); ++aph;
                tc_fence_after();
                TRACE_M(13); wsum = 0;
                issue_block(ao, D1, 1, h_ready);             // h[:, :C] complete (own barrier): its ReLU / re-split runs under the next block
                issue_block(ao, D2, 1, d_ready);
                TRACE_M(14);
                TRACE_W(15);
                // seg 4: mlp.2 on relu(h): [:, :C] from X as soon as that half is split (x_ready), [:, C:] from A
                TRACE_M(16);
                mbar_wait(x_ready, vph & 1); ++vph;
                tc_fence_after();
                TRACE_M(17); wsum = 0;
                issue_block(xo, D0, 0, 0);
                mbar_wait(a_ready, aph & 1); ++aph;
                tc_fence_after();
                issue_block(ao, D0, 1, d_ready);
                TRACE_M(18);
                TRACE_W(19);
            }
        }
    } else {
        // ===================== epilogue warps: thread = (token row, column quarter) =====================
        const int ew = warp - 2;                         // 0..15
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int colq = ew >> 2;                        // column quarter: 32 columns = heads 2 colq, 2 colq + 1
        const int row = quad * 32 + lane;
        const int col0 = colq * 32;
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
        uint32_t dph = 0, hph = 0;
        float2* red = reinterpret_cast<float2*>(a_base);  // LayerNorm partials: 4 KB of the A operand, dead whenever a LayerNorm runs

        float bad = 0.f;                                 // running max |operand value| of the rows that exist (RANGE_CHECK)
        bool pre_valid = false;                          // this call's x rows are arriving in X as a raw fp32 image (prefetch_tile_async)
        const int g8 = lane >> 2, t4 = lane & 3, r8 = lane & 7, tl = lane >> 3;      // mma.sync / ldmatrix lane coordinates
        const uint32_t stage_u32 = smem_u32(x_base);     // attention staging lives in X (dead between the q/k/v MMAs and the x reload)

        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
        for (int ci = 0; ci < P.n_calls; ++ci) {
            const CallArgs& A = P.call[ci];
            const int pts = min(PTS, P.N - tile * PTS);
            const int valid_rows = pts * TOK;
            const size_t g_row0 = (size_t)tile * ROWS;
            const float* xg = A.x + g_row0 * CH;
            const float* sg = A.src + g_row0 * CH;

            // ---- E0: X = x, A = src (cross); after a hand-off X already holds src and x goes to A ---------------
            TRACE_E(0);
            const bool handoff = ci > 0 && !A.is_self && A.src == P.call[ci - 1].out;      // same predicate as the MMA warp
            if (handoff) {
                publish(x_ready);                        // X = split(source), written by the previous call's LayerNorm 2: k and v may go
                load_tile_to_oper(a_base, xg, valid_rows, ew, lane, bad);
            } else {
                if (pre_valid) {
                    float4 v[8];
                    read_prefetched_tile(v, x_base, valid_rows, ew, lane);
                    epi_bar();                           // every thread holds its pieces: the split may overwrite the raw image
                    split_tile_regs(x_base, v, ew, lane, bad);
                } else {
                    load_tile_to_oper(x_base, xg, valid_rows, ew, lane, bad);
                }
                if (!A.is_self) load_tile_to_oper(a_base, sg, valid_rows, ew, lane, bad);
            }
            publish(a_ready);
            TRACE_E(1);

            // ---- E1: linear attention (linear_attention.py:29-43) ---------------------------
            // The reference divides V by S and multiplies the message by S (fp16 range safety,
            // linear_attention.py:39-43); in fp32 the two cancel to rounding, so neither is applied.
            mbar_wait(d_ready, dph & 1); ++dph;          // q, k, v in D0, D1, D2
            tc_fence_after();
            TRACE_E(2);
            {
                // One quarter (heads 2 hq, 2 hq + 1 = 32 columns) at a time:
                // (1) every thread moves 8 columns of q, k, v of its token row from TMEM, applies elu + 1 to q and k, splits
                //     each value into fp16 hi / lo and stores 16-byte pieces into the six staging arrays;
                // (2) warp w < 12 owns the (point, head) pair w and runs it on the tensor cores (mma.sync.m16n8k16, fp32
                //     accumulate, 3-term split):  KV^T[v][d] = sum_s V[s][v] K'[s][d]  (A = V^T and B = K' both through
                //     ldmatrix.trans; tokens 21..31 of the second k step are masked out of A) together with
                //     Ksum[d] = sum_s K'[s][d] as an all-ones A;  the accumulator fragment of KV^T is exactly the B fragment
                //     of  msg[l][v] = sum_d Q'[l][d] KV[d][v]  (rows = tokens, 2 m tiles), the normaliser Q' . Ksum is one
                //     more n tile of the same product.  Each message block goes straight into the A operand (hi / lo, swizzled)
                //     for the merge projection - the staging lives in X, so nothing is held back in registers.
                const int pair = ew, p = pair >> 1, hl = pair & 1;
                const bool worker = ew < 12 && p < pts;
#pragma unroll
                for (int hq = 0; hq < 4; ++hq) {
                    {
                        uint32_t qr[8], kr[8], vr[8];
                        TMEM_LD8(trow + D0 + hq * 32 + colq * 8, qr);
                        TMEM_LD8(trow + D1 + hq * 32 + colq * 8, kr);
                        TMEM_LD8(trow + D2 + hq * 32 + colq * 8, vr);
                        tmem_wait_ld();
                        if (hq == 1) TRACE_E(20);
                        uint4 h4, l4;
                        uint8_t* srow = x_base + row * SROW + colq * 16;
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = elu1(acc_f(qr[i]));
                        RANGE_CHECK(f[0], f[1]); RANGE_CHECK(f[2], f[3]); RANGE_CHECK(f[4], f[5]); RANGE_CHECK(f[6], f[7]);
                        split2(f[0], f[1], h4.x, l4.x); split2(f[2], f[3], h4.y, l4.y); split2(f[4], f[5], h4.z, l4.z); split2(f[6], f[7], h4.w, l4.w);
                        *reinterpret_cast<uint4*>(srow) = h4; *reinterpret_cast<uint4*>(srow + SARR) = l4;
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = elu1(acc_f(kr[i]));
                        RANGE_CHECK(f[0], f[1]); RANGE_CHECK(f[2], f[3]); RANGE_CHECK(f[4], f[5]); RANGE_CHECK(f[6], f[7]);
                        split2(f[0], f[1], h4.x, l4.x); split2(f[2], f[3], h4.y, l4.y); split2(f[4], f[5], h4.z, l4.z); split2(f[6], f[7], h4.w, l4.w);
                        *reinterpret_cast<uint4*>(srow + 2 * SARR) = h4; *reinterpret_cast<uint4*>(srow + 3 * SARR) = l4;
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = acc_f(vr[i]);
                        RANGE_CHECK(f[0], f[1]); RANGE_CHECK(f[2], f[3]); RANGE_CHECK(f[4], f[5]); RANGE_CHECK(f[6], f[7]);
                        split2(f[0], f[1], h4.x, l4.x); split2(f[2], f[3], h4.y, l4.y); split2(f[4], f[5], h4.z, l4.z); split2(f[6], f[7], h4.w, l4.w);
                        *reinterpret_cast<uint4*>(srow + 4 * SARR) = h4; *reinterpret_cast<uint4*>(srow + 5 * SARR) = l4;
                    }
                    if (hq == 1) TRACE_E(21);
                    epi_bar();
                    if (hq == 1) TRACE_E(22);
                    if (worker) {
                        float kvt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                        float ksm[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            // A = V^T: tile tl -> tokens ks*16 + 8*(tl>>1) .., value columns 8*(tl&1) ..
                            const int rowA = min(p * TOK + ks * 16 + r8 + 8 * (tl >> 1), TILE_M - 1);
                            const uint32_t aV = stage_u32 + 4 * SARR + rowA * SROW + (hl * 16 + 8 * (tl & 1)) * 2;
                            uint32_t vh[4], vl[4], one[4] = {0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u};
                            ldsm_x4_t(aV, vh); ldsm_x4_t(aV + SARR, vl);
                            if (ks == 1) {                       // tokens 16 + 2 t4, 17 + 2 t4 in a0 / a1; 24.. in a2 / a3: keep < 21
                                const uint32_t mk = t4 < 2 ? 0xffffffffu : (t4 == 2 ? 0x0000ffffu : 0u);
                                vh[0] &= mk; vh[1] &= mk; vl[0] &= mk; vl[1] &= mk; one[0] &= mk; one[1] &= mk;
                                vh[2] = vh[3] = vl[2] = vl[3] = one[2] = one[3] = 0u;
                            }
                            // B = K': tile tl -> tokens ks*16 + 8*(tl&1) .., d columns 8*(tl>>1) ..
                            const int rowB = min(p * TOK + ks * 16 + r8 + 8 * (tl & 1), TILE_M - 1);
                            const uint32_t aK = stage_u32 + 2 * SARR + rowB * SROW + (hl * 16 + 8 * (tl >> 1)) * 2;
                            uint32_t kh[4], kl[4];
                            ldsm_x4_t(aK, kh); ldsm_x4_t(aK + SARR, kl);
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
                                mma16816(kvt[nt], vl, kh[2 * nt], kh[2 * nt + 1]);
                                mma16816(kvt[nt], vh, kl[2 * nt], kl[2 * nt + 1]);
                                mma16816(kvt[nt], vh, kh[2 * nt], kh[2 * nt + 1]);
                                mma16816(ksm[nt], one, kl[2 * nt], kl[2 * nt + 1]);
                                mma16816(ksm[nt], one, kh[2 * nt], kh[2 * nt + 1]);
                            }
                        }
                        if (hq == 1) TRACE_E(23);
                        // KV^T accumulators -> B fragments of Q' KV: n tile nv holds v = 8 nv + g8, b0 = d 0..7, b1 = d 8..15
                        uint32_t bh[2][2], bl[2][2], sh[2], sl[2];
                        RANGE_CHECK(kvt[0][0], kvt[0][1]); RANGE_CHECK(kvt[1][0], kvt[1][1]); RANGE_CHECK(kvt[0][2], kvt[0][3]); RANGE_CHECK(kvt[1][2], kvt[1][3]);
                        RANGE_CHECK(ksm[0][0], ksm[0][1]); RANGE_CHECK(ksm[1][0], ksm[1][1]);
                        split2(kvt[0][0], kvt[0][1], bh[0][0], bl[0][0]); split2(kvt[1][0], kvt[1][1], bh[0][1], bl[0][1]);
                        split2(kvt[0][2], kvt[0][3], bh[1][0], bl[1][0]); split2(kvt[1][2], kvt[1][3], bh[1][1], bl[1][1]);
                        split2(ksm[0][0], ksm[0][1], sh[0], sl[0]); split2(ksm[1][0], ksm[1][1], sh[1], sl[1]);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            // A = Q': tile tl -> tokens mt*16 + 8*(tl&1) .., d columns 8*(tl>>1) ..
                            const int rowQ = min(p * TOK + mt * 16 + r8 + 8 * (tl & 1), TILE_M - 1);
                            const uint32_t aQ = stage_u32 + rowQ * SROW + (hl * 16 + 8 * (tl >> 1)) * 2;
                            uint32_t qh[4], ql[4];
                            ldsm_x4(aQ, qh); ldsm_x4(aQ + SARR, ql);
                            float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, dn[4] = {0.f, 0.f, 0.f, 0.f};
                            mma16816(o0, ql, bh[0][0], bh[0][1]); mma16816(o0, qh, bl[0][0], bl[0][1]); mma16816(o0, qh, bh[0][0], bh[0][1]);
                            mma16816(o1, ql, bh[1][0], bh[1][1]); mma16816(o1, qh, bl[1][0], bl[1][1]); mma16816(o1, qh, bh[1][0], bh[1][1]);
                            mma16816(dn, ql, sh[0], sh[1]); mma16816(dn, qh, sl[0], sl[1]); mma16816(dn, qh, sh[0], sh[1]);
                            const float z0 = __fdividef(1.0f, dn[0] + P.attn_eps), z1 = __fdividef(1.0f, dn[2] + P.attn_eps);
                            // message block -> A operand (hi / lo, swizzled), token rows l0 = 16 mt + g8 and l0 + 8, columns of head
                            // 2 hq + hl: two values (v = 8 nv + 2 t4, + 1) per 4-byte store; the 8 rows x 4 lanes of a warp hit 32 banks
                            const int l0 = mt * 16 + g8;
                            const int colb = (2 * hq + hl) * 16 + 2 * t4;
                            if (l0 < TOK) {
                                const int rr = p * TOK + l0;
                                uint32_t h0, l0w, h1, l1w;
                                RANGE_CHECK(o0[0] * z0, o0[1] * z0); RANGE_CHECK(o1[0] * z0, o1[1] * z0);
                                split2(o0[0] * z0, o0[1] * z0, h0, l0w); split2(o1[0] * z0, o1[1] * z0, h1, l1w);
                                const uint32_t off0 = (uint32_t)(colb >> 6) * SLAB + sw128(rr, (colb & 63) >> 3) + (uint32_t)(colb & 7) * 2u;
                                const uint32_t off1 = (uint32_t)(colb >> 6) * SLAB + sw128(rr, ((colb + 8) & 63) >> 3) + (uint32_t)(colb & 7) * 2u;
                                *reinterpret_cast<uint32_t*>(a_base + off0) = h0; *reinterpret_cast<uint32_t*>(a_base + 2 * SLAB + off0) = l0w;
                                *reinterpret_cast<uint32_t*>(a_base + off1) = h1; *reinterpret_cast<uint32_t*>(a_base + 2 * SLAB + off1) = l1w;
                            }
                            if (l0 + 8 < TOK) {
                                const int rr = p * TOK + l0 + 8;
                                uint32_t h0, l0w, h1, l1w;
                                RANGE_CHECK(o0[2] * z1, o0[3] * z1); RANGE_CHECK(o1[2] * z1, o1[3] * z1);
                                split2(o0[2] * z1, o0[3] * z1, h0, l0w); split2(o1[2] * z1, o1[3] * z1, h1, l1w);
                                const uint32_t off0 = (uint32_t)(colb >> 6) * SLAB + sw128(rr, (colb & 63) >> 3) + (uint32_t)(colb & 7) * 2u;
                                const uint32_t off1 = (uint32_t)(colb >> 6) * SLAB + sw128(rr, ((colb + 8) & 63) >> 3) + (uint32_t)(colb & 7) * 2u;
                                *reinterpret_cast<uint32_t*>(a_base + off0) = h0; *reinterpret_cast<uint32_t*>(a_base + 2 * SLAB + off0) = l0w;
                                *reinterpret_cast<uint32_t*>(a_base + off1) = h1; *reinterpret_cast<uint32_t*>(a_base + 2 * SLAB + off1) = l1w;
                            }
                        }
                    }
                    if (hq == 0) TRACE_E(3);
                    if (hq == 1) TRACE_E(24);
                    fence_async_smem();                  // this quarter's message columns -> visible to the MMA (async proxy)
                    epi_bar();                           // (2) done reading: the next quarter may overwrite the staging
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(m_ready(hq));   // (one lane per warp) the merge MMAs of these two k steps may go
                    if (hq == 1) TRACE_E(25);
                }
            }
            TRACE_E(4);
            // ---- x back into X (the staging is dead: barrier above) for its half of mlp.0; L2 hit, runs under the merge MMAs
            load_tile_to_oper(x_base, xg, valid_rows, ew, lane, bad);
            publish(x_ready);

            // ---- E2: m = LayerNorm1(merge(msg)) -> A ----------------------------------------
            mbar_wait(d_ready, dph & 1); ++dph;          // merge done (D3)
            tc_fence_after();
            TRACE_E(5);
            {
                uint32_t r[32];
                float v[32];
                TMEM_LD32(trow + D3 + col0, r);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = acc_f(r[i]);
                float mean, rstd;
                row_stats(v, row, colq, red, P.ln_eps, mean, rstd);
                epi_bar();                               // every thread has read the partials: m may overwrite them
                float rbad = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float4 g = __ldg(reinterpret_cast<const float4*>(A.ln1_g + col0) + 2 * j + e);
                        const float4 b = __ldg(reinterpret_cast<const float4*>(A.ln1_b + col0) + 2 * j + e);
                        y[4 * e + 0] = (v[8 * j + 4 * e + 0] - mean) * rstd * g.x + b.x;
                        y[4 * e + 1] = (v[8 * j + 4 * e + 1] - mean) * rstd * g.y + b.y;
                        y[4 * e + 2] = (v[8 * j + 4 * e + 2] - mean) * rstd * g.z + b.z;
                        y[4 * e + 3] = (v[8 * j + 4 * e + 3] - mean) * rstd * g.w + b.w;
                    }
                    split_store8(a_base, row, col0 + j * 8, make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), rbad);
                }
                // rows beyond the tile's points (absent points of a ragged tile, the two pad rows) carry whatever the shared
                // memory held: they never reach an output row and must not raise the range flag
                if (row < valid_rows) bad = fmaxf(bad, rbad);
            }
            publish(a_ready);
            TRACE_E(6);

            // ---- E3: relu(h[:, :C]) -> X (x is dead: its MMAs precede the m half in issue order), then relu(h[:, C:]) -> A (m dead)
            float hbad = 0.f;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                // hf = 0: h[:, :C] in D1 (the second block still reads m from A); hf = 1: h[:, C:] in D2
                if (hf == 0) { mbar_wait(h_ready, hph & 1); ++hph; } else { mbar_wait(d_ready, dph & 1); ++dph; }
                tc_fence_after();
                if (hf == 0) TRACE_E(7);
                uint32_t hr[32];
                TMEM_LD32(trow + (hf ? D2 : D1) + col0, hr);
                tmem_wait_ld();
                uint8_t* dst = hf ? a_base : x_base;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    split_store8(dst, row, col0 + j * 8,
                                 make_float4(fmaxf(acc_f(hr[8 * j]), 0.f), fmaxf(acc_f(hr[8 * j + 1]), 0.f),
                                             fmaxf(acc_f(hr[8 * j + 2]), 0.f), fmaxf(acc_f(hr[8 * j + 3]), 0.f)),
                                 make_float4(fmaxf(acc_f(hr[8 * j + 4]), 0.f), fmaxf(acc_f(hr[8 * j + 5]), 0.f),
                                             fmaxf(acc_f(hr[8 * j + 6]), 0.f), fmaxf(acc_f(hr[8 * j + 7]), 0.f)), hbad);
                if (hf == 0) publish(x_ready);           // the first K half of mlp.2 goes while the second half is split
            }
            if (row < valid_rows) bad = fmaxf(bad, hbad);
            publish(a_ready);
            TRACE_E(8);

            // ---- E4: out = x + LayerNorm2(mlp(...)) -----------------------------------------
            // what this thread does next: the following call of this tile, or call 0 of the CTA's next tile
            const bool last_call = ci + 1 == P.n_calls;
            const int n_tile = last_call ? tile + (int)gridDim.x : tile;
            const CallArgs& NX = P.call[last_call ? 0 : ci + 1];
            const bool has_next = n_tile < P.n_tiles;
            const bool next_handoff = !last_call && !NX.is_self && NX.src == A.out;      // its source = the rows written below
            mbar_wait(d_ready, dph & 1); ++dph;
            tc_fence_after();
            TRACE_E(9);
            // X is dead (mlp.2 complete).  Unless it is about to receive the next call's source, start fetching the next call's x
            // into it now: rows this call does not write (the other map, or the next tile), ahead of this call's store burst
            pre_valid = has_next && !next_handoff && NX.x != A.out;
            if (pre_valid)
                prefetch_tile_async(x_base, NX.x + (size_t)n_tile * ROWS * CH, min(PTS, P.N - n_tile * PTS) * TOK, ew, lane);
            {
                uint32_t r[32];
                float v[32];
                TMEM_LD32(trow + D0 + col0, r);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = acc_f(r[i]);
                float mean, rstd;
                row_stats(v, row, colq, red, P.ln_eps, mean, rstd);
                epi_bar();                               // every thread has read the partials: the staging below may overwrite them
                // stage the normalised row (float4 index XOR (row & 31)) so the global traffic is coalesced (writing each thread's
                // 128 contiguous bytes straight to global memory was measured: 6.7 us instead of 3.3 for this phase)
                float* stage_f = reinterpret_cast<float*>(a_base);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(A.ln2_g + col0) + j);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(A.ln2_b + col0) + j);
                    float4 y;
                    y.x = (v[4 * j + 0] - mean) * rstd * g.x + b.x;
                    y.y = (v[4 * j + 1] - mean) * rstd * g.y + b.y;
                    y.z = (v[4 * j + 2] - mean) * rstd * g.z + b.z;
                    y.w = (v[4 * j + 3] - mean) * rstd * g.w + b.w;
                    reinterpret_cast<float4*>(stage_f + row * CH)[(colq * 8 + j) ^ (row & 31)] = y;
                }
            }
            epi_bar();
            {
                const float* stage_f = reinterpret_cast<const float*>(a_base);
                float* og = A.out + g_row0 * CH;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r2 = ew * 8 + i;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r2 < valid_rows) {
                        const float4 y = reinterpret_cast<const float4*>(stage_f + r2 * CH)[lane ^ (r2 & 31)];
                        const float4 xv = __ldcg(reinterpret_cast<const float4*>(xg + (size_t)r2 * CH) + lane);
                        o = make_float4(xv.x + y.x, xv.y + y.y, xv.z + y.z, xv.w + y.w);
                        __stcg(reinterpret_cast<float4*>(og + (size_t)r2 * CH) + lane, o);
                    }
                    if (next_handoff) {
                        // these rows are the next call's source: split them into X right here (rows that do not exist: zeros)
                        const int col = lane * 4;
                        const uint32_t off = (uint32_t)(col >> 6) * SLAB + sw128(r2, (col & 63) >> 3) + (uint32_t)(col & 7) * 2u;
                        RANGE_CHECK(o.x, o.y); RANGE_CHECK(o.z, o.w);
                        uint2 hi, lo;
                        split2(o.x, o.y, hi.x, lo.x); split2(o.z, o.w, hi.y, lo.y);
                        *reinterpret_cast<uint2*>(x_base + off) = hi;
                        *reinterpret_cast<uint2*>(x_base + 2 * SLAB + off) = lo;
                    }
                }
                epi_bar();                               // staging reads done (and this call's rows written) before the next call
            }
            TRACE_E(10);
            tc_fence_before();
        }
        if (!(bad <= 65504.0f) && P.overflow) atomicOr(P.overflow, 1);
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------ weight packing
// block b (issue order) -> (matrix, first row, first column, leading dimension); chunk = block * 2 + K half
struct PackArgs {
    const float* wq; const float* wk; const float* wv; const float* wm; const float* w1; const float* w2;
    uint8_t* dst;
};

__global__ void __launch_bounds__(256)
loftr_pack_weights_kernel(const PackArgs P) {
    const int t = blockIdx.x * 256 + threadIdx.x;                // one 16-byte segment (8 fp16) each
    if (t >= NCHUNK * 128 * 8) return;
    const int seg = t & 7, row = (t >> 3) & 127, c = t >> 10, blk = c >> 1, kb = c & 1;
    const float* W; int r0 = 0, c0 = 0, ld = CH;
    switch (blk) {
        case 0: W = P.wk; break;                                 // k, v before q: after a hand-off their operand is ready first
        case 1: W = P.wv; break;
        case 2: W = P.wq; break;
        case 3: W = P.wm; break;
        case 4: W = P.w1; ld = 2 * CH; break;                    // mlp.0[:C, :C]   (x half of cat([x, message]))
        case 5: W = P.w1; ld = 2 * CH; r0 = CH; break;           // mlp.0[C:, :C]
        case 6: W = P.w1; ld = 2 * CH; c0 = CH; break;           // mlp.0[:C, C:]   (message half)
        case 7: W = P.w1; ld = 2 * CH; r0 = CH; c0 = CH; break;  // mlp.0[C:, C:]
        case 8: W = P.w2; ld = 2 * CH; break;                    // mlp.2[:, :C]
        default: W = P.w2; ld = 2 * CH; c0 = CH; break;          // mlp.2[:, C:]
    }
    const float4* src = reinterpret_cast<const float4*>(W + (size_t)(r0 + row) * ld + c0 + kb * 64 + seg * 8);
    float4 a = __ldg(src), b = __ldg(src + 1);
    a.x *= W_SCALE; a.y *= W_SCALE; a.z *= W_SCALE; a.w *= W_SCALE;
    b.x *= W_SCALE; b.y *= W_SCALE; b.z *= W_SCALE; b.w *= W_SCALE;
    uint4 hi, lo;
    split2(a.x, a.y, hi.x, lo.x); split2(a.z, a.w, hi.y, lo.y);
    split2(b.x, b.y, hi.z, lo.z); split2(b.z, b.w, hi.w, lo.w);
    uint8_t* d = P.dst + (size_t)c * CHUNK + sw128(row, seg);
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + SLAB) = lo;
}

// ------------------------------------------------------------------------------------ self-test GEMM
// D (128 x 128) = A (128 x K) * B (128 x K)^T, K in {64, 128}, through exactly the staging /
// descriptor / TMEM helpers of the layer kernel.  terms = 1: plain fp16; terms = 3: hi/lo split.
__global__ void __launch_bounds__(128, 1)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int terms) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + OPER_BYTES;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, row = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float bad = 0.f;
    for (int c = 0; c < K; c += 8) {
        const float4* pa = reinterpret_cast<const float4*>(A + (size_t)row * K + c);
        const float4* pb = reinterpret_cast<const float4*>(B + (size_t)row * K + c);
        split_store8(a_base, row, c, pa[0], pa[1], bad);
        split_store8(b_base, row, c, pb[0], pb[1], bad);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a_hi = smem_u32(a_base), a_lo = a_hi + 2 * SLAB, b_hi = smem_u32(b_base), b_lo = b_hi + 2 * SLAB;
        constexpr uint32_t IDESC = idesc_f16(128);
        uint32_t acc = 0;
        for (int kb = 0; kb < K / 64; ++kb)
            for (int k16 = 0; k16 < 4; ++k16) {
                const uint64_t adv = (uint64_t)(k16 * 2);
                if (terms == 3) {
                    umma_f16(tmem, smem_desc(a_lo + kb * SLAB) + adv, smem_desc(b_hi + kb * SLAB) + adv, IDESC, acc); acc = 1;
                    umma_f16(tmem, smem_desc(a_hi + kb * SLAB) + adv, smem_desc(b_lo + kb * SLAB) + adv, IDESC, 1);
                }
                umma_f16(tmem, smem_desc(a_hi + kb * SLAB) + adv, smem_desc(b_hi + kb * SLAB) + adv, IDESC, acc); acc = 1;
            }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int part = 0; part < 4; ++part) {
        uint32_t r[32];
        TMEM_LD32(trow + part * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) D[(size_t)row * 128 + part * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// ------------------------------------------------------------------------------------ MMA rate probe
// Times `count` back-to-back tcgen05.mma (M = 128, K = 32 bytes) issued by one thread against zeroed shared-memory
// operands: mode bits 0..1 = N (0: 64, 1: 128, 2: 256), bits 4..7 = number of accumulators rotated over (1..4),
// bit 8 = kind::tf32 instead of kind::f16.  cycles[blockIdx.x] = clock64 delta from first issue to commit arrival.
__global__ void __launch_bounds__(128, 1)
tc_mma_rate_kernel(int mode, int count, long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 6 * SLAB / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t n = 64u << (mode & 3), nacc = (uint32_t)((mode >> 4) & 15), tf32 = (mode >> 8) & 1;
    long long t0 = 0;
    if (warp == 0) {
        const uint32_t a = smem_u32(smem), b = a + 2 * SLAB;
        const uint32_t idesc = tf32 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24)) : idesc_f16(n);
        uint32_t dd[8];
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            dd[u] = tmem + ((uint32_t)u % nacc) * n;
            ad[u] = smem_desc(a + (u & 1) * SLAB) + (uint64_t)((u >> 1) & 3) * 2;
            bd[u] = smem_desc(b + (u & 1) * SLAB) + (uint64_t)((u >> 1) & 3) * 2;
        }
        t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < count; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (tf32)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(dd[u]), "l"(ad[u]), "l"(bd[u]), "r"(idesc), "r"(1u) : "memory");
                    else
                        umma_f16(dd[u], ad[u], bd[u], idesc, 1);
                }
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    tc_fence_after();
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace tc

// ------------------------------------------------------------------------------------ host side
static long long* g_loftr_trace = nullptr;     // device buffer of 256 int64, set by riders_debug_loftr_trace
bool loftr_tc_built() { return true; }

size_t loftr_tc_workspace_bytes(int N, int tokens, int C) {
    // packed weights of up to MAX_LAYERS layers (the tokens live in the caller's output buffers)
    (void)N; (void)tokens; (void)C;
    return (size_t)tc::MAX_LAYERS * tc::LAYER_PACK_BYTES + 1024;
}

static uint8_t* pack_base(void* tc_ws) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_ws) + 255) & ~(uintptr_t)255);
}

int loftr_tc_pack(const riders_loftr_weights& w, int layer_slot, void* tc_ws, cudaStream_t st) {
    RIDERS_REQUIRE(layer_slot >= 0 && layer_slot < tc::MAX_LAYERS, "loftr(tcgen05): at most %d layers per call", tc::MAX_LAYERS);
    tc::PackArgs P;
    P.wq = w.wq; P.wk = w.wk; P.wv = w.wv; P.wm = w.wm; P.w1 = w.w1; P.w2 = w.w2;
    P.dst = pack_base(tc_ws) + (size_t)layer_slot * tc::LAYER_PACK_BYTES;
    count_launches(1);
    tc::loftr_pack_weights_kernel<<<ceil_div(tc::NCHUNK * 128 * 8, 256), 256, 0, st>>>(P);
    return check_launch("loftr_pack_weights");
}

// One launch for a list of layer calls (all of a transformer, or a single layer): call i reads x[i] / src[i], writes out[i]
// (out[i] may alias x[i]) with the packed weights of layer slot[i].  `overflow` (device int, zeroed by the caller) is set when an
// operand left the fp16 range.
int loftr_tc_run_calls(int n_calls, const float* const* x, const float* const* src, float* const* out,
                       const riders_loftr_weights* const* w, const int* slot,
                       int N, float attn_eps, float ln_eps, void* tc_ws, int* overflow, cudaStream_t st) {
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(tc::loftr_layer_tc_kernel, (size_t)(tc::SMEM_BYTES), configured, "loftr(tcgen05): shared memory attribute")) return rc_smem;
    RIDERS_REQUIRE(n_calls >= 1 && n_calls <= tc::MAX_CALLS, "loftr(tcgen05): 1..%d layer calls per launch (got %d)", tc::MAX_CALLS, n_calls);
    tc::LayerArgs P;
    for (int i = 0; i < n_calls; ++i) {
        RIDERS_REQUIRE(out[i] != src[i] || x[i] == src[i], "loftr(tcgen05): out must not alias the source of a cross layer");
        tc::CallArgs& A = P.call[i];
        A.x = x[i]; A.src = src[i]; A.out = out[i];
        A.wpack = pack_base(tc_ws) + (size_t)slot[i] * tc::LAYER_PACK_BYTES;
        A.ln1_g = w[i]->ln1_g; A.ln1_b = w[i]->ln1_b; A.ln2_g = w[i]->ln2_g; A.ln2_b = w[i]->ln2_b;
        A.is_self = (x[i] == src[i]) ? 1 : 0;
    }
    P.n_calls = n_calls;
    P.N = N; P.n_tiles = ceil_div(N, tc::PTS);
    P.attn_eps = attn_eps; P.ln_eps = ln_eps;
    P.trace = g_loftr_trace;
    P.overflow = overflow;
    const int grid = P.n_tiles < sm_count() ? P.n_tiles : sm_count();
    count_launches(1);
    tc::loftr_layer_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(P);
    return check_launch("loftr_layer_tc");
}

}  // namespace riders

extern "C" int riders_debug_tc_gemm(const float* a, const float* b, float* d, int K, int terms, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(a && b && d, "debug_tc_gemm: null pointer");
    RIDERS_REQUIRE((K == 64 || K == 128) && (terms == 1 || terms == 3), "debug_tc_gemm: K=%d terms=%d unsupported", K, terms);
    const int smem = 1024 + 2 * tc::OPER_BYTES;
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(tc::tc_gemm_selftest_kernel, (size_t)(smem), configured, "debug_tc_gemm: shared memory attribute")) return rc_smem;
    count_launches(1);
    tc::tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, K, terms);
    return check_launch("debug_tc_gemm");
}

extern "C" int riders_debug_loftr_trace(long long* device_buffer_256) {
    riders::g_loftr_trace = device_buffer_256;
    return RIDERS_OK;
}

extern "C" int riders_debug_tc_mma_rate(int mode, int count, int ctas, long long* device_cycles, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(device_cycles && count > 0 && ctas > 0, "debug_tc_mma_rate: bad arguments");
    const int nacc = (mode >> 4) & 15, n = 64 << (mode & 3);
    RIDERS_REQUIRE((mode & 3) < 3 && nacc >= 1 && nacc * n <= 512, "debug_tc_mma_rate: mode %d unsupported", mode);
    const int smem = 1024 + 6 * tc::SLAB;
    static size_t configured[64] = {0};
    if (int rc_smem = ensure_dynamic_smem(tc::tc_mma_rate_kernel, (size_t)(smem), configured, "debug_tc_mma_rate: shared memory attribute")) return rc_smem;
    count_launches(1);
    tc::tc_mma_rate_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(mode, count, device_cycles);
    return check_launch("debug_tc_mma_rate");
}
