// A3 fused LoFTREncoderLayer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
// Reference semantics: RCNet/linear_attention.py:112-135 (layer) and :18-45 (linear attention).
//
// Specialised for RC-Net's transformer: d_model 128, 8 heads of 16, 21 tokens per radar point
// (networks.py:378, :441).  One CTA per SM, persistent over tiles of 6 points (126 token rows
// padded to the UMMA M of 128).  Every dense contraction of the layer - q/k/v projections,
// merge, the 256->256->128 MLP - is a tcgen05.mma kind::tf32 with both operands in shared
// memory (K-major, 128-byte swizzle) and the fp32 accumulator in tensor memory.  fp32 accuracy
// comes from the 3-term split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  with hi = the top 19
// bits (exactly representable in tf32) and lo = a - hi, all three accumulated in TMEM.
//
// Roles (192 threads): warp 0 streams the pre-split, pre-swizzled weight chunks (32 KB: hi|lo
// of a 128 x 32 block) from L2 with cp.async.bulk through a 3-stage mbarrier ring; lane 0 of
// warp 1 issues the MMAs and commits their completion to mbarriers; warps 2..5 own one TMEM lane
// quadrant each (thread = token row) and run every epilogue: elu+1, the per-(point, head) KV /
// normaliser contraction (2.5 % of the layer's FLOPs, CUDA cores), LayerNorms, ReLU, residual,
// and the hi/lo re-split of the next A operand straight into the swizzled layout.
//
// TMEM columns (512 allocated): D0 [0,128) q -> merge -> mlp.2 | D1 [128,256) k -> msg scratch
//                               D2 [256,384) v | D3 [256,512) mlp.0 (after v is consumed)
// Shared memory: A operand hi[4]|lo[4] slabs of 128 x 128 B (128 KB, also the attention / output
// staging) + 3 weight stages of 32 KB.
#include "common.cuh"

namespace riders {
namespace tc {

constexpr int TOK = 21, CH = 128, NHEAD = 8, HD = 16;
constexpr int TILE_M = 128, PTS = 6, ROWS = PTS * TOK;        // 126 valid rows per tile
constexpr int SLAB = 128 * 128;                                // bytes: 128 rows x 32 fp32
constexpr int A_BYTES = 8 * SLAB;
constexpr int CHUNK = 2 * SLAB;                                // hi + lo of one 128 x 32 weight block
constexpr int NSTAGE = 3;
constexpr int NCHUNK = 40;
constexpr int LAYER_PACK_BYTES = NCHUNK * CHUNK;               // 1.25 MB per layer
constexpr int MAX_LAYERS = 16;
constexpr int THREADS = 320;                                   // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int TAIL_BYTES = 128 + 2048;                         // barriers + TMEM slot | LayerNorm partials (4 x 128 floats)
constexpr int SMEM_BYTES = A_BYTES + NSTAGE * CHUNK + TAIL_BYTES;      // 231552 of the 232448 available; base must be 1024-aligned
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t D0 = 0, D1 = 128, D2 = 256, D3 = 256;
// instruction descriptor: D fp32, A/B tf32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

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

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte-swizzled operand slab (rows x 32 fp32): 8-row groups 1024 B apart
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
#define TMEM_ST16(taddr, r)                                                                                     \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                    \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),     \
                   "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), \
                   "r"(r[15]) : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// ------------------------------------------------------------------------------------ operand staging
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// byte offset of 16-byte segment `seg` (0..7) of row `row` inside one swizzled slab
__device__ __forceinline__ uint32_t sw128(int row, int seg) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((seg ^ (row & 7)) << 4));
}
// split 4 consecutive K elements (col % 4 == 0) of row `row` into the hi / lo A slabs
__device__ __forceinline__ void split_store4(uint8_t* a_base, int row, int col, float4 v) {
    const uint32_t off = (uint32_t)(col >> 5) * SLAB + sw128(row, (col & 31) >> 2);
    float4 hi, lo;
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
    *reinterpret_cast<float4*>(a_base + off) = hi;
    *reinterpret_cast<float4*>(a_base + 4 * SLAB + off) = lo;
}

// coalesced: epilogue warp `ew` (0..7) loads rows ew*16.. of a (rows x 128) fp32 tile and splits them into A
__device__ __forceinline__ void load_tile_to_A(uint8_t* a_base, const float* __restrict__ g, int valid_rows, int ew, int lane) {
#pragma unroll 1
    for (int r0 = 0; r0 < 16; r0 += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = ew * 16 + r0 + i;
            v[i] = (row < valid_rows) ? __ldg(reinterpret_cast<const float4*>(g + (size_t)row * CH) + lane)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) split_store4(a_base, ew * 16 + r0 + i, lane * 4, v[i]);
    }
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.0f : __expf(x); }   // elu(x) + 1 (SFU exp2, ~2 ulp)

// LayerNorm statistics of a 128-wide row held as 64 + 64 values by the two threads (column halves) of the row:
// partial sums go through shared memory and are added in a fixed order, so both threads get identical bits.
__device__ __forceinline__ void row_stats(const uint32_t (&v)[64], int row, int half, float* red, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += __uint_as_float(v[i]);
    red[half * 128 + row] = s;
    epi_bar();
    mean = (red[row] + red[128 + row]) * (1.0f / CH);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) { const float d = __uint_as_float(v[i]) - mean; q = fmaf(d, d, q); }
    red[256 + half * 128 + row] = q;
    epi_bar();
    rstd = 1.0f / sqrtf((red[256 + row] + red[384 + row]) * (1.0f / CH) + eps);
}

// signal "A operand (and my TMEM reads) done" to the MMA warp
__device__ __forceinline__ void publish_A(uint32_t a_ready) {
    fence_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);
}

struct LayerArgs {
    const float* x; const float* src; float* out;
    const uint8_t* wpack;                       // NCHUNK x 32 KB, pre-split + pre-swizzled
    const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;
    int N;                                      // radar points
    int n_tiles;
    int is_self;                                // src == x
    float attn_eps, ln_eps;
    long long* trace;                           // optional phase timestamps (riders_debug_loftr_trace), else null
};

// trace layout per traced CTA (blockIdx.x == 0 or gridDim.x / 2): 128 slots; [0..63] epilogue thread, [64..127] MMA thread
#define TRACE_E(k) do { if (P.trace && tr_on && lane == 0 && warp == 4 && tile == (int)blockIdx.x) P.trace[tr_base + (k)] = clock64(); } while (0)
#define TRACE_M(k) do { if (P.trace && tr_on && tile == (int)blockIdx.x) P.trace[tr_base + 64 + (k)] = clock64(); } while (0)

// ------------------------------------------------------------------------------------ the fused layer
__global__ void __launch_bounds__(THREADS, 1)
loftr_layer_tc_kernel(const LayerArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;                                 // SWIZZLE_128B slabs need a 1024-byte aligned base
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* a_base = smem;                                   // A hi[4] | lo[4]
    uint8_t* b_base = smem + A_BYTES;                         // NSTAGE x (hi | lo)
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + NSTAGE * CHUNK);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);
    const uint32_t bfull = smem_u32(bars), bempty = smem_u32(bars + NSTAGE);
    const uint32_t a_ready = smem_u32(bars + 2 * NSTAGE), d_ready = smem_u32(bars + 2 * NSTAGE + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr_on = blockIdx.x == 0 || blockIdx.x == gridDim.x / 2;
    const int tr_base = (blockIdx.x == 0 ? 0 : 128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
        mbar_init(a_ready, 256);
        mbar_init(d_ready, 1);
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
                for (int c = 0; c < NCHUNK; ++c, ++n) {
                    const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1;
                    mbar_wait(bempty + 8 * s, ph ^ 1);
                    mbar_expect_tx(bfull + 8 * s, CHUNK);
                    bulk_g2s(smem_u32(b_base + s * CHUNK), P.wpack + (size_t)c * CHUNK, CHUNK, bfull + 8 * s);
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t n = 0, aph = 0;
            const uint32_t a_hi = smem_u32(a_base), a_lo = a_hi + 4 * SLAB, b0 = smem_u32(b_base);
            // segment ends (exclusive chunk index): after each the accumulators are committed to d_ready
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                int c = 0;
#pragma unroll 1
                for (int seg = 0; seg < 7; ++seg) {
                    int c_end;
                    switch (seg) {
                        case 0: c_end = P.is_self ? 12 : 4; break;      // A = x   : q (+ k, v when self)
                        case 1: c_end = 12; break;                       // A = src : k, v (cross only)
                        case 2: c_end = 16; break;                       // A = msg : merge
                        case 3: c_end = 24; break;                       // A = m   : mlp.0[:, C:]
                        case 4: c_end = 32; break;                       // A = x   : mlp.0[:, :C]
                        case 5: c_end = 36; break;                       // A = relu(h)[:, :C]  : mlp.2
                        default: c_end = 40; break;                      // A = relu(h)[:, C:]  : mlp.2
                    }
                    if (c == c_end) continue;                            // empty segment (self layer, seg 1)
                    TRACE_M(seg * 4 + 0);
                    mbar_wait(a_ready, aph & 1); ++aph;
                    tc_fence_after();
                    TRACE_M(seg * 4 + 1);
                    long long wsum = 0;
                    for (; c < c_end; ++c, ++n) {
                        const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1;
                        const long long tw = P.trace ? clock64() : 0;
                        mbar_wait(bfull + 8 * s, ph);
                        if (P.trace) wsum += clock64() - tw;
                        tc_fence_after();
                        const int kc = c & 3;
                        uint32_t dcol, acc0;
                        if (c < 12)      { dcol = (uint32_t)(c >> 2) * 128; acc0 = kc > 0; }
                        else if (c < 16) { dcol = D0; acc0 = kc > 0; }
                        else if (c < 24) { dcol = D3 + (uint32_t)((c - 16) >> 2) * 128; acc0 = kc > 0; }
                        else if (c < 32) { dcol = D3 + (uint32_t)((c - 24) >> 2) * 128; acc0 = 1; }
                        else if (c < 36) { dcol = D0; acc0 = kc > 0; }
                        else             { dcol = D0; acc0 = 1; }
                        const uint64_t ah = smem_desc(a_hi + kc * SLAB), al = smem_desc(a_lo + kc * SLAB);
                        const uint64_t bh = smem_desc(b0 + s * CHUNK), bl = smem_desc(b0 + s * CHUNK + SLAB);
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {                 // 8 tf32 = 32 bytes per MMA along K
                            const uint64_t adv = (uint64_t)(k8 * 2);
                            umma_tf32(tmem + dcol, al + adv, bh + adv, acc0 | (k8 > 0));
                            umma_tf32(tmem + dcol, ah + adv, bl + adv, 1);
                            umma_tf32(tmem + dcol, ah + adv, bh + adv, 1);
                        }
                        umma_commit(bempty + 8 * s);                     // stage reusable once these MMAs retire
                    }
                    umma_commit(d_ready);
                    TRACE_M(seg * 4 + 2);
                    if (P.trace && tr_on && tile == (int)blockIdx.x) P.trace[tr_base + 64 + seg * 4 + 3] = wsum;
                }
            }
        }
    } else {
        // ===================== epilogue warps: (thread pair) = token row, each thread 64 of its 128 columns =====
        const int ew = warp - 2;                         // 0..7
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                        // column half: warps 2..5 -> [0,64), warps 6..9 -> [64,128)
        const int row = quad * 32 + lane;
        const int col0 = half * 64;
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
        uint32_t dph = 0;
        float* stage_f = reinterpret_cast<float*>(a_base);

        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
            const int pts = min(PTS, P.N - tile * PTS);
            const int valid_rows = pts * TOK;
            const size_t g_row0 = (size_t)tile * ROWS;
            const float* xg = P.x + g_row0 * CH;
            const float* sg = P.src + g_row0 * CH;

            // ---- E0: A = x ------------------------------------------------------------------
            TRACE_E(0);
            load_tile_to_A(a_base, xg, valid_rows, ew, lane);
            publish_A(a_ready);
            TRACE_E(1);
            if (!P.is_self) {
                mbar_wait(d_ready, dph & 1); ++dph;      // q done: A is free again
                tc_fence_after();
                load_tile_to_A(a_base, sg, valid_rows, ew, lane);
                publish_A(a_ready);
            }

            // ---- E1: linear attention (linear_attention.py:29-43) ---------------------------
            // The reference divides V by S and multiplies the message by S (fp16 range safety,
            // linear_attention.py:39-43); in fp32 the two cancel to rounding, so neither is applied.
            mbar_wait(d_ready, dph & 1); ++dph;          // q, k, v in D0, D1, D2
            tc_fence_after();
            TRACE_E(2);
            float msg[2][2][16];                         // [head half hh][head e][v]: heads hh*4 + half*2 + e
            {
                // staging inside the (now idle) A region, all fp32:
                //   Kst [128][64] @0, Vst [128][64] @32K (float4 index XOR (row & 15)), KVs [24][16][16] @64K, Ksum [24][16] @88K
                float* Kst = stage_f;
                float* Vst = stage_f + 8192;
                float* KVs = stage_f + 16384;
                float* Ksum = stage_f + 16384 + 6144;
                const int p_row = min(row / TOK, PTS - 1);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    // (1) K' = elu(k)+1 and V of heads 4hh..4hh+3 -> smem (this thread: 32 of the 64 columns)
                    {
                        uint32_t kr[32], vr[32];
                        TMEM_LD32(trow + D1 + hh * 64 + half * 32, kr);
                        TMEM_LD32(trow + D2 + hh * 64 + half * 32, vr);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 kk, vv;
                            kk.x = elu1(__uint_as_float(kr[4 * j + 0])); kk.y = elu1(__uint_as_float(kr[4 * j + 1]));
                            kk.z = elu1(__uint_as_float(kr[4 * j + 2])); kk.w = elu1(__uint_as_float(kr[4 * j + 3]));
                            vv.x = __uint_as_float(vr[4 * j + 0]); vv.y = __uint_as_float(vr[4 * j + 1]);
                            vv.z = __uint_as_float(vr[4 * j + 2]); vv.w = __uint_as_float(vr[4 * j + 3]);
                            const int f4 = (half * 8 + j) ^ (row & 15);
                            reinterpret_cast<float4*>(Kst + row * 64)[f4] = kk;
                            reinterpret_cast<float4*>(Vst + row * 64)[f4] = vv;
                        }
                    }
                    epi_bar();
                    // (2) KV[p][h] = K'^T V, Ksum[p][h] = sum_s K'   (24 (point, head) pairs, 3 per warp;
                    //     lane -> (d, 8 consecutive v); 7 tokens per batch so 21 loads are in flight)
#pragma unroll 1
                    for (int q3 = 0; q3 < 3; ++q3) {
                        const int pair = ew * 3 + q3, p = pair >> 2, hl = pair & 3;
                        const int d = lane >> 1, v0 = (lane & 1) * 8;
                        float acc[8], ks = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                        if (p < pts) {
#pragma unroll 1
                            for (int s0 = 0; s0 < TOK; s0 += 7) {
                                float kk[7];
                                float4 va[7], vb[7];
#pragma unroll
                                for (int u = 0; u < 7; ++u) {
                                    const int r = p * TOK + s0 + u, sw = r & 15;
                                    kk[u] = Kst[r * 64 + (((hl * 4 + (d >> 2)) ^ sw) << 2) + (d & 3)];
                                    va[u] = reinterpret_cast<const float4*>(Vst + r * 64)[(hl * 4 + (v0 >> 2)) ^ sw];
                                    vb[u] = reinterpret_cast<const float4*>(Vst + r * 64)[(hl * 4 + (v0 >> 2) + 1) ^ sw];
                                }
#pragma unroll
                                for (int u = 0; u < 7; ++u) {
                                    ks += kk[u];
                                    acc[0] = fmaf(kk[u], va[u].x, acc[0]); acc[1] = fmaf(kk[u], va[u].y, acc[1]);
                                    acc[2] = fmaf(kk[u], va[u].z, acc[2]); acc[3] = fmaf(kk[u], va[u].w, acc[3]);
                                    acc[4] = fmaf(kk[u], vb[u].x, acc[4]); acc[5] = fmaf(kk[u], vb[u].y, acc[5]);
                                    acc[6] = fmaf(kk[u], vb[u].z, acc[6]); acc[7] = fmaf(kk[u], vb[u].w, acc[7]);
                                }
                            }
                        }
                        float4* dst = reinterpret_cast<float4*>(KVs + pair * 256 + d * 16 + v0);
                        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                        if ((lane & 1) == 0) Ksum[pair * 16 + d] = ks;
                    }
                    epi_bar();
                    // (3) message rows of heads hh*4 + half*2 + {0,1} -> registers
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int hl = half * 2 + e;
                        uint32_t qr[16];
                        TMEM_LD16(trow + D0 + (hh * 4 + hl) * 16, qr);
                        tmem_wait_ld();
                        const int pair = p_row * 4 + hl;
                        float Q[16], den = 0.f;
#pragma unroll
                        for (int dd = 0; dd < 16; ++dd) { Q[dd] = elu1(__uint_as_float(qr[dd])); den = fmaf(Q[dd], Ksum[pair * 16 + dd], den); }
#pragma unroll
                        for (int v = 0; v < 16; ++v) msg[hh][e][v] = 0.f;
#pragma unroll
                        for (int dd = 0; dd < 16; ++dd) {
                            const float4* kv = reinterpret_cast<const float4*>(KVs + pair * 256 + dd * 16);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 t = kv[j];
                                msg[hh][e][4 * j + 0] = fmaf(Q[dd], t.x, msg[hh][e][4 * j + 0]); msg[hh][e][4 * j + 1] = fmaf(Q[dd], t.y, msg[hh][e][4 * j + 1]);
                                msg[hh][e][4 * j + 2] = fmaf(Q[dd], t.z, msg[hh][e][4 * j + 2]); msg[hh][e][4 * j + 3] = fmaf(Q[dd], t.w, msg[hh][e][4 * j + 3]);
                            }
                        }
                        const float z = __fdividef(1.0f, den + P.attn_eps);
#pragma unroll
                        for (int v = 0; v < 16; ++v) msg[hh][e][v] *= z;
                    }
                    // the barrier after the next half's stage (1) orders (3) before the next (2)
                }
                epi_bar();                               // staging dead: A may be overwritten
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            split_store4(a_base, row, (hh * 4 + half * 2 + e) * 16 + j * 4,
                                         make_float4(msg[hh][e][4 * j], msg[hh][e][4 * j + 1], msg[hh][e][4 * j + 2], msg[hh][e][4 * j + 3]));
            }
            publish_A(a_ready);
            TRACE_E(3);

            // ---- E2: m = LayerNorm1(merge(msg)) -> A ----------------------------------------
            mbar_wait(d_ready, dph & 1); ++dph;
            tc_fence_after();
            TRACE_E(4);
            {
                uint32_t v[64];
                TMEM_LD32(trow + D0 + col0, (v + 0)); TMEM_LD32(trow + D0 + col0 + 32, (v + 32));
                tmem_wait_ld();
                float mean, rstd;
                row_stats(v, row, half, red, P.ln_eps, mean, rstd);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(P.ln1_g + col0) + j);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(P.ln1_b + col0) + j);
                    float4 y;
                    y.x = (__uint_as_float(v[4 * j + 0]) - mean) * rstd * g.x + b.x;
                    y.y = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * g.y + b.y;
                    y.z = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * g.z + b.z;
                    y.w = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * g.w + b.w;
                    split_store4(a_base, row, col0 + j * 4, y);
                }
            }
            publish_A(a_ready);
            TRACE_E(5);

            // ---- E3: A = x again for mlp.0[:, :C] -------------------------------------------
            mbar_wait(d_ready, dph & 1); ++dph;
            tc_fence_after();
            TRACE_E(6);
            load_tile_to_A(a_base, xg, valid_rows, ew, lane);
            publish_A(a_ready);
            TRACE_E(7);

            // ---- E4 / E5: A = relu(h) halves for mlp.2 --------------------------------------
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                mbar_wait(d_ready, dph & 1); ++dph;
                tc_fence_after();
                TRACE_E(8 + 2 * hf);
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t hr[32];
                    TMEM_LD32(trow + D3 + hf * 128 + col0 + part * 32, hr);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        split_store4(a_base, row, col0 + part * 32 + j * 4,
                                     make_float4(fmaxf(__uint_as_float(hr[4 * j]), 0.f), fmaxf(__uint_as_float(hr[4 * j + 1]), 0.f),
                                                 fmaxf(__uint_as_float(hr[4 * j + 2]), 0.f), fmaxf(__uint_as_float(hr[4 * j + 3]), 0.f)));
                }
                publish_A(a_ready);
                TRACE_E(9 + 2 * hf);
            }

            // ---- E6: out = x + LayerNorm2(mlp(...)) -----------------------------------------
            mbar_wait(d_ready, dph & 1); ++dph;
            tc_fence_after();
            TRACE_E(12);
            {
                uint32_t v[64];
                TMEM_LD32(trow + D0 + col0, (v + 0)); TMEM_LD32(trow + D0 + col0 + 32, (v + 32));
                tmem_wait_ld();
                float mean, rstd;
                row_stats(v, row, half, red, P.ln_eps, mean, rstd);
                // stage the normalised row (float4 index XOR (row & 31)) so the global traffic is coalesced
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(P.ln2_g + col0) + j);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(P.ln2_b + col0) + j);
                    float4 y;
                    y.x = (__uint_as_float(v[4 * j + 0]) - mean) * rstd * g.x + b.x;
                    y.y = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * g.y + b.y;
                    y.z = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * g.z + b.z;
                    y.w = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * g.w + b.w;
                    reinterpret_cast<float4*>(stage_f + row * CH)[(half * 16 + j) ^ (row & 31)] = y;
                }
                epi_bar();
                float* og = P.out + g_row0 * CH;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = ew * 16 + i;
                    if (r < valid_rows) {
                        const float4 y = reinterpret_cast<const float4*>(stage_f + r * CH)[lane ^ (r & 31)];
                        const float4 xv = __ldg(reinterpret_cast<const float4*>(xg + (size_t)r * CH) + lane);
                        reinterpret_cast<float4*>(og + (size_t)r * CH)[lane] = make_float4(xv.x + y.x, xv.y + y.y, xv.z + y.z, xv.w + y.w);
                    }
                }
                epi_bar();                               // staging reads done before the next tile's split stores
            }
            TRACE_E(13);
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------ weight packing
// chunk c -> (matrix, first row, first column, leading dimension); see the file header for the order
struct PackArgs {
    const float* wq; const float* wk; const float* wv; const float* wm; const float* w1; const float* w2;
    uint8_t* dst;
};

__global__ void __launch_bounds__(256)
loftr_pack_weights_kernel(const PackArgs P) {
    const int t = blockIdx.x * 256 + threadIdx.x;                // one 16-byte segment each
    if (t >= NCHUNK * 128 * 8) return;
    const int seg = t & 7, row = (t >> 3) & 127, c = t >> 10, kc = c & 3;
    const float* W; int r0 = 0, c0 = kc * 32, ld = CH;
    if (c < 4) W = P.wq;
    else if (c < 8) W = P.wk;
    else if (c < 12) W = P.wv;
    else if (c < 16) W = P.wm;
    else if (c < 24) { W = P.w1; ld = 2 * CH; r0 = ((c - 16) >> 2) * 128; c0 = CH + kc * 32; }
    else if (c < 32) { W = P.w1; ld = 2 * CH; r0 = ((c - 24) >> 2) * 128; }
    else if (c < 36) { W = P.w2; ld = 2 * CH; }
    else { W = P.w2; ld = 2 * CH; c0 = CH + kc * 32; }
    const float4 v = __ldg(reinterpret_cast<const float4*>(W + (size_t)(r0 + row) * ld + c0 + seg * 4));
    float4 hi, lo;
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
    uint8_t* d = P.dst + (size_t)c * CHUNK + sw128(row, seg);
    *reinterpret_cast<float4*>(d) = hi;
    *reinterpret_cast<float4*>(d + SLAB) = lo;
}

// ------------------------------------------------------------------------------------ self-test GEMM
// D (128 x 128) = A (128 x K) * B (128 x K)^T, K in {32, 64}, through exactly the staging /
// descriptor / TMEM helpers of the layer kernel.  terms = 1: plain tf32; terms = 3: hi/lo split.
__global__ void __launch_bounds__(128, 1)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int terms) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + 2 * SLAB;                        // B: hi in slabs 2,3 / lo in slabs 6,7 of the same region
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int c = 0; c < K; c += 4) {
        split_store4(a_base, row, c, *reinterpret_cast<const float4*>(A + (size_t)row * K + c));
        split_store4(b_base, row, c, *reinterpret_cast<const float4*>(B + (size_t)row * K + c));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a_hi = smem_u32(a_base), a_lo = a_hi + 4 * SLAB, b_hi = smem_u32(b_base), b_lo = b_hi + 4 * SLAB;
        uint32_t acc = 0;
        for (int kc = 0; kc < K / 32; ++kc)
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint64_t adv = (uint64_t)(k8 * 2);
                if (terms == 3) {
                    umma_tf32(tmem, smem_desc(a_lo + kc * SLAB) + adv, smem_desc(b_hi + kc * SLAB) + adv, acc); acc = 1;
                    umma_tf32(tmem, smem_desc(a_hi + kc * SLAB) + adv, smem_desc(b_lo + kc * SLAB) + adv, 1);
                }
                umma_tf32(tmem, smem_desc(a_hi + kc * SLAB) + adv, smem_desc(b_hi + kc * SLAB) + adv, acc); acc = 1;
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
    (void)lane;
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

}  // namespace tc

// ------------------------------------------------------------------------------------ host side
static long long* g_loftr_trace = nullptr;     // device buffer of 256 int64, set by riders_debug_loftr_trace
bool loftr_tc_built() { return true; }

size_t loftr_tc_workspace_bytes(int N, int tokens, int C) {
    // 4 ping-pong token buffers (transformer) + packed weights of up to MAX_LAYERS layers
    return (size_t)N * tokens * C * 4 * sizeof(float) + (size_t)tc::MAX_LAYERS * tc::LAYER_PACK_BYTES + 1024;
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

int loftr_tc_run(const float* x, const float* src, const riders_loftr_weights& w, int layer_slot,
                 int N, float attn_eps, float ln_eps, float* out, void* tc_ws, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc::loftr_layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES) != cudaSuccess)
            return check_launch("loftr(tcgen05): shared memory attribute");
        configured = true;
    }
    RIDERS_REQUIRE(out != x && out != src, "loftr(tcgen05): out must not alias an input");
    tc::LayerArgs P;
    P.x = x; P.src = src; P.out = out;
    P.wpack = pack_base(tc_ws) + (size_t)layer_slot * tc::LAYER_PACK_BYTES;
    P.ln1_g = w.ln1_g; P.ln1_b = w.ln1_b; P.ln2_g = w.ln2_g; P.ln2_b = w.ln2_b;
    P.N = N; P.n_tiles = ceil_div(N, tc::PTS); P.is_self = (x == src) ? 1 : 0;
    P.attn_eps = attn_eps; P.ln_eps = ln_eps;
    P.trace = g_loftr_trace;
    const int grid = P.n_tiles < sm_count() ? P.n_tiles : sm_count();
    count_launches(1);
    tc::loftr_layer_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(P);
    return check_launch("loftr_layer_tc");
}

}  // namespace riders

extern "C" int riders_debug_tc_gemm(const float* a, const float* b, float* d, int K, int terms, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(a && b && d, "debug_tc_gemm: null pointer");
    RIDERS_REQUIRE((K == 32 || K == 64) && (terms == 1 || terms == 3), "debug_tc_gemm: K=%d terms=%d unsupported", K, terms);
    const int smem = 1024 + tc::A_BYTES;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc::tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("debug_tc_gemm: shared memory attribute");
        configured = true;
    }
    count_launches(1);
    tc::tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, K, terms);
    return check_launch("debug_tc_gemm");
}

extern "C" int riders_debug_loftr_trace(long long* device_buffer_256) {
    riders::g_loftr_trace = device_buffer_256;
    return RIDERS_OK;
}
