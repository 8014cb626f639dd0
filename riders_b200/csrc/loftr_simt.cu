// A3/A4, fp32 FFMA path: LoFTREncoderLayer / LocalFeatureTransformer
// (RCNet/linear_attention.py:112-135, 159-184 of the reference).
//
// Straightforward, un-fused restatement used (a) as the on-device cross-check of the fused
// tcgen05 layer and (b) for shapes the fused kernel does not cover.  Per layer:
//   q = x Wq^T, k = src Wk^T, v = src Wv^T           (sgemm_nt)
//   msg = LinearAttention(q, k, v)                    (linear_attention.cu)
//   m = LN1(msg Wm^T)                                 (sgemm_nt + layer_norm)
//   h = relu(x W1[:, :C]^T + m W1[:, C:]^T)           (two accumulating sgemm_nt, no concat copy)
//   out = x + LN2(h W2^T)                             (sgemm_nt + layer_norm with residual)
// This file also owns the C-ABI entry points of A3/A4 and dispatches to the tcgen05 layer.
#include "common.cuh"

extern "C" int riders_linear_attention_fwd(const float*, const float*, const float*, const float*, const float*,
                                           int, int, int, int, int, float, float*, void*);

namespace riders {

// ----------------------------------------------------------------------------------------------
// C[M x N] (+)= A[M x K] (lda) * W[N x K]^T (ldw), optional ReLU.  128x128x16 tiles, 8x8 / thread.
// ----------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wt, int ldw,
                float* __restrict__ Cm, int ldc, int M, int N, int K, int accumulate, int relu) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Ws[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;            // 16 x 16 threads, 8 x 8 outputs each
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // loader mapping: 128 rows x 16 k = 512 float4, two per thread
    const int lr = tid / 4;                              // 0..63 (+64)
    const int lk = (tid % 4) * 4;                        // 0,4,8,12
    float4 ra[2], rw[2];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = lr + u * 64;
            const int gm = m0 + row, gn = n0 + row;
            ra[u] = (gm < M) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + k0 + lk) : make_float4(0, 0, 0, 0);
            rw[u] = (gn < N) ? __ldg(reinterpret_cast<const float4*>(Wt + (size_t)gn * ldw + k0 + lk)) : make_float4(0, 0, 0, 0);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = lr + u * 64;
            As[buf][lk + 0][row] = ra[u].x; As[buf][lk + 1][row] = ra[u].y;
            As[buf][lk + 2][row] = ra[u].z; As[buf][lk + 3][row] = ra[u].w;
            Ws[buf][lk + 0][row] = rw[u].x; Ws[buf][lk + 1][row] = rw[u].y;
            Ws[buf][lk + 2][row] = rw[u].z; Ws[buf][lk + 3][row] = rw[u].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int nk = K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int gn = n0 + jh * 64 + tx * 4;
            if (gn >= N) continue;                        // N % 4 == 0 is required by the host
            float4 o = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            float4* dst = reinterpret_cast<float4*>(Cm + (size_t)gm * ldc + gn);
            if (accumulate) { const float4 c = *dst; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *dst = o;
        }
    }
}

static void sgemm_nt(const float* A, int lda, const float* Wt, int ldw, float* C, int ldc,
                     int M, int N, int K, bool accumulate, bool relu, cudaStream_t st) {
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    count_launches(1);
    sgemm_nt_kernel<<<grid, SG_THREADS, 0, st>>>(A, lda, Wt, ldw, C, ldc, M, N, K, accumulate ? 1 : 0, relu ? 1 : 0);
}

// ----------------------------------------------------------------------------------------------
// out = (residual ? residual : 0) + LayerNorm(in) * gamma + beta ; one warp per row
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layer_norm_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ residual, float* __restrict__ out, int M, int C, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* p = in + (size_t)row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += p[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = p[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)C + eps);
    for (int c = lane; c < C; c += 32) {
        float y = (p[c] - mean) * rstd * gamma[c] + beta[c];
        if (residual) y += residual[(size_t)row * C + c];
        out[(size_t)row * C + c] = y;
    }
}

struct SimtScratch { float *q, *k, *v, *msg, *m, *h, *ping0, *ping1; };

constexpr int STATUS_FLOATS = 64;             // workspace status block (256 B): word 0 = fp16-range overflow flag of the tcgen05 path

static size_t simt_scratch_floats(int N, int tokens, int C) {
    const size_t rows = (size_t)N * tokens;
    return rows * C * 7 + rows * 2 * C;     // q,k,v,msg,m,ping0,ping1 + h
}

static SimtScratch carve(float* ws, int N, int tokens, int C) {
    const size_t a = (size_t)N * tokens * C;
    SimtScratch s;
    s.q = ws; s.k = ws + a; s.v = ws + 2 * a; s.msg = ws + 3 * a; s.m = ws + 4 * a;
    s.ping0 = ws + 5 * a; s.ping1 = ws + 6 * a; s.h = ws + 7 * a;
    return s;
}

static int loftr_layer_simt(const float* x, const float* src, const riders_loftr_weights& w,
                            int N, int L, int S, int C, int nhead, float attn_eps, float ln_eps,
                            float* out, const SimtScratch& s, cudaStream_t st) {
    const int Mx = N * L, Ms = N * S;
    sgemm_nt(x, C, w.wq, C, s.q, C, Mx, C, C, false, false, st);
    sgemm_nt(src, C, w.wk, C, s.k, C, Ms, C, C, false, false, st);
    sgemm_nt(src, C, w.wv, C, s.v, C, Ms, C, C, false, false, st);
    int rc = riders_linear_attention_fwd(s.q, s.k, s.v, nullptr, nullptr, N, L, S, nhead, C / nhead, attn_eps, s.msg, st);
    if (rc != RIDERS_OK) return rc;
    sgemm_nt(s.msg, C, w.wm, C, s.q, C, Mx, C, C, false, false, st);                 // merge -> q buffer
    count_launches(2);
    layer_norm_kernel<<<ceil_div(Mx, 8), 256, 0, st>>>(s.q, w.ln1_g, w.ln1_b, nullptr, s.m, Mx, C, ln_eps);
    sgemm_nt(x, C, w.w1, 2 * C, s.h, 2 * C, Mx, 2 * C, C, false, false, st);         // x   * W1[:, :C]^T
    sgemm_nt(s.m, C, w.w1 + C, 2 * C, s.h, 2 * C, Mx, 2 * C, C, true, true, st);     // + m * W1[:, C:]^T, relu
    sgemm_nt(s.h, 2 * C, w.w2, 2 * C, s.q, C, Mx, C, 2 * C, false, false, st);
    layer_norm_kernel<<<ceil_div(Mx, 8), 256, 0, st>>>(s.q, w.ln2_g, w.ln2_b, x, out, Mx, C, ln_eps);
    return check_launch("loftr_layer_simt");
}

// implemented in loftr_tcgen05.cu
bool loftr_tc_built();
size_t loftr_tc_workspace_bytes(int N, int tokens, int C);
int loftr_tc_pack(const riders_loftr_weights& w, int layer_slot, void* tc_ws, cudaStream_t st);
int loftr_tc_run_calls(int n_calls, const float* const* x, const float* const* src, float* const* out,
                       const riders_loftr_weights* const* w, const int* slot,
                       int N, float attn_eps, float ln_eps, void* tc_ws, int* overflow, cudaStream_t st);

static int validate(const riders_loftr_weights* w, int N, int L, int S, int C, int nhead, int impl) {
    RIDERS_REQUIRE(N >= 0 && L > 0 && S > 0 && C > 0 && nhead > 0, "loftr: bad shape");
    RIDERS_REQUIRE(C % nhead == 0 && C % 16 == 0 && C / nhead <= 64, "loftr: d_model=%d nhead=%d unsupported", C, nhead);
    RIDERS_REQUIRE(impl == RIDERS_LOFTR_SIMT || impl == RIDERS_LOFTR_TCGEN05, "loftr: unknown impl %d", impl);
    RIDERS_REQUIRE(w && w->wq && w->wk && w->wv && w->wm && w->w1 && w->w2 && w->ln1_g && w->ln1_b && w->ln2_g && w->ln2_b,
                   "loftr: null weight pointer");
    if (impl == RIDERS_LOFTR_TCGEN05)
        RIDERS_REQUIRE(C == 128 && nhead == 8 && L == 21 && S == 21,
                       "loftr: the tcgen05 layer is specialised for d_model=128, nhead=8, 21 tokens (got C=%d H=%d L=%d S=%d)", C, nhead, L, S);
    return RIDERS_OK;
}

}  // namespace riders

extern "C" int riders_loftr_impl_available(int impl, int d_model, int nhead, int tokens) {
    if (impl == RIDERS_LOFTR_SIMT) return d_model > 0 && nhead > 0 && d_model % nhead == 0 && d_model % 16 == 0 && d_model / nhead <= 64;
    if (impl == RIDERS_LOFTR_TCGEN05) return riders::loftr_tc_built() && d_model == 128 && nhead == 8 && tokens == 21;
    return 0;
}

extern "C" size_t riders_loftr_workspace_bytes(int N, int tokens, int d_model, int impl) {
    if (N <= 0 || tokens <= 0 || d_model <= 0) return 256;
    size_t b = riders::simt_scratch_floats(N, tokens, d_model) * sizeof(float);
    if (impl == RIDERS_LOFTR_TCGEN05) b = riders::loftr_tc_workspace_bytes(N, tokens, d_model);
    return b + 256 + riders::STATUS_FLOATS * sizeof(float);
}

extern "C" int riders_loftr_layer_fwd(const float* x, const float* src, const riders_loftr_weights* w,
                                      int N, int L, int S, int d_model, int nhead,
                                      float attn_eps, float ln_eps, int impl,
                                      float* out, void* workspace, void* stream) {
    using namespace riders;
    int rc = validate(w, N, L, S, d_model, nhead, impl);
    if (rc != RIDERS_OK) return rc;
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && src && out && workspace, "loftr_layer: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == RIDERS_LOFTR_TCGEN05) {
        RIDERS_REQUIRE(out != x && out != src, "loftr_layer: out must not alias an input");
        int* overflow = (int*)workspace;                      // status block: word 0 = fp16-range overflow flag
        void* tc_ws = (float*)workspace + STATUS_FLOATS;
        if (cudaMemsetAsync(overflow, 0, sizeof(int), st) != cudaSuccess) return check_launch("loftr_layer: status reset");
        if ((rc = loftr_tc_pack(*w, 0, tc_ws, st)) != RIDERS_OK) return rc;
        const int slot = 0;
        return loftr_tc_run_calls(1, &x, &src, &out, &w, &slot, N, attn_eps, ln_eps, tc_ws, overflow, st);
    }
    const SimtScratch s = carve((float*)workspace + STATUS_FLOATS, N, L > S ? L : S, d_model);
    return loftr_layer_simt(x, src, *w, N, L, S, d_model, nhead, attn_eps, ln_eps, out, s, st);
}

extern "C" int riders_loftr_transformer_fwd(const float* feat0, const float* feat1,
                                            const riders_loftr_weights* layers, const int* layer_kind, int n_layers,
                                            int N, int L, int S, int d_model, int nhead,
                                            float attn_eps, float ln_eps, int impl,
                                            float* out0, float* out1, void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(n_layers > 0 && layers && layer_kind, "loftr_transformer: no layers");
    for (int i = 0; i < n_layers; ++i) {
        int rc = validate(&layers[i], N, L, S, d_model, nhead, impl);
        if (rc != RIDERS_OK) return rc;
        RIDERS_REQUIRE(layer_kind[i] == 0 || layer_kind[i] == 1, "loftr_transformer: layer_kind[%d]=%d", i, layer_kind[i]);
    }
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(feat0 && feat1 && out0 && out1 && workspace, "loftr_transformer: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int tokens = L > S ? L : S;
    if (impl == RIDERS_LOFTR_TCGEN05) {
        // the whole transformer in ONE launch: every layer's weights split + swizzled once, then the list of 2 n_layers
        // layer calls; a tile's tokens stay with its CTA through all of them (in place in out0 / out1 after their first use)
        RIDERS_REQUIRE(2 * n_layers <= 16, "loftr_transformer(tcgen05): at most 8 layers (16 layer calls) per call");
        RIDERS_REQUIRE(out0 != feat0 && out0 != feat1 && out1 != feat0 && out1 != feat1 && out0 != out1,
                       "loftr_transformer: outputs must not alias the inputs");
        int* overflow = (int*)workspace;
        void* tc_ws = (float*)workspace + STATUS_FLOATS;
        if (cudaMemsetAsync(overflow, 0, sizeof(int), st) != cudaSuccess) return check_launch("loftr_transformer: status reset");
        const float* xs[16]; const float* ss[16]; float* os[16]; const riders_loftr_weights* ws_[16]; int slots[16];
        const float* c0 = feat0;
        const float* c1 = feat1;
        int nc = 0;
        for (int i = 0; i < n_layers; ++i) {
            int rc = loftr_tc_pack(layers[i], i, tc_ws, st);
            if (rc != RIDERS_OK) return rc;
            if (layer_kind[i] == 0) {                   // 'self'  (linear_attention.py:175-177)
                xs[nc] = c0; ss[nc] = c0; os[nc] = out0; ws_[nc] = &layers[i]; slots[nc++] = i;
                xs[nc] = c1; ss[nc] = c1; os[nc] = out1; ws_[nc] = &layers[i]; slots[nc++] = i;
            } else {                                    // 'cross' (linear_attention.py:178-180): the second call sees the new feat0
                xs[nc] = c0; ss[nc] = c1; os[nc] = out0; ws_[nc] = &layers[i]; slots[nc++] = i;
                xs[nc] = c1; ss[nc] = out0; os[nc] = out1; ws_[nc] = &layers[i]; slots[nc++] = i;
            }
            c0 = out0; c1 = out1;
        }
        return loftr_tc_run_calls(nc, xs, ss, os, ws_, slots, N, attn_eps, ln_eps, tc_ws, overflow, st);
    }
    float* ws = (float*)workspace + STATUS_FLOATS;
    const SimtScratch s = carve(ws, N, tokens, d_model);
    // SIMT path: two ping-pong buffer pairs.  The layer at even distance from the end writes pair B, odd
    // distance pair A, so the last layer can always write out0/out1 without reading them.
    float *pa0 = s.ping0, *pa1 = s.ping1, *pb0 = out0, *pb1 = out1;     // out buffers double as pair B
    const float* c0 = feat0;
    const float* c1 = feat1;
    for (int i = 0; i < n_layers; ++i) {
        const int to_end = n_layers - 1 - i;
        float* n0 = (to_end == 0) ? out0 : ((to_end & 1) ? pa0 : pb0);
        float* n1 = (to_end == 0) ? out1 : ((to_end & 1) ? pa1 : pb1);
        auto run = [&](const float* xx, const float* ss, int Lx, int Ss, float* oo) -> int {
            return loftr_layer_simt(xx, ss, layers[i], N, Lx, Ss, d_model, nhead, attn_eps, ln_eps, oo, s, st);
        };
        int rc;
        if (layer_kind[i] == 0) {                       // 'self'  (linear_attention.py:175-177)
            if ((rc = run(c0, c0, L, L, n0)) != RIDERS_OK) return rc;
            if ((rc = run(c1, c1, S, S, n1)) != RIDERS_OK) return rc;
        } else {                                        // 'cross' (linear_attention.py:178-180)
            if ((rc = run(c0, c1, L, S, n0)) != RIDERS_OK) return rc;
            if ((rc = run(c1, n0, S, L, n1)) != RIDERS_OK) return rc;
        }
        c0 = n0; c1 = n1;
    }
    return RIDERS_OK;
}
