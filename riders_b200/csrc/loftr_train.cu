// A3 training path: LoFTREncoderLayer forward that keeps what the backward needs, and the backward itself
// (RCNet/linear_attention.py:112-135 / 18-45 differentiated by hand; the reference gets these gradients from torch autograd
// inside rcnet_main.train, RCNet/rcnet_main.py:342-358).
//
//   forward (riders_loftr_layer_fwd_train):   q = x Wq^T, k = s Wk^T, v = s Wv^T            -> saved q, k, v
//        msg = LinearAttention(q, k, v)                                                       -> saved msg
//        y1 = msg Wm^T, m = LN1(y1)                                                           -> saved y1, m
//        h = x W1[:, :C]^T + m W1[:, C:]^T,  y2 = relu(h) W2^T,  out = x + LN2(y2)           -> saved h (pre-ReLU), y2
//   backward (riders_loftr_layer_bwd), given dOut:
//        LN2 backward -> dy2, dgamma2, dbeta2;  dW2 += dy2^T relu(h);  dh = (dy2 W2) * [h > 0]
//        dW1 += dh^T [x | m];  dx += dh W1[:, :C];  dm = dh W1[:, C:]
//        LN1 backward -> dy1, dgamma1, dbeta1;  dWm += dy1^T msg;  dmsg = dy1 Wm
//        attention backward per (point, head): with Q' = elu(q)+1, K' = elu(k)+1, KV = K'^T V, Ksum, den = Q'.Ksum + eps:
//            dnum = dmsg / den, dden = -(dmsg . msg) / den, dQ' = dnum KV^T + dden Ksum, dKV = Q'^T dnum, dKsum = Q'^T dden,
//            dK' = V dKV^T + dKsum, dV = K' dKV;  dq = dQ' * elu'(q), dk = dK' * elu'(k)   (elu'(t) = 1 for t > 0, else elu(t)+1)
//        dWq += dq^T x, dx += dq Wq;  dWk += dk^T s, ds += dk Wk;  dWv += dv^T s, ds += dv Wv;  dx += dOut
//
// dtype: token tensors crossing the ABI (x, src, out, dOut, dx, dsrc) are float32 or bfloat16 (RIDERS_DTYPE_*); weights,
// weight gradients and the saved intermediates are float32 (fp32 master weights, fp32 accumulation everywhere).  Weight
// gradients ACCUMULATE into the caller's buffers (torch semantics of .grad, and a layer is called twice per transformer layer).
// The GEMMs are fp32 FFMA tiles: a training step sees B x K = 4 x 30 radar points (RCNet/train_rcnet_zju.py:28-30), i.e. 20
// tiles of tokens, so this path is sized for exactness and determinism (fixed summation order, no atomics), not for throughput.
#include "common.cuh"
#include <cuda_bf16.h>

extern "C" int riders_linear_attention_fwd(const float*, const float*, const float*, const float*, const float*,
                                           int, int, int, int, int, float, float*, void*);

namespace riders {
namespace train {

__device__ __forceinline__ float ld_tok(const void* p, size_t i, int dt) {
    return dt == RIDERS_DTYPE_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_tok(void* p, size_t i, int dt, float v) {
    if (dt == RIDERS_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(p)[i] = v;
}

// ----------------------------------------------------------------------------------------------
// C[M x N] (ldc) (+)= op(A)[M x K] * op(B)[K x N];  op(A)(m,k) = ta ? A[k lda + m] : A[m lda + k], likewise B (tb: B[n ldb + k]).
// Every operand is float32 or bfloat16 (token tensors), accumulation in float32.  64 x 64 x 16 tiles, 4 x 4 outputs per thread,
// fixed summation order.
// ----------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 32;
// blockIdx.z = K split: this CTA covers k in [z * kchunk, min(K, (z + 1) * kchunk)).  With gridDim.z > 1 the tile goes to
// part[z][M][N] (float32) and splitk_reduce_kernel adds the splits in index order (deterministic) into C.  The next K tile is
// fetched into registers while the current one is multiplied (the K loop was global-load-latency bound: the weight-gradient
// products have K = all tokens and only a handful of output tiles - 63 of the first 87 ms of a training step).
// TMv = 32 halves the tile height for products that would otherwise leave SMs idle (2520 tokens x 128 columns = 80 tiles of
// 64 x 64 on 148 SMs); TK = 32 halves the number of load -> barrier -> multiply rounds of the latency-bound K loop.
template <int TMv>
__global__ void __launch_bounds__(256)
gemm_kernel(const void* __restrict__ A, int lda, int ta, int dta, const void* __restrict__ B, int ldb, int tb, int dtb,
            void* __restrict__ Cm, int ldc, int dtc, int M, int N, int K, int accumulate, int kchunk, float* __restrict__ part) {
    constexpr int RM = TMv / 16;                         // output rows per thread
    constexpr int UA = TMv * TK / 256, UB = TN * TK / 256;
    __shared__ float As[TK][TMv + 1];
    __shared__ float Bs[TK][TN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * TMv, n0 = blockIdx.y * TN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    float acc[RM][4] = {};
    float ra[UA], rb[UB];
    auto fetch = [&](int k0) {
        // consecutive threads walk the contiguous dimension of each operand
#pragma unroll
        for (int u = 0; u < UA; ++u) {
            const int e = tid + u * 256;
            const int mi = ta ? e % TMv : e / TK, ki = ta ? e / TMv : e % TK;
            const int gm = m0 + mi, gk = k0 + ki;
            ra[u] = (gm < M && gk < kend) ? ld_tok(A, ta ? (size_t)gk * lda + gm : (size_t)gm * lda + gk, dta) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = tid + u * 256;
            const int ni = tb ? e / TK : e % TN, kj = tb ? e % TK : e / TN;
            const int gn = n0 + ni, gk2 = k0 + kj;
            rb[u] = (gn < N && gk2 < kend) ? ld_tok(B, tb ? (size_t)gn * ldb + gk2 : (size_t)gk2 * ldb + gn, dtb) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int u = 0; u < UA; ++u) {
            const int e = tid + u * 256;
            const int mi = ta ? e % TMv : e / TK, ki = ta ? e / TMv : e % TK;
            As[ki][mi] = ra[u];
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = tid + u * 256;
            const int ni = tb ? e / TK : e % TN, kj = tb ? e % TK : e / TN;
            Bs[kj][ni] = rb[u];
        }
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
        stash();
        __syncthreads();
        if (k0 + TK < kend) fetch(k0 + TK);
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[RM], b[4];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int gm = m0 + ty * RM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            if (gridDim.z > 1) { part[((size_t)blockIdx.z * M + gm) * N + gn] = acc[i][j]; continue; }
            const size_t o = (size_t)gm * ldc + gn;
            st_tok(Cm, o, dtc, accumulate ? ld_tok(Cm, o, dtc) + acc[i][j] : acc[i][j]);
        }
    }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, int splits, void* __restrict__ Cm, int ldc, int dtc, int M, int N, int accumulate) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= M * N) return;
    const int gm = e / N, gn = e % N;
    float a = 0.f;
    for (int z = 0; z < splits; ++z) a += part[(size_t)z * M * N + e];
    const size_t o = (size_t)gm * ldc + gn;
    st_tok(Cm, o, dtc, accumulate ? ld_tok(Cm, o, dtc) + a : a);
}

constexpr int SPLITK_CHUNK = 256, SPLITK_MAX = 64;
static size_t splitk_floats(int C) { return (size_t)SPLITK_MAX * 2 * C * 2 * C; }

// `part` (splitk_floats) enables the K split for products with few output tiles and a long K (the weight gradients)
static void gemm(const void* A, int lda, bool ta, int dta, const void* B, int ldb, bool tb, int dtb, void* C, int ldc, int dtc,
                 int M, int N, int K, bool accumulate, cudaStream_t st, float* part = nullptr) {
    int splits = 1, kchunk = K;
    if (part && K > 2 * SPLITK_CHUNK && ceil_div(M, TM) * ceil_div(N, TN) < 64) {
        splits = ceil_div(K, SPLITK_CHUNK);
        if (splits > SPLITK_MAX) splits = SPLITK_MAX;
        kchunk = ceil_div(ceil_div(K, splits), TK) * TK;
        splits = ceil_div(K, kchunk);
    }
    count_launches(splits > 1 ? 2 : 1);
    if (ceil_div(M, TM) * ceil_div(N, TN) * splits < 148)        // would leave SMs idle: half-height tiles
        gemm_kernel<32><<<dim3(ceil_div(M, 32), ceil_div(N, TN), splits), 256, 0, st>>>(A, lda, ta, dta, B, ldb, tb, dtb, C, ldc, dtc, M, N, K,
                                                                                        accumulate, kchunk, part);
    else
        gemm_kernel<TM><<<dim3(ceil_div(M, TM), ceil_div(N, TN), splits), 256, 0, st>>>(A, lda, ta, dta, B, ldb, tb, dtb, C, ldc, dtc, M, N, K,
                                                                                        accumulate, kchunk, part);
    if (splits > 1)
        splitk_reduce_kernel<<<ceil_div(M * N, 256), 256, 0, st>>>(part, splits, C, ldc, dtc, M, N, accumulate);
}
constexpr int F32 = RIDERS_DTYPE_F32;

// out[dt] = (residual[dt] ? residual : 0) + LN(in) * gamma + beta; one warp per row
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
              const void* __restrict__ residual, int dtr, void* __restrict__ out, int dto, int M, int C, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
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
        if (residual) y += ld_tok(residual, (size_t)row * C + c, dtr);
        st_tok(out, (size_t)row * C + c, dto, y);
    }
}

// LayerNorm backward of one row per warp: dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy * gamma.
// Per-CTA partial sums of dgamma / dbeta (8 rows) go to `part` [ctas][2][C]; ln_bwd_reduce_kernel adds them in CTA order.
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ y_in, const float* __restrict__ gamma, const void* __restrict__ dy, int dtdy,
              float* __restrict__ dx, float* __restrict__ part, int M, int C, float eps) {
    extern __shared__ float sm[];                        // [8][2][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    float* mine = sm + (size_t)warp * 2 * C;
    for (int c = lane; c < 2 * C; c += 32) mine[c] = 0.f;
    if (row < M) {
        const float* p = y_in + (size_t)row * C;
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
        float sg = 0.f, sgx = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float d = ld_tok(dy, (size_t)row * C + c, dtdy), xh = (p[c] - mean) * rstd, g = d * gamma[c];
            sg += g; sgx = fmaf(g, xh, sgx);
            mine[c] = d * xh;                              // dgamma contribution
            mine[C + c] = d;                               // dbeta contribution
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, o); sgx += __shfl_xor_sync(0xffffffffu, sgx, o); }
        const float mg = sg / (float)C, mgx = sgx / (float)C;
        for (int c = lane; c < C; c += 32) {
            const float d = ld_tok(dy, (size_t)row * C + c, dtdy), xh = (p[c] - mean) * rstd;
            dx[(size_t)row * C + c] = rstd * (d * gamma[c] - mg - xh * mgx);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += sm[(size_t)w * 2 * C + c];
        part[(size_t)blockIdx.x * 2 * C + c] = a;
    }
}

__global__ void __launch_bounds__(256)
ln_bwd_reduce_kernel(const float* __restrict__ part, int ctas, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= 2 * C) return;
    float a = 0.f;
    for (int i = 0; i < ctas; ++i) a += part[(size_t)i * 2 * C + c];
    if (c < C) dgamma[c] += a; else dbeta[c - C] += a;
}

__global__ void __launch_bounds__(256)
relu_kernel(const float* __restrict__ h, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = fmaxf(h[i], 0.f);
}
__global__ void __launch_bounds__(256)
relu_mask_kernel(const float* __restrict__ h, float* __restrict__ g, size_t n) {           // g *= [h > 0]
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) if (!(h[i] > 0.f)) g[i] = 0.f;
}
// dst[dt] (+)= src (+ extra[dt])
__global__ void __launch_bounds__(256)
add_store_kernel(const float* __restrict__ a, const void* __restrict__ extra, int dte, void* __restrict__ dst, int dtd, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        st_tok(dst, i, dtd, a[i] + (extra ? ld_tok(extra, i, dte) : 0.f));
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.0f : expf(x); }

// Attention backward: one CTA per (point, head).  D <= 32, L, S <= 64.
constexpr int AB_THREADS = 128, AB_D = 32, AB_T = 64;
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     const float* __restrict__ msg, const float* __restrict__ dmsg, int L, int S, int H, int D, float eps,
                     float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
    __shared__ float Q[AB_T][AB_D + 1], K[AB_T][AB_D + 1], V[AB_T][AB_D + 1], dN[AB_T][AB_D + 1];
    __shared__ float KV[AB_D][AB_D + 1], dKV[AB_D][AB_D + 1], Ksum[AB_D], dKsum[AB_D], dden[AB_T];
    const int n = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
    const size_t tok = (size_t)H * D;
    const size_t qb = (size_t)n * L * tok + (size_t)h * D, kb = (size_t)n * S * tok + (size_t)h * D;
    for (int e = tid; e < L * D; e += AB_THREADS) Q[e / D][e % D] = elu1(q[qb + (size_t)(e / D) * tok + e % D]);
    for (int e = tid; e < S * D; e += AB_THREADS) {
        K[e / D][e % D] = elu1(k[kb + (size_t)(e / D) * tok + e % D]);
        V[e / D][e % D] = v[kb + (size_t)(e / D) * tok + e % D];
    }
    __syncthreads();
    for (int e = tid; e < D * D; e += AB_THREADS) {
        const int d = e / D, c = e % D;
        float a = 0.f;
        for (int s = 0; s < S; ++s) a = fmaf(K[s][d], V[s][c], a);
        KV[d][c] = a;
    }
    for (int d = tid; d < D; d += AB_THREADS) { float a = 0.f; for (int s = 0; s < S; ++s) a += K[s][d]; Ksum[d] = a; }
    __syncthreads();
    for (int l = tid; l < L; l += AB_THREADS) {
        float den = eps;
        for (int d = 0; d < D; ++d) den = fmaf(Q[l][d], Ksum[d], den);
        float dot = 0.f;
        for (int c = 0; c < D; ++c) {
            const float g = dmsg[qb + (size_t)l * tok + c];
            dN[l][c] = g / den;
            dot = fmaf(g, msg[qb + (size_t)l * tok + c], dot);
        }
        dden[l] = -dot / den;
    }
    __syncthreads();
    for (int e = tid; e < D * D; e += AB_THREADS) {
        const int d = e / D, c = e % D;
        float a = 0.f;
        for (int l = 0; l < L; ++l) a = fmaf(Q[l][d], dN[l][c], a);
        dKV[d][c] = a;
    }
    for (int d = tid; d < D; d += AB_THREADS) { float a = 0.f; for (int l = 0; l < L; ++l) a = fmaf(Q[l][d], dden[l], a); dKsum[d] = a; }
    for (int e = tid; e < L * D; e += AB_THREADS) {
        const int l = e / D, d = e % D;
        float a = dden[l] * Ksum[d];
        for (int c = 0; c < D; ++c) a = fmaf(dN[l][c], KV[d][c], a);
        const float qraw = q[qb + (size_t)l * tok + d];
        dq[qb + (size_t)l * tok + d] = a * (qraw > 0.f ? 1.0f : Q[l][d]);
    }
    __syncthreads();
    for (int e = tid; e < S * D; e += AB_THREADS) {
        const int s = e / D, d = e % D;
        float a = dKsum[d], b = 0.f;
        for (int c = 0; c < D; ++c) { a = fmaf(dKV[d][c], V[s][c], a); b = fmaf(K[s][c], dKV[c][d], b); }
        const float kraw = k[kb + (size_t)s * tok + d];
        dk[kb + (size_t)s * tok + d] = a * (kraw > 0.f ? 1.0f : K[s][d]);
        dv[kb + (size_t)s * tok + d] = b;                  // dV[s][v=d] = sum_c K'[s][c] dKV[c][d]
    }
}

struct Saved { float *q, *k, *v, *msg, *y1, *m, *h, *y2; };
static Saved carve_saved(float* p, size_t tx, size_t ts, int C) {
    Saved s;
    s.q = p; p += tx * C; s.k = p; p += ts * C; s.v = p; p += ts * C; s.msg = p; p += tx * C;
    s.y1 = p; p += tx * C; s.m = p; p += tx * C; s.h = p; p += tx * 2 * C; s.y2 = p;
    return s;
}

}  // namespace train
}  // namespace riders

extern "C" size_t riders_loftr_saved_bytes(int N, int L, int S, int d_model) {
    if (N <= 0) return 256;
    return ((size_t)N * L * 7 + (size_t)N * S * 2) * d_model * sizeof(float) + 256;
}
extern "C" size_t riders_loftr_train_workspace_bytes(int N, int L, int S, int d_model) {
    if (N <= 0) return 256;
    const size_t t = (size_t)N * (L > S ? L : S);
    // 2C-wide scratch (relu(h) / dh) + five C-wide scratch rows + LayerNorm partials
    return (t * 2 * d_model * 2 + t * d_model * 6) * sizeof(float) + ((t + 7) / 8) * 2 * d_model * sizeof(float) +
           riders::train::splitk_floats(d_model) * sizeof(float) + 1024;
}

static int check_train(const riders_loftr_weights* w, int N, int L, int S, int C, int nhead, int dtype) {
    RIDERS_REQUIRE(N >= 0 && L > 0 && S > 0 && C > 0 && nhead > 0 && C % nhead == 0, "loftr_train: bad shape");
    RIDERS_REQUIRE(C / nhead <= riders::train::AB_D && L <= riders::train::AB_T && S <= riders::train::AB_T,
                   "loftr_train: head dim <= %d and <= %d tokens supported", riders::train::AB_D, riders::train::AB_T);
    RIDERS_REQUIRE(dtype == RIDERS_DTYPE_F32 || dtype == RIDERS_DTYPE_BF16, "loftr_train: dtype %d", dtype);
    RIDERS_REQUIRE(w && w->wq && w->wk && w->wv && w->wm && w->w1 && w->w2 && w->ln1_g && w->ln1_b && w->ln2_g && w->ln2_b,
                   "loftr_train: null weight pointer");
    return RIDERS_OK;
}

extern "C" int riders_loftr_layer_fwd_train(const void* x, const void* src, const riders_loftr_weights* w,
                                            int N, int L, int S, int d_model, int nhead, float attn_eps, float ln_eps, int dtype,
                                            void* out, void* saved, void* workspace, void* stream) {
    using namespace riders;
    using namespace riders::train;
    int rc = check_train(w, N, L, S, d_model, nhead, dtype);
    if (rc != RIDERS_OK) return rc;
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && src && out && saved && workspace, "loftr_layer_fwd_train: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = d_model, Mx = N * L, Ms = N * S;
    const Saved sv = carve_saved((float*)saved, Mx, Ms, C);
    float* hr = (float*)workspace;                                        // relu(h), Mx x 2C
    gemm(x, C, false, dtype, w->wq, C, true, F32, sv.q, C, RIDERS_DTYPE_F32, Mx, C, C, false, st);
    gemm(src, C, false, dtype, w->wk, C, true, F32, sv.k, C, RIDERS_DTYPE_F32, Ms, C, C, false, st);
    gemm(src, C, false, dtype, w->wv, C, true, F32, sv.v, C, RIDERS_DTYPE_F32, Ms, C, C, false, st);
    rc = riders_linear_attention_fwd(sv.q, sv.k, sv.v, nullptr, nullptr, N, L, S, nhead, C / nhead, attn_eps, sv.msg, stream);
    if (rc != RIDERS_OK) return rc;
    gemm(sv.msg, C, false, RIDERS_DTYPE_F32, w->wm, C, true, F32, sv.y1, C, RIDERS_DTYPE_F32, Mx, C, C, false, st);
    count_launches(4);
    ln_fwd_kernel<<<ceil_div(Mx, 8), 256, 0, st>>>(sv.y1, w->ln1_g, w->ln1_b, nullptr, 0, sv.m, RIDERS_DTYPE_F32, Mx, C, ln_eps);
    gemm(x, C, false, dtype, w->w1, 2 * C, true, F32, sv.h, 2 * C, RIDERS_DTYPE_F32, Mx, 2 * C, C, false, st);               // x   W1[:, :C]^T
    gemm(sv.m, C, false, RIDERS_DTYPE_F32, w->w1 + C, 2 * C, true, F32, sv.h, 2 * C, RIDERS_DTYPE_F32, Mx, 2 * C, C, true, st);  // + m W1[:, C:]^T
    relu_kernel<<<1024, 256, 0, st>>>(sv.h, hr, (size_t)Mx * 2 * C);
    gemm(hr, 2 * C, false, RIDERS_DTYPE_F32, w->w2, 2 * C, true, F32, sv.y2, C, RIDERS_DTYPE_F32, Mx, C, 2 * C, false, st);
    ln_fwd_kernel<<<ceil_div(Mx, 8), 256, 0, st>>>(sv.y2, w->ln2_g, w->ln2_b, x, dtype, out, dtype, Mx, C, ln_eps);
    return check_launch("loftr_layer_fwd_train");
}

extern "C" int riders_loftr_layer_bwd(const void* x, const void* src, const riders_loftr_weights* w, const void* saved,
                                      const void* d_out, int N, int L, int S, int d_model, int nhead,
                                      float attn_eps, float ln_eps, int dtype,
                                      void* d_x, void* d_src, const riders_loftr_weights* d_w,
                                      void* workspace, void* stream) {
    using namespace riders;
    using namespace riders::train;
    int rc = check_train(w, N, L, S, d_model, nhead, dtype);
    if (rc != RIDERS_OK) return rc;
    RIDERS_REQUIRE(d_w && d_w->wq && d_w->wk && d_w->wv && d_w->wm && d_w->w1 && d_w->w2 && d_w->ln1_g && d_w->ln1_b && d_w->ln2_g && d_w->ln2_b,
                   "loftr_layer_bwd: null weight-gradient pointer");
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(x && src && saved && d_out && d_x && workspace, "loftr_layer_bwd: null pointer");
    const bool self = (x == src);
    RIDERS_REQUIRE(self || d_src, "loftr_layer_bwd: d_src is required for a cross layer");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = d_model, Mx = N * L, Ms = N * S;
    const size_t T = (size_t)N * (L > S ? L : S);
    const Saved sv = carve_saved((float*)const_cast<void*>(saved), Mx, Ms, C);
    float* p = (float*)workspace;
    float* hr = p; p += T * 2 * C;                       // relu(h)
    float* dh = p; p += T * 2 * C;
    float* dy = p; p += T * C;                           // dy2, then dy1
    float* dm = p; p += T * C;                           // dm, then dmsg, then the float32 accumulator of dsrc
    float* dq = p; p += T * C;
    float* dk = p; p += T * C;
    float* dv = p; p += T * C;
    float* dxa = p; p += T * C;                          // float32 accumulator of dx
    float* part = p; p += (size_t)((T + 7) / 8) * 2 * C;  // LayerNorm partials [ctas][2][C]
    float* sk = p;                                       // split-K partials of the weight-gradient products
    const int lnc = ceil_div(Mx, 8);
    auto G = [](const float* q) { return const_cast<float*>(q); };
    const size_t ln_smem = (size_t)8 * 2 * C * sizeof(float);
    count_launches(9);
    // ---- out = x + LN2(y2)
    ln_bwd_kernel<<<lnc, 256, ln_smem, st>>>(sv.y2, w->ln2_g, d_out, dtype, dy, part, Mx, C, ln_eps);
    ln_bwd_reduce_kernel<<<ceil_div(2 * C, 256), 256, 0, st>>>(part, lnc, C, G(d_w->ln2_g), G(d_w->ln2_b));
    // ---- y2 = relu(h) W2^T
    relu_kernel<<<1024, 256, 0, st>>>(sv.h, hr, (size_t)Mx * 2 * C);
    gemm(dy, C, true, F32, hr, 2 * C, false, F32, G(d_w->w2), 2 * C, F32, C, 2 * C, Mx, true, st, sk);          // dW2 += dy2^T relu(h)
    gemm(dy, C, false, F32, w->w2, 2 * C, false, F32, dh, 2 * C, F32, Mx, 2 * C, C, false, st);             // dh = dy2 W2
    relu_mask_kernel<<<1024, 256, 0, st>>>(sv.h, dh, (size_t)Mx * 2 * C);
    // ---- h = [x | m] W1^T
    gemm(dh, 2 * C, true, F32, x, C, false, dtype, G(d_w->w1), 2 * C, F32, 2 * C, C, Mx, true, st, sk);        // dW1[:, :C] += dh^T x
    gemm(dh, 2 * C, true, F32, sv.m, C, false, F32, G(d_w->w1) + C, 2 * C, F32, 2 * C, C, Mx, true, st, sk);   // dW1[:, C:] += dh^T m
    gemm(dh, 2 * C, false, F32, w->w1, 2 * C, false, F32, dxa, C, F32, Mx, C, 2 * C, false, st);            // dx  = dh W1[:, :C]
    gemm(dh, 2 * C, false, F32, w->w1 + C, 2 * C, false, F32, dm, C, F32, Mx, C, 2 * C, false, st);         // dm  = dh W1[:, C:]
    // ---- m = LN1(y1), y1 = msg Wm^T
    ln_bwd_kernel<<<lnc, 256, ln_smem, st>>>(sv.y1, w->ln1_g, dm, F32, dy, part, Mx, C, ln_eps);
    ln_bwd_reduce_kernel<<<ceil_div(2 * C, 256), 256, 0, st>>>(part, lnc, C, G(d_w->ln1_g), G(d_w->ln1_b));
    gemm(dy, C, true, F32, sv.msg, C, false, F32, G(d_w->wm), C, F32, C, C, Mx, true, st, sk);                  // dWm += dy1^T msg
    gemm(dy, C, false, F32, w->wm, C, false, F32, dm, C, F32, Mx, C, C, false, st);                         // dmsg = dy1 Wm
    // ---- attention
    attention_bwd_kernel<<<N * nhead, AB_THREADS, 0, st>>>(sv.q, sv.k, sv.v, sv.msg, dm, L, S, nhead, C / nhead, attn_eps, dq, dk, dv);
    // ---- q = x Wq^T, k = s Wk^T, v = s Wv^T
    gemm(dq, C, true, F32, x, C, false, dtype, G(d_w->wq), C, F32, C, C, Mx, true, st, sk);
    gemm(dk, C, true, F32, src, C, false, dtype, G(d_w->wk), C, F32, C, C, Ms, true, st, sk);
    gemm(dv, C, true, F32, src, C, false, dtype, G(d_w->wv), C, F32, C, C, Ms, true, st, sk);
    gemm(dq, C, false, F32, w->wq, C, false, F32, dxa, C, F32, Mx, C, C, true, st);                          // dx += dq Wq
    float* dsa = self ? dxa : dm;                                                                            // dmsg is dead
    gemm(dk, C, false, F32, w->wk, C, false, F32, dsa, C, F32, Ms, C, C, self, st);                          // ds (+)= dk Wk
    gemm(dv, C, false, F32, w->wv, C, false, F32, dsa, C, F32, Ms, C, C, true, st);                          // ds += dv Wv
    // ---- residual, and out to the caller's dtype
    add_store_kernel<<<1024, 256, 0, st>>>(dxa, d_out, dtype, d_x, dtype, (size_t)Mx * C);
    if (!self) add_store_kernel<<<1024, 256, 0, st>>>(dsa, nullptr, 0, d_src, dtype, (size_t)Ms * C);
    return check_launch("loftr_layer_bwd");
}
