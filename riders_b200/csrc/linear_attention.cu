// A2: LinearAttention.forward (RCNet/linear_attention.py:18-45 of the reference).
//   Q = elu(q)+1, K = elu(k)+1, V = v/S;  KV[d][v] = sum_s K[s][d] V[s][v];
//   Z[l] = 1/(Q[l].sum_s K[s] + eps);     out[l][v] = (sum_d Q[l][d] KV[d][v]) * Z[l] * S
// One CTA per (sample n, head h).  KV (D x D) and Ksum (D) live in shared memory; the S loop
// streams K/V rows from global memory (21 tokens in RC-Net, any S supported).  Stand-alone
// entry point for the `LinearAttention` class boundary; the fused layer kernels contain the
// same math inline.
#include "common.cuh"

namespace riders {

constexpr int LA_THREADS = 128;
constexpr int LA_MAX_D = 64;

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.0f : expf(x); }   // elu(x)+1; expm1(x)+1 == exp(x) to 1 ulp

__global__ void __launch_bounds__(LA_THREADS)
linear_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                        const float* __restrict__ q_mask, const float* __restrict__ kv_mask,
                        int L, int S, int H, int D, float eps, float* __restrict__ out) {
    __shared__ float KV[LA_MAX_D * LA_MAX_D];
    __shared__ float Ksum[LA_MAX_D];
    const int n = blockIdx.x / H, h = blockIdx.x % H;
    const size_t tok = (size_t)H * D;                       // floats per token
    const float* kb = k + (size_t)n * S * tok + (size_t)h * D;
    const float* vb = v + (size_t)n * S * tok + (size_t)h * D;
    const float* qb = q + (size_t)n * L * tok + (size_t)h * D;
    float* ob = out + (size_t)n * L * tok + (size_t)h * D;
    const float fS = (float)S;

    for (int e = threadIdx.x; e < D * D; e += LA_THREADS) {
        const int d = e / D, c = e % D;
        float acc = 0.f;
        for (int s = 0; s < S; ++s) {
            const float m = kv_mask ? kv_mask[(size_t)n * S + s] : 1.0f;
            const float kk = kv_mask ? elu1(kb[s * tok + d]) * m : elu1(kb[s * tok + d]);
            float vv = kv_mask ? vb[s * tok + c] * m : vb[s * tok + c];
            vv = __fdiv_rn(vv, fS);                          // values / v_length (linear_attention.py:40)
            acc = fmaf(kk, vv, acc);
        }
        KV[e] = acc;
    }
    for (int d = threadIdx.x; d < D; d += LA_THREADS) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) {
            const float kk = elu1(kb[s * tok + d]);
            acc += kv_mask ? kk * kv_mask[(size_t)n * S + s] : kk;
        }
        Ksum[d] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < L * D; e += LA_THREADS) {
        const int l = e / D, c = e % D;
        const float qm = q_mask ? q_mask[(size_t)n * L + l] : 1.0f;
        float num = 0.f, den = 0.f;
        for (int d = 0; d < D; ++d) {
            float qq = elu1(qb[l * tok + d]);
            if (q_mask) qq *= qm;
            num = fmaf(qq, KV[d * D + c], num);
            den = fmaf(qq, Ksum[d], den);
        }
        const float z = __fdiv_rn(1.0f, den + eps);
        ob[l * tok + c] = num * z * fS;
    }
}

}  // namespace riders

extern "C" int riders_linear_attention_fwd(const float* q, const float* k, const float* v,
                                           const float* q_mask, const float* kv_mask,
                                           int N, int L, int S, int H, int D, float eps,
                                           float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(N >= 0 && L > 0 && S > 0 && H > 0 && D > 0, "linear_attention: bad shape");
    RIDERS_REQUIRE(D <= LA_MAX_D, "linear_attention: head dim %d > %d", D, LA_MAX_D);
    if (N == 0) return RIDERS_OK;
    RIDERS_REQUIRE(q && k && v && out, "linear_attention: null pointer");
    RIDERS_REQUIRE((long long)N * H < (1ll << 31), "linear_attention: too many (sample, head) pairs");
    count_launches(1);
    linear_attention_kernel<<<N * H, LA_THREADS, 0, (cudaStream_t)stream>>>(q, k, v, q_mask, kv_mask, L, S, H, D, eps, out);
    return check_launch("linear_attention_fwd");
}
