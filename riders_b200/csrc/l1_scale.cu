// N1: Optimizer.optimize_scale / apply_scale / clamp_min_max, modules/estimator.py:129-185
// of the reference - the DEFAULT alignment of every shipped script (val_zju.py:386 global_alignment='s').
//
// The reference minimises f(x) = sum_valid |x * pred - target| with scipy.optimize.minimize_scalar
// (method 'bounded': Brent's fminbound, xatol 1e-5, maxiter 500; third-party, published algorithm
// restated below) re-reading the full H x W maps ~30 times on the CPU.  Here:
//   compact_valid_kernel : row-major ordered compaction of the valid pixels' (pred, target) per frame
//   l1_brent_kernel      : one CTA per frame; every thread runs the same fminbound state machine, the
//                          objective is a deterministic fp64 block reduction over the compacted pairs
//   l1_apply_kernel      : out = fl32(pred * fl32(scale)) + the two clamps (estimator.py:161-176)
#include "common.cuh"
#include "ordered_compact.cuh"

namespace riders {

constexpr int BR_THREADS = 256;

__global__ void __launch_bounds__(OC_THREADS)
compact_valid_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                     const uint8_t* __restrict__ valid, int npx, int cap,
                     float2* __restrict__ pairs, int32_t* __restrict__ count, unsigned* aggregates) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * npx;
    float2* dst = pairs + (size_t)frame * cap;
    const int upto = ordered_compact_slice(valid + fo, npx, [&](int pos, int i) {
        if (pos < cap) dst[pos] = make_float2(pred[fo + i], target[fo + i]);
    }, aggregates ? aggregates + (size_t)frame * gridDim.x : nullptr);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) count[frame] = upto;
}

// deterministic block sum of |x*p - t| over the frame's pairs; every thread returns the same value
__device__ __forceinline__ double l1_objective(double x, const float2* __restrict__ pairs, int n, double* sm) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += BR_THREADS) {
        const float2 pt = pairs[i];
        a += fabs(x * (double)pt.x - (double)pt.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    __syncthreads();                                   // previous readers of sm[] are done
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < BR_THREADS / 32; ++w) s += sm[w];
    return s;
}

__device__ __forceinline__ double sgn_nz(double v) { return (v < 0.0) ? -1.0 : 1.0; }   // sign(v) + (v == 0)

__global__ void __launch_bounds__(BR_THREADS)
l1_brent_kernel(const float2* __restrict__ pairs_all, const int32_t* __restrict__ count, int cap,
                double x1, double x2, double xatol, int maxfun,
                double* __restrict__ scale, double* __restrict__ fval, int32_t* __restrict__ nfev) {
    __shared__ double sm[BR_THREADS / 32];
    const int frame = blockIdx.x;
    const float2* pairs = pairs_all + (size_t)frame * cap;
    const int n = min(count[frame], cap);
    // scipy.optimize._optimize._minimize_scalar_bounded, statement for statement, in fp64
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double a = x1, b = x2;
    double fulc = a + golden_mean * (b - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = l1_objective(x, pairs, n, sm);
    int num = 1;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        bool golden = true;
        if (fabs(e) > tol1) {                          // parabolic fit
            golden = false;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                rat = (p + 0.0) / q;
                x = xf + rat;
                if (((x - a) < tol2) || ((b - x) < tol2)) rat = tol1 * sgn_nz(xm - xf);
            } else {
                golden = true;
            }
        }
        if (golden) {
            e = (xf >= xm) ? a - xf : b - xf;
            rat = golden_mean * e;
        }
        x = xf + sgn_nz(rat) * fmax(fabs(rat), tol1);
        const double fu = l1_objective(x, pairs, n, sm);
        ++num;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fnfc) || (nfc == xf)) {
                fulc = nfc; ffulc = fnfc;
                nfc = x; fnfc = fu;
            } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
                fulc = x; ffulc = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= maxfun) break;
    }
    if (threadIdx.x == 0) {
        scale[frame] = xf;
        if (fval) fval[frame] = fx;
        if (nfev) nfev[frame] = num;
    }
}

__global__ void __launch_bounds__(256)
l1_apply_kernel(const float* __restrict__ pred, const double* __restrict__ scale, int npx,
                float* __restrict__ out, float clamp_hi, float clamp_lo, int clamp_flags) {
    const int frame = blockIdx.y;
    const float s = (float)scale[frame];               // estimate(float32) * scale, numpy 1.23 value-based casting
    const size_t fo = (size_t)frame * npx;
    const bool hi_on = clamp_flags & 1, lo_on = clamp_flags & 2;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < npx; i += gridDim.x * 256) {
        float r = __fmul_rn(pred[fo + i], s);
        if (hi_on && r > clamp_hi) r = clamp_hi;
        if (lo_on && r < clamp_lo) r = clamp_lo;
        out[fo + i] = r;
    }
}

}  // namespace riders

static size_t l1_pairs_bytes(int B, int max_valid) {
    return (((size_t)B * max_valid * sizeof(float2)) + 255) & ~(size_t)255;
}

extern "C" size_t riders_l1_scale_workspace_bytes(int B, int max_valid) {
    if (B <= 0 || max_valid <= 0) return 256;
    return l1_pairs_bytes(B, max_valid) + (size_t)B * sizeof(int32_t) + 256;
}

extern "C" int riders_l1_scale_align(const float* pred, const float* target, const uint8_t* mask,
                                     int B, int H, int W, int max_valid,
                                     double lo, double hi, double xatol, int maxiter,
                                     double* scale, double* fval, int32_t* nfev,
                                     float* out, float clamp_hi, float clamp_lo, int clamp_flags,
                                     void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0 && max_valid > 0, "l1_scale_align: bad shape");
    RIDERS_REQUIRE((long long)H * W < (1ll << 30) && B <= 65535, "l1_scale_align: frame too large / too many frames");
    RIDERS_REQUIRE(lo <= hi && xatol > 0 && maxiter > 0, "l1_scale_align: bad bounds / tolerance");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(pred && target && mask && scale && workspace, "l1_scale_align: null pointer");
    RIDERS_REQUIRE((((uintptr_t)workspace) & 255) == 0, "l1_scale_align: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int npx = H * W;
    float2* pairs = (float2*)workspace;
    int32_t* count = (int32_t*)((uint8_t*)workspace + l1_pairs_bytes(B, max_valid));
    count_launches(out ? 3 : 2);
    unsigned* agg = oc_scratch(B, ceil_div(npx, OC_SLICE), st);
    compact_valid_kernel<<<dim3(ceil_div(npx, OC_SLICE), B), OC_THREADS, 0, st>>>(pred, target, mask, npx, max_valid, pairs, count, agg);
    if (agg) cudaFreeAsync(agg, st);
    l1_brent_kernel<<<B, BR_THREADS, 0, st>>>(pairs, count, max_valid, lo, hi, xatol, maxiter, scale, fval, nfev);
    if (out) {
        dim3 grid(min(ceil_div(npx, 256 * 4), 1024), B);
        l1_apply_kernel<<<grid, 256, 0, st>>>(pred, scale, npx, out, clamp_hi, clamp_lo, clamp_flags);
    }
    return check_launch("l1_scale_align");
}
