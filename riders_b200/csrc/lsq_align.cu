// A6/A7: global scale/shift least squares + apply + clamp (modules/estimator.py:5-29,108-125).
//
// Pass 1 (lsq_partial_kernel): every CTA owns a fixed, contiguous PX_PER_CTA-pixel (8192) slice of one
//   frame, reads the mask first and touches pred/target only where the mask is set, accumulates
//   the five normal-equation sums in fp64 and reduces them with a fixed shuffle/smem tree ->
//   partial[frame][cta][5].  The slice size does not depend on the SM count, so the result is
//   bit-reproducible on any device.
// Pass 2 (lsq_apply_kernel): every CTA re-adds its frame's partials in index order (same value
//   in every CTA), solves the 2x2 system in fp64 (det > 0 guard, estimator.py:22-27), and
//   streams out = fl32(fl32(pred*s) + t) with the two clamps.  CTA 0 of each frame publishes
//   sums / scale / shift.
//
// Default: pass 2 is a programmatic dependent launch (lsq_apply_pdl_kernel) that loads its `pred` pixels while pass 1 still
//   runs.  Also kept (mode 3): lsq_single_kernel - ONE ordinary launch with a per-frame flag barrier between the passes;
//   measured SLOWER than the PDL pair (26.6 vs 21.5 us at 16 frames, 80 vs 54 us at 64: the waiting CTAs hold the SM slots the
//   not-yet-started CTAs of later frames need, profiles/r02_microbench_frontend.jsonl).  Also kept: lsq_fused_kernel, both passes in ONE cooperative grid with a grid-wide barrier (measured SLOWER at
//   1 and 16 frames: the cooperative launch itself costs more than the second launch it saves - profiles/r02_microbench_*),
//   and the two plain launches.  All three use the same slices and summation order: bit-identical results.
//
// HBM roofline: algorithmic 13 B/pixel = pred + target (8) + mask (1) read, output (4) written.  Actual
// traffic is lower: pass 1 reads the mask (1 B/pixel) and touches pred / target only in the 32-byte sectors that hold a
// valid pixel; pass 2 reads pred once (4 B) and writes the output (4 B).
#include "common.cuh"
#include <cooperative_groups.h>
#include <atomic>

namespace riders {

// 8192-pixel slices (256 threads x 32 pixels): a 640x512 frame is 40 slices, 16 frames are 640 CTAs = ONE wave of the 148 SMs at
// <= 48 registers (round 1's 4096-pixel slices were 1280 CTAs = two waves for both passes; the slice size is a constant, so
// results stay bit-reproducible on any device)
constexpr int LSQ_THREADS = 256;
constexpr int PX_PER_THREAD = 32;
constexpr int PX_PER_CTA = LSQ_THREADS * PX_PER_THREAD;     // 8192
constexpr int PDL_PREFETCH = 4;                             // float4 per thread loaded before the dependency wait (16 pixels)

struct Sums5 { double v[5]; };

__device__ __forceinline__ void accumulate(Sums5& s, float pf, float tf) {
    const double p = (double)pf, t = (double)tf;
    s.v[0] += p * p;     // a00 = sum m*p*p   (products of two float32 are exact in fp64)
    s.v[1] += p;         // a01 = sum m*p
    s.v[2] += 1.0;       // a11 = sum m
    s.v[3] += p * t;     // b0  = sum m*p*t
    s.v[4] += t;         // b1  = sum m*t
}

__device__ __forceinline__ void block_reduce5(Sums5& s, double (*sm)[5]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s.v[q] += __shfl_down_sync(0xffffffffu, s.v[q], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 5; ++q) sm[warp][q] = s.v[q];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < LSQ_THREADS / 32; ++w) a += sm[w][threadIdx.x];
        sm[0][threadIdx.x] = a;
    }
    __syncthreads();
}

__device__ __forceinline__ void partial_slice(const float* __restrict__ pred, const float* __restrict__ target,
                                              const uint8_t* __restrict__ mask, int npx, int ctas_per_frame,
                                              double* __restrict__ partial, int frame, int cta, double (*sm)[5]) {
    const size_t fo = (size_t)frame * npx;
    const float* p = pred + fo;
    const float* t = target + fo;
    const uint8_t* m = mask + fo;
    const int begin = cta * PX_PER_CTA;
    const int end = min(begin + PX_PER_CTA, npx);
    Sums5 s = {{0, 0, 0, 0, 0}};

    const bool vec_ok = ((npx & 3) == 0) && ((((uintptr_t)m) & 3) == 0) &&
                        ((((uintptr_t)p) & 15) == 0) && ((((uintptr_t)t) & 15) == 0);
    if (vec_ok) {
#pragma unroll
        for (int it = 0; it < PX_PER_THREAD / 4; ++it) {
            const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
            if (i < end) {
                const uint32_t mm = *reinterpret_cast<const uint32_t*>(m + i);
                if (mm) {
                    const float4 pv = *reinterpret_cast<const float4*>(p + i);
                    const float4 tv = *reinterpret_cast<const float4*>(t + i);
                    if (mm & 0x000000ffu) accumulate(s, pv.x, tv.x);
                    if (mm & 0x0000ff00u) accumulate(s, pv.y, tv.y);
                    if (mm & 0x00ff0000u) accumulate(s, pv.z, tv.z);
                    if (mm & 0xff000000u) accumulate(s, pv.w, tv.w);
                }
            }
        }
    } else {
        // same pixel -> thread assignment as the vector path, so the sums are layout-independent
        for (int it = 0; it < PX_PER_THREAD / 4; ++it) {
            const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
            for (int e = 0; e < 4; ++e)
                if (i + e < end && m[i + e]) accumulate(s, p[i + e], t[i + e]);
        }
    }
    block_reduce5(s, sm);
    if (threadIdx.x < 5) {
        partial[((size_t)frame * ctas_per_frame + cta) * 5 + threadIdx.x] = sm[0][threadIdx.x];
        __threadfence();                               // device-visible before anything this CTA signals after the barrier below
    }
    __syncthreads();                                   // sm is reused by the next slice of a persistent CTA
}

__global__ void __launch_bounds__(LSQ_THREADS)
lsq_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                   const uint8_t* __restrict__ mask, int npx, int ctas_per_frame,
                   double* __restrict__ partial) {
    __shared__ double sm[LSQ_THREADS / 32][5];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a programmatic dependent pass 2 may become resident now
    partial_slice(pred, target, mask, npx, ctas_per_frame, partial, blockIdx.y, blockIdx.x, sm);
}

__device__ __forceinline__ void solve_frame(const double* partial, int ctas_per_frame,
                                            double* sm5, double* sm_ss) {
    // fixed-order sum of the per-CTA partials (identical in every CTA of the frame):
    // warp q adds sum q; lane l takes CTAs l, l+32, ... in order, then a fixed shuffle tree
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    if (q < 5) {
        double a = 0.0;
        for (int c = lane; c < ctas_per_frame; c += 32) a += __ldcg(partial + (size_t)c * 5 + q);   // L2: written by other CTAs of this launch in the fused form
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0) sm5[q] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double a00 = sm5[0], a01 = sm5[1], a11 = sm5[2], b0 = sm5[3], b1 = sm5[4];
        const double det = a00 * a11 - a01 * a01;                 // estimator.py:22
        double x0 = 0.0, x1 = 0.0;
        if (det > 0.0) {                                           // estimator.py:24
            x0 = (a11 * b0 - a01 * b1) / det;
            x1 = (-a01 * b0 + a00 * b1) / det;
        }
        sm_ss[0] = x0;
        sm_ss[1] = x1;
    }
    __syncthreads();
}

__device__ __forceinline__ void apply_slice(const float* __restrict__ pred, const double* partial,
                                            int npx, int ctas_per_frame,
                                            double* __restrict__ sums, double* __restrict__ scale_shift,
                                            float* __restrict__ out, float clamp_hi, float clamp_lo, int clamp_flags,
                                            int frame, int cta, double* sm5, double* sm_ss) {
    solve_frame(partial + (size_t)frame * ctas_per_frame * 5, ctas_per_frame, sm5, sm_ss);
    if (cta == 0) {
        if (sums && threadIdx.x < 5) sums[frame * 5 + threadIdx.x] = sm5[threadIdx.x];
        if (scale_shift && threadIdx.x < 2) scale_shift[frame * 2 + threadIdx.x] = sm_ss[threadIdx.x];
    }
    if (!out) return;
    const float s = (float)sm_ss[0], t = (float)sm_ss[1];          // 0-d float32 in the reference
    const size_t fo = (size_t)frame * npx;
    const float* p = pred + fo;
    float* o = out + fo;
    const int begin = cta * PX_PER_CTA;
    const int end = min(begin + PX_PER_CTA, npx);
    const bool hi_on = clamp_flags & 1, lo_on = clamp_flags & 2;
    auto f = [&](float v) {
        float r = __fadd_rn(__fmul_rn(v, s), t);                   // two roundings, no FMA (estimator.py:113)
        if (hi_on && r > clamp_hi) r = clamp_hi;
        if (lo_on && r < clamp_lo) r = clamp_lo;
        return r;
    };
    const bool vec_ok = ((npx & 3) == 0) && ((((uintptr_t)p) & 15) == 0) && ((((uintptr_t)o) & 15) == 0);
#pragma unroll
    for (int it = 0; it < PX_PER_THREAD / 4; ++it) {
        const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
        if (vec_ok) {
            if (i < end) {
                float4 v = *reinterpret_cast<const float4*>(p + i);
                v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
                *reinterpret_cast<float4*>(o + i) = v;
            }
        } else {
            for (int e = 0; e < 4; ++e)
                if (i + e < end) o[i + e] = f(p[i + e]);
        }
    }
}

__global__ void __launch_bounds__(LSQ_THREADS)
lsq_apply_kernel(const float* __restrict__ pred, const double* __restrict__ partial,
                 int npx, int ctas_per_frame,
                 double* __restrict__ sums, double* __restrict__ scale_shift,
                 float* __restrict__ out, float clamp_hi, float clamp_lo, int clamp_flags) {
    __shared__ double sm5[5];
    __shared__ double sm_ss[2];
    apply_slice(pred, partial, npx, ctas_per_frame, sums, scale_shift, out, clamp_hi, clamp_lo, clamp_flags,
                blockIdx.y, blockIdx.x, sm5, sm_ss);
}

// Pass 2 as a PROGRAMMATIC DEPENDENT LAUNCH (the default): launched with programmatic stream serialisation it becomes
// resident while pass 1 is still running, loads its 16 pixels of `pred` per thread (they do not depend on pass 1), and only
// then waits for pass 1's partials (griddepcontrol.wait) - the HBM read of pass 2 and the launch latency of the second kernel
// hide under pass 1.  Same arithmetic as lsq_apply_kernel, bit for bit.
__global__ void __launch_bounds__(LSQ_THREADS, 5)
lsq_apply_pdl_kernel(const float* __restrict__ pred, const double* partial,
                     int npx, int ctas_per_frame,
                     double* __restrict__ sums, double* __restrict__ scale_shift,
                     float* __restrict__ out, float clamp_hi, float clamp_lo, int clamp_flags) {
    __shared__ double sm5[5];
    __shared__ double sm_ss[2];
    const int frame = blockIdx.y, cta = blockIdx.x;
    const size_t fo = (size_t)frame * npx;
    const float* p = pred + fo;
    float* o = out ? out + fo : nullptr;
    const int begin = cta * PX_PER_CTA;
    const int end = min(begin + PX_PER_CTA, npx);
    const bool vec_ok = out && ((npx & 3) == 0) && ((((uintptr_t)p) & 15) == 0) && ((((uintptr_t)o) & 15) == 0);
    auto load4 = [&](int it) {
        const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec_ok) { if (i < end) v = *reinterpret_cast<const float4*>(p + i); }
        else {
            if (i < end) v.x = p[i];
            if (i + 1 < end) v.y = p[i + 1];
            if (i + 2 < end) v.z = p[i + 2];
            if (i + 3 < end) v.w = p[i + 3];
        }
        return v;
    };
    float4 v[PDL_PREFETCH];
    if (out) {
#pragma unroll
        for (int it = 0; it < PDL_PREFETCH; ++it) v[it] = load4(it);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");            // pass 1 has completed and its partials are visible
    solve_frame(partial + (size_t)frame * ctas_per_frame * 5, ctas_per_frame, sm5, sm_ss);
    if (cta == 0) {
        if (sums && threadIdx.x < 5) sums[frame * 5 + threadIdx.x] = sm5[threadIdx.x];
        if (scale_shift && threadIdx.x < 2) scale_shift[frame * 2 + threadIdx.x] = sm_ss[threadIdx.x];
    }
    if (!out) return;
    const float s = (float)sm_ss[0], t = (float)sm_ss[1];
    const bool hi_on = clamp_flags & 1, lo_on = clamp_flags & 2;
    auto f = [&](float x) {
        float r = __fadd_rn(__fmul_rn(x, s), t);
        if (hi_on && r > clamp_hi) r = clamp_hi;
        if (lo_on && r < clamp_lo) r = clamp_lo;
        return r;
    };
    auto store4 = [&](int it, float4 x) {
        const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
        if (vec_ok) {
            if (i < end) *reinterpret_cast<float4*>(o + i) = make_float4(f(x.x), f(x.y), f(x.z), f(x.w));
        } else {
            if (i < end) o[i] = f(x.x);
            if (i + 1 < end) o[i + 1] = f(x.y);
            if (i + 2 < end) o[i + 2] = f(x.z);
            if (i + 3 < end) o[i + 3] = f(x.w);
        }
    };
    float4 w[PX_PER_THREAD / 4 - PDL_PREFETCH];
#pragma unroll
    for (int it = PDL_PREFETCH; it < PX_PER_THREAD / 4; ++it) w[it - PDL_PREFETCH] = load4(it);    // the rest streams while the first part is written
#pragma unroll
    for (int it = 0; it < PDL_PREFETCH; ++it) store4(it, v[it]);
#pragma unroll
    for (int it = PDL_PREFETCH; it < PX_PER_THREAD / 4; ++it) store4(it, w[it - PDL_PREFETCH]);
}

// Both passes in ONE ordinary launch (mode 3): the grid of pass 1, a per-frame flag barrier, then solve + apply.  A frame's CTAs
// only wait for the CTAs of the SAME frame: every CTA publishes its partial sums followed by a flag word carrying this call's
// nonce (st.release; the workspace needs no zeroing - stale memory cannot hold the nonce), polls the flags of its frame
// (ld.acquire), and goes on.  No second launch, no programmatic dependency, no cooperative-launch cost (but see the
// measurement in the file header: it loses to the PDL pair); `pred` is prefetched
// before the wait like the PDL form does.  Deadlock-free as long as one frame's CTAs fit the device together (the host
// checks; CTAs are dispatched frame by frame, earlier frames never wait for later ones).  Same slices, thread -> pixel
// assignment and summation order as the other forms: bit-identical results.
__global__ void __launch_bounds__(LSQ_THREADS, 5)
lsq_single_kernel(const float* __restrict__ pred, const float* __restrict__ target, const uint8_t* __restrict__ mask,
                  int npx, int ctas_per_frame, double* partial, unsigned long long* flags, unsigned long long nonce,
                  double* __restrict__ sums, double* __restrict__ scale_shift,
                  float* __restrict__ out, float clamp_hi, float clamp_lo, int clamp_flags) {
    __shared__ double sm[LSQ_THREADS / 32][5];
    __shared__ double sm5[5];
    __shared__ double sm_ss[2];
    const int frame = blockIdx.y, cta = blockIdx.x;
    const size_t fo = (size_t)frame * npx;
    const float* p = pred + fo;
    float* o = out ? out + fo : nullptr;
    const int begin = cta * PX_PER_CTA;
    const int end = min(begin + PX_PER_CTA, npx);
    const bool vec_ok = out && ((npx & 3) == 0) && ((((uintptr_t)p) & 15) == 0) && ((((uintptr_t)o) & 15) == 0);
    auto load4 = [&](int it) {
        const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec_ok) { if (i < end) v = *reinterpret_cast<const float4*>(p + i); }
        else {
            if (i < end) v.x = p[i];
            if (i + 1 < end) v.y = p[i + 1];
            if (i + 2 < end) v.z = p[i + 2];
            if (i + 3 < end) v.w = p[i + 3];
        }
        return v;
    };
    float4 v[PDL_PREFETCH];
    if (out) {
#pragma unroll
        for (int it = 0; it < PDL_PREFETCH; ++it) v[it] = load4(it);
    }
    partial_slice(pred, target, mask, npx, ctas_per_frame, partial, frame, cta, sm);
    unsigned long long* fl = flags + (size_t)frame * ctas_per_frame;
    if (threadIdx.x == 0) {
        __threadfence();                                           // the five partial sums (written by threads 0..4) before the flag
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(fl + cta), "l"(nonce) : "memory");
    }
    if (!out && cta != 0) return;                                  // sums / scale only: one CTA per frame finishes the job
    for (int c = threadIdx.x; c < ctas_per_frame; c += LSQ_THREADS) {
        unsigned long long f;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(f) : "l"(fl + c) : "memory");
        } while (f != nonce);
    }
    __syncthreads();
    solve_frame(partial + (size_t)frame * ctas_per_frame * 5, ctas_per_frame, sm5, sm_ss);
    if (cta == 0) {
        if (sums && threadIdx.x < 5) sums[frame * 5 + threadIdx.x] = sm5[threadIdx.x];
        if (scale_shift && threadIdx.x < 2) scale_shift[frame * 2 + threadIdx.x] = sm_ss[threadIdx.x];
    }
    if (!out) return;
    const float sc = (float)sm_ss[0], sh = (float)sm_ss[1];
    const bool hi_on = clamp_flags & 1, lo_on = clamp_flags & 2;
    auto f = [&](float x) {
        float r = __fadd_rn(__fmul_rn(x, sc), sh);
        if (hi_on && r > clamp_hi) r = clamp_hi;
        if (lo_on && r < clamp_lo) r = clamp_lo;
        return r;
    };
    auto store4 = [&](int it, float4 x) {
        const int i = begin + (it * LSQ_THREADS + threadIdx.x) * 4;
        if (vec_ok) {
            if (i < end) *reinterpret_cast<float4*>(o + i) = make_float4(f(x.x), f(x.y), f(x.z), f(x.w));
        } else {
            if (i < end) o[i] = f(x.x);
            if (i + 1 < end) o[i + 1] = f(x.y);
            if (i + 2 < end) o[i + 2] = f(x.z);
            if (i + 3 < end) o[i + 3] = f(x.w);
        }
    };
    float4 w[PX_PER_THREAD / 4 - PDL_PREFETCH];
#pragma unroll
    for (int it = PDL_PREFETCH; it < PX_PER_THREAD / 4; ++it) w[it - PDL_PREFETCH] = load4(it);    // the rest streams while the first part is written
#pragma unroll
    for (int it = 0; it < PDL_PREFETCH; ++it) store4(it, v[it]);
#pragma unroll
    for (int it = PDL_PREFETCH; it < PX_PER_THREAD / 4; ++it) store4(it, w[it - PDL_PREFETCH]);
}

// Both passes in one cooperative launch: every CTA walks the (frame, slice) items with the grid stride, the grid barrier
// separates "all partials written" from "solve + apply".  Slices, thread -> pixel assignment and summation order are those
// of the two kernels above, so the outputs are bit-identical to theirs on any grid size.
__global__ void __launch_bounds__(LSQ_THREADS)
lsq_fused_kernel(const float* __restrict__ pred, const float* __restrict__ target, const uint8_t* __restrict__ mask,
                 int B, int npx, int ctas_per_frame, double* partial,
                 double* __restrict__ sums, double* __restrict__ scale_shift,
                 float* __restrict__ out, float clamp_hi, float clamp_lo, int clamp_flags) {
    __shared__ double sm[LSQ_THREADS / 32][5];
    __shared__ double sm5[5];
    __shared__ double sm_ss[2];
    const int items = B * ctas_per_frame;
    for (int it = blockIdx.x; it < items; it += gridDim.x)
        partial_slice(pred, target, mask, npx, ctas_per_frame, partial, it / ctas_per_frame, it % ctas_per_frame, sm);
    __threadfence();
    cooperative_groups::this_grid().sync();
    if (out) {
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            apply_slice(pred, partial, npx, ctas_per_frame, sums, scale_shift, out, clamp_hi, clamp_lo, clamp_flags,
                        it / ctas_per_frame, it % ctas_per_frame, sm5, sm_ss);
            __syncthreads();                           // sm5 / sm_ss are rewritten by the next item
        }
    } else {
        for (int frame = blockIdx.x; frame < B; frame += gridDim.x) {
            apply_slice(pred, partial, npx, ctas_per_frame, sums, scale_shift, nullptr, clamp_hi, clamp_lo, clamp_flags,
                        frame, 0, sm5, sm_ss);
            __syncthreads();
        }
    }
}

static int g_lsq_mode = 2;                             // 2: two kernels, the second a programmatic dependent launch (default);
                                                       // 3: ONE ordinary launch with a per-frame flag barrier (measured slower);
                                                       // 1: one cooperative launch; 0: two plain kernels
void set_lsq_mode(int m) { g_lsq_mode = m; }

}  // namespace riders

extern "C" size_t riders_lsq_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t ctas = (size_t)riders::ceil_div(H * W, riders::PX_PER_CTA);
    return (size_t)B * ctas * (5 * sizeof(double) + sizeof(unsigned long long));        // partial sums + the flag words of mode 3
}

extern "C" int riders_lsq_align(const float* pred, const float* target, const uint8_t* mask,
                                int B, int H, int W, double* sums, double* scale_shift,
                                float* out, float clamp_hi, float clamp_lo, int clamp_flags,
                                void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0, "lsq_align: bad shape B=%d H=%d W=%d", B, H, W);
    RIDERS_REQUIRE((long long)H * W < (1ll << 30), "lsq_align: frame too large");
    RIDERS_REQUIRE(B <= 65535, "lsq_align: at most 65535 frames per call");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(pred && target && mask && workspace, "lsq_align: null pointer");
    RIDERS_REQUIRE((((uintptr_t)workspace) & 7) == 0, "lsq_align: workspace must be 8-byte aligned");
    const int npx = H * W;
    const int ctas = ceil_div(npx, PX_PER_CTA);
    cudaStream_t st = (cudaStream_t)stream;
    if (g_lsq_mode == 1) {
        // co-resident capacity of this device for the fused kernel (cached per device)
        static int capacity[64] = {0};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        if (!capacity[dev]) {
            int per_sm = 0, coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
            if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lsq_fused_kernel, LSQ_THREADS, 0) == cudaSuccess && per_sm > 0)
                capacity[dev] = per_sm * sm_count();
            else
                capacity[dev] = -1;
            cudaGetLastError();
        }
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (capacity[dev] > 0 && cap == cudaStreamCaptureStatusNone) {
            const int items = B * ctas;
            int grid1 = items < capacity[dev] ? items : capacity[dev];
            int npx_ = npx, ctas_ = ctas, B_ = B;
            double* part = (double*)workspace;
            void* args[] = {(void*)&pred, (void*)&target, (void*)&mask, (void*)&B_, (void*)&npx_, (void*)&ctas_, (void*)&part,
                            (void*)&sums, (void*)&scale_shift, (void*)&out, (void*)&clamp_hi, (void*)&clamp_lo, (void*)&clamp_flags};
            if (cudaLaunchCooperativeKernel((const void*)lsq_fused_kernel, dim3(grid1), dim3(LSQ_THREADS), args, 0, st) == cudaSuccess) {
                count_launches(1);
                return check_launch("lsq_align (fused)");
            }
            cudaGetLastError();                        // refused (e.g. MPS / partitioned device): take the two-kernel form
        }
    }
    dim3 grid(ctas, B);
    if (g_lsq_mode == 3) {
        // one frame's CTAs must be able to be resident together (they wait for each other): cached per device
        static int capacity[64] = {0};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        if (!capacity[dev]) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lsq_single_kernel, LSQ_THREADS, 0) == cudaSuccess && per_sm > 0)
                capacity[dev] = per_sm * sm_count();
            else
                capacity[dev] = -1;
            cudaGetLastError();
        }
        // a captured graph would replay the same nonce on the same workspace: capture takes the two-kernel form
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (capacity[dev] > 0 && ctas <= capacity[dev] / 4 && cap == cudaStreamCaptureStatusNone) {
            static std::atomic<unsigned long long> nonce{0x9E3779B97F4A7C15ull ^ (unsigned long long)(uintptr_t)&capacity};
            const unsigned long long my = nonce.fetch_add(0x9E3779B97F4A7C15ull, std::memory_order_relaxed) | 1ull;
            double* part = (double*)workspace;
            unsigned long long* flags = (unsigned long long*)(part + (size_t)B * ctas * 5);
            count_launches(1);
            lsq_single_kernel<<<grid, LSQ_THREADS, 0, st>>>(pred, target, mask, npx, ctas, part, flags, my, sums, scale_shift, out,
                                                            clamp_hi, clamp_lo, clamp_flags);
            return check_launch("lsq_align (single launch)");
        }
    }
    count_launches(2);
    lsq_partial_kernel<<<grid, LSQ_THREADS, 0, st>>>(pred, target, mask, npx, ctas, (double*)workspace);
    dim3 grid2(out ? ctas : 1, B);
    if (g_lsq_mode >= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid2; cfg.blockDim = dim3(LSQ_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const double* part = (const double*)workspace;
        if (cudaLaunchKernelEx(&cfg, lsq_apply_pdl_kernel, pred, part, npx, ctas, sums, scale_shift, out, clamp_hi, clamp_lo,
                               clamp_flags) == cudaSuccess)
            return check_launch("lsq_align (pdl)");
        cudaGetLastError();                            // attribute refused: plain second launch below
    }
    lsq_apply_kernel<<<grid2, LSQ_THREADS, 0, st>>>(pred, (const double*)workspace, npx, ctas, sums, scale_shift,
                                                    out, clamp_hi, clamp_lo, clamp_flags);
    return check_launch("lsq_align");
}
