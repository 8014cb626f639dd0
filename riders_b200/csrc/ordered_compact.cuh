// Row-major ordered compaction of the set bytes of a per-frame mask, many CTAs per frame.
// Used by extract_knots_kernel (A8, modules/interpolator.py:28-32: knots in np.nonzero order) and
// compact_valid_kernel (N1).  grid = (slices, frames); a CTA owns OC_SLICE consecutive pixels.
// Its output offset is the number of set bytes before its slice:
//   * with a zeroed `aggregates[frames][slices]` scratch (decoupled look-back): every CTA publishes the count of its own
//     slice (flagged) and adds up the published counts of the slices before it - it only ever waits for CTAs with a lower
//     block index, which are resident or finished by the time it runs.  The mask is read once.
//   * without scratch: the CTA counts the set bytes before its slice itself (re-reads <= npx bytes from L2; the round-1
//     form: the last slice of a 640x512 frame re-read 320 KB in 80 dependent steps, 34 us for one frame).
// Either way: one launch, no atomics on the output position, and the order is exactly the scan order.
#pragma once
#include "common.cuh"

namespace riders {

constexpr int OC_THREADS = 256;
constexpr int OC_PER_THREAD = 32;
constexpr int OC_SLICE = OC_THREADS * OC_PER_THREAD;      // 8192 pixels per CTA

__device__ __forceinline__ int nonzero_bytes(uint32_t w) {
    return __popc((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u);
}

__device__ __forceinline__ int block_sum_broadcast(int v, int* sm /* OC_THREADS/32 + 1 ints */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < OC_THREADS / 32; ++w) t += sm[w];
    __syncthreads();
    return t;
}

// emit(position, pixel_index) is called for every set byte of this CTA's slice, in order.
// Returns (to every thread) the number of set bytes in [0, end of this slice).
constexpr unsigned OC_READY = 0x80000000u;

// `aggregates`: this frame's row of the look-back scratch (gridDim.x ints, zero before the launch) or nullptr.
template <class Emit>
__device__ __forceinline__ int ordered_compact_slice(const uint8_t* __restrict__ v, int npx, Emit emit,
                                                     unsigned* aggregates = nullptr) {
    __shared__ int sm_red[OC_THREADS / 32 + 1];
    __shared__ int sm_warp[OC_THREADS / 32];
    const int begin = blockIdx.x * OC_SLICE;
    const bool vec = ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    // ---- this slice: 32 consecutive pixels per thread -------------------------------------------
    const int t0 = begin + threadIdx.x * OC_PER_THREAD;
    uint32_t flags = 0;
    if (vec && t0 + OC_PER_THREAD <= npx) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(v + t0));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(v + t0 + 16));
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if ((w[k] >> (8 * e)) & 0xffu) flags |= 1u << (4 * k + e);
    } else {
        for (int e = 0; e < OC_PER_THREAD; ++e)
            if (t0 + e < npx && v[t0 + e]) flags |= 1u << e;
    }
    const int cnt = __popc(flags);
    // ---- set bytes before this slice ----------------------------------------------------------
    int before = 0;
    if (aggregates) {
        const int mine = block_sum_broadcast(cnt, sm_red);
        if (threadIdx.x == 0) atomicExch(aggregates + blockIdx.x, (unsigned)mine | OC_READY);
        for (int j = threadIdx.x; j < (int)blockIdx.x; j += OC_THREADS) {
            unsigned a;
            do { a = *reinterpret_cast<volatile unsigned*>(aggregates + j); } while (!(a & OC_READY));
            before += (int)(a & ~OC_READY);
        }
    } else if (vec) {
        for (int i = threadIdx.x * 16; i < begin; i += OC_THREADS * 16) {       // begin is a multiple of 8192
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(v + i));
            before += nonzero_bytes(q.x) + nonzero_bytes(q.y) + nonzero_bytes(q.z) + nonzero_bytes(q.w);
        }
    } else {
        for (int i = threadIdx.x; i < begin; i += OC_THREADS) before += v[i] ? 1 : 0;
    }
    before = block_sum_broadcast(before, sm_red);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm_warp[warp] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < OC_THREADS / 32; ++w) { const int c = sm_warp[w]; if (w < warp) woff += c; total += c; }
    int pos = before + woff + inc - cnt;
    while (flags) {
        const int e = __ffs(flags) - 1;
        flags &= flags - 1;
        emit(pos++, t0 + e);
    }
    return before + total;
}

// Stream-ordered zeroed look-back scratch for one launch (frames x slices words); nullptr if the allocation is refused
// (the kernels then take the scratch-free form).  Free with cudaFreeAsync on the same stream after the launch.
inline unsigned* oc_scratch(int frames, int slices, cudaStream_t st) {
    // keep freed blocks cached in the device's default pool (its default threshold of 0 returns them to the driver at every
    // synchronisation, which made each cudaMallocAsync a 0.2 ms re-allocation)
    static bool pool_set[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !pool_set[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = 64ull << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        pool_set[dev] = true;
    }
    void* p = nullptr;
    const size_t bytes = (size_t)frames * slices * sizeof(unsigned);
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemsetAsync(p, 0, bytes, st) != cudaSuccess) { cudaGetLastError(); cudaFreeAsync(p, st); return nullptr; }
    return (unsigned*)p;
}

}  // namespace riders
