// A8: Interpolator2D / interpolate_knots(method='linear') - modules/interpolator.py:7-50 of the
// reference, which delegates to scipy.interpolate.griddata (Qhull Delaunay + barycentric
// LinearNDInterpolator, fill_value=1.0 outside the hull, float64 math, float32 store).
//
//   extract_knots_kernel : row-major scan of `valid` -> knot (x, y), scale = sparse/pred,
//                          shift = sparse - pred            (interpolator.py:28-32)
//   delaunay_kernel      : one CTA per frame.  Sort knots by (x, y) (bitonic, shared memory),
//                          Guibas-Stolfi divide and conquer with exact integer predicates
//                          (delaunay_core.cuh): the nodes of one recursion depth are merged
//                          by independent threads, depth by depth; then emit triangle records.
//   interp_fill_kernel   : 32x32 pixel tiles; each tile culls the frame's triangles by bounding
//                          box, then every pixel evaluates exact integer edge functions (a pixel
//                          on the hull boundary is inside, like SciPy's eps test) and the fp64
//                          barycentric sum; no triangle -> `fill`.
//
// Roofline: the fill is HBM-write bound (4 B/pixel out); the triangulation is latency bound
// (pointer chasing) and only amortises over a batch of frames - reported separately.
#include "common.cuh"
#include "delaunay_core.cuh"
#include "ordered_compact.cuh"

namespace riders {

using dt::Pt;
using dt::Hull;
using dt::Mesh;

constexpr int DT_THREADS = 512;

struct TriRec {            // 48 bytes
    int ax, ay, bx, by, cx, cy;
    float va, vb, vc;
    int key_hi, key_lo;    // canonical order of the triangle (sorted knot ranks) for tie-breaking
    int pad;
};

struct FrameWs {           // device-side view of one frame's workspace slice
    unsigned long long* keys;
    Pt* pts;
    int* perm;
    int* next;
    int* org;
    Hull* hulls;
    int* counters;         // [0] fresh quads, [1] overflow, [2] triangles
    TriRec* tris;
};

struct WsLayout {
    size_t keys, pts, perm, next, org, hulls, counters, tris, total;
    int P, cap;
};

__host__ __device__ inline int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

__host__ __device__ inline WsLayout ws_layout(int max_knots) {
    WsLayout L;
    L.P = next_pow2(max_knots < 2 ? 2 : max_knots);
    L.cap = 8 * max_knots + 64;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    L.keys = take(sizeof(unsigned long long) * L.P);
    L.pts = take(sizeof(Pt) * (size_t)(max_knots + 1));
    L.perm = take(sizeof(int) * (size_t)(max_knots + 1));
    L.next = take(sizeof(int) * 4 * (size_t)L.cap);
    L.org = take(sizeof(int) * 4 * (size_t)L.cap);
    L.hulls = take(sizeof(Hull) * 4 * (size_t)L.P);
    L.counters = take(sizeof(int) * 8);
    L.tris = take(sizeof(TriRec) * 2 * (size_t)(max_knots + 1));
    L.total = o;
    return L;
}

__device__ inline FrameWs frame_ws(uint8_t* base, const WsLayout& L, int frame) {
    uint8_t* p = base + (size_t)frame * L.total;
    FrameWs w;
    w.keys = (unsigned long long*)(p + L.keys);
    w.pts = (Pt*)(p + L.pts);
    w.perm = (int*)(p + L.perm);
    w.next = (int*)(p + L.next);
    w.org = (int*)(p + L.org);
    w.hulls = (Hull*)(p + L.hulls);
    w.counters = (int*)(p + L.counters);
    w.tris = (TriRec*)(p + L.tris);
    return w;
}

// ---------------------------------------------------------------------------------------------
// knot extraction
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OC_THREADS)
extract_knots_kernel(const float* __restrict__ pred, const float* __restrict__ sparse,
                     const uint8_t* __restrict__ valid, int H, int W, int max_knots,
                     int32_t* __restrict__ knot_xy, float* __restrict__ knot_scale,
                     float* __restrict__ knot_shift, int32_t* __restrict__ knot_count) {
    const int frame = blockIdx.y;
    const int npx = H * W;
    const size_t fo = (size_t)frame * npx;
    const size_t ko = (size_t)frame * max_knots;
    const int upto = ordered_compact_slice(valid + fo, npx, [&](int pos, int i) {
        if (pos < max_knots) {
            const float p = pred[fo + i], s = sparse[fo + i];
            knot_xy[(ko + pos) * 2 + 0] = i % W;
            knot_xy[(ko + pos) * 2 + 1] = i / W;
            knot_scale[ko + pos] = __fdiv_rn(s, p);          // interpolator.py:31
            knot_shift[ko + pos] = __fsub_rn(s, p);          // interpolator.py:32
        }
    });
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) knot_count[frame] = upto;
}

struct SmemMesh { size_t next, org, pts, perm, hulls, total; };

__host__ __device__ inline SmemMesh smem_mesh_layout(int max_knots, const WsLayout& L) {
    SmemMesh S;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
    S.next = take(sizeof(int) * 4 * (size_t)L.cap);    // also holds the P sort keys (8 B each, P <= 2 * max_knots)
    S.org = take(sizeof(int) * 4 * (size_t)L.cap);
    S.pts = take(sizeof(Pt) * (size_t)(max_knots + 1));
    S.perm = take(sizeof(int) * (size_t)(max_knots + 1));
    S.hulls = take(sizeof(Hull) * 4 * (size_t)L.P);
    S.total = o;
    return S;
}

// ---------------------------------------------------------------------------------------------
// triangulation: one CTA per frame
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DT_THREADS)
delaunay_kernel(const int32_t* __restrict__ knot_xy, const float* __restrict__ knot_val,
                const int32_t* __restrict__ knot_count, int max_knots, uint8_t* ws_base, WsLayout L,
                int32_t* __restrict__ status, int32_t* __restrict__ tri_out, int32_t* __restrict__ tri_count, int mesh_in_smem) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ int s_counters[2];                       // [0] fresh quads, [1] overflow (shared-memory mesh only)
    const int frame = blockIdx.x;
    FrameWs w = frame_ws(ws_base, L, frame);
    int* fresh = &w.counters[0];
    int* overflow = &w.counters[1];
    if (mesh_in_smem) {
        // the pointer-chasing structures live in shared memory (~30-cycle instead of L2 latency per hop);
        // only the triangle records and their count go to the global workspace for the fill kernel
        const SmemMesh S = smem_mesh_layout(max_knots, L);
        w.next = (int*)(dsm + S.next); w.org = (int*)(dsm + S.org);
        w.keys = (unsigned long long*)(dsm + S.next);   // the sort finishes before the mesh is touched
        w.pts = (Pt*)(dsm + S.pts); w.perm = (int*)(dsm + S.perm); w.hulls = (Hull*)(dsm + S.hulls);
        fresh = &s_counters[0]; overflow = &s_counters[1];
        if (threadIdx.x < 2) s_counters[threadIdx.x] = 0;
    }
    const int n = min(knot_count[frame], max_knots);
    const int tid = threadIdx.x;
    const int32_t* xy = knot_xy + (size_t)frame * max_knots * 2;
    const float* val = knot_val + (size_t)frame * max_knots;

    if (tid < 8) w.counters[tid] = 0;
    if (knot_count[frame] > max_knots) {
        // more valid pixels than the caller's capacity: never triangulate a truncated knot set silently
        if (tid == 0) { status[frame] = RIDERS_EINTERNAL; if (tri_count) tri_count[frame] = 0; }
        return;
    }
    if (n < 3) {
        if (tid == 0) { status[frame] = RIDERS_EQHULL; if (tri_count) tri_count[frame] = 0; }
        return;
    }
    // ---- sort by (x, y): key = x:16 | y:16 | index:16 --------------------------------------
    const int P = next_pow2(n);
    for (int i = tid; i < P; i += DT_THREADS)
        w.keys[i] = (i < n) ? (((unsigned long long)(unsigned)xy[2 * i] << 32) | ((unsigned long long)(unsigned)xy[2 * i + 1] << 16) | (unsigned)i)
                            : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += DT_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = w.keys[i], b = w.keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { w.keys[i] = b; w.keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += DT_THREADS) {
        const unsigned long long key = w.keys[i];
        Pt p;
        p.x = (int)(key >> 32); p.y = (int)((key >> 16) & 0xffff);
        w.pts[i] = p;
        w.perm[i] = (int)(key & 0xffff);
    }
    __syncthreads();

    // ---- divide and conquer, deepest level first ----------------------------------------------
    Mesh m;
    m.next = w.next; m.org = w.org; m.fresh = fresh; m.capacity = L.cap; m.pts = w.pts; m.overflow = overflow;
    const int D = dt::tree_depth(n);
    for (int d = D; d >= 0; --d) {
        for (int j = tid; j < (1 << d); j += DT_THREADS) dt::process_node(m, n, d, j, w.hulls);
        __syncthreads();
    }

    // ---- triangle records ---------------------------------------------------------------------
    const int quads = min(*fresh, L.cap);
    for (int e2 = tid; e2 < 2 * quads; e2 += DT_THREADS) {
        const int e = 4 * (e2 >> 1) + 2 * (e2 & 1);
        int a, b, c;
        if (!dt::left_triangle(m, e, a, b, c)) continue;
        const int slot = atomicAdd(&w.counters[2], 1);
        if (slot >= 2 * max_knots) continue;
        TriRec t;
        t.ax = w.pts[a].x; t.ay = w.pts[a].y; t.bx = w.pts[b].x; t.by = w.pts[b].y; t.cx = w.pts[c].x; t.cy = w.pts[c].y;
        t.va = val[w.perm[a]]; t.vb = val[w.perm[b]]; t.vc = val[w.perm[c]];
        int s0 = a, s1 = b, s2 = c, tmp;                       // sorted ranks -> canonical key
        if (s0 > s1) { tmp = s0; s0 = s1; s1 = tmp; }
        if (s1 > s2) { tmp = s1; s1 = s2; s2 = tmp; }
        if (s0 > s1) { tmp = s0; s0 = s1; s1 = tmp; }
        t.key_hi = s0; t.key_lo = (s1 << 16) | s2;
        t.pad = 0;
        w.tris[slot] = t;
        if (tri_out) {
            int32_t* o = tri_out + ((size_t)frame * 2 * max_knots + slot) * 3;
            o[0] = w.perm[a]; o[1] = w.perm[b]; o[2] = w.perm[c];
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int nt = w.counters[2];
        status[frame] = *overflow ? RIDERS_EINTERNAL : (nt == 0 ? RIDERS_EQHULL : RIDERS_OK);
        if (tri_count) tri_count[frame] = nt;
    }
}

// ---------------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------------
constexpr int FILL_THREADS = 256;
constexpr int FT_W = 32, FT_H = 32, FILL_ROWS = FT_H / (FILL_THREADS / FT_W);   // 4 rows per thread

__global__ void __launch_bounds__(FILL_THREADS)
interp_fill_kernel(uint8_t* ws_base, WsLayout L, int max_knots, int H, int W, float fill, float* __restrict__ out) {
    __shared__ TriRec cand[FILL_THREADS];
    __shared__ int warp_count[FILL_THREADS / 32];
    __shared__ int n_cand;
    const int frame = blockIdx.z;
    const FrameWs w = frame_ws(ws_base, L, frame);
    const int nt = min(w.counters[2], 2 * max_knots);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * FT_W, ty0 = blockIdx.y * FT_H;
    const int px = tx0 + lane;
    const int py0 = ty0 + warp;                                  // rows py0 + 8*m

    float res[FILL_ROWS];
    unsigned long long best[FILL_ROWS];
#pragma unroll
    for (int mI = 0; mI < FILL_ROWS; ++mI) { res[mI] = fill; best[mI] = ~0ull; }

    for (int base = 0; base < nt; base += FILL_THREADS) {
        const int ti = base + tid;
        bool hit = false;
        TriRec t;
        if (ti < nt) {
            t = w.tris[ti];
            const int x0 = min(t.ax, min(t.bx, t.cx)), x1 = max(t.ax, max(t.bx, t.cx));
            const int y0 = min(t.ay, min(t.by, t.cy)), y1 = max(t.ay, max(t.by, t.cy));
            hit = x0 < tx0 + FT_W && x1 >= tx0 && y0 < ty0 + FT_H && y1 >= ty0;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_count[warp] = __popc(ballot);
        __syncthreads();
        int offset = 0;
#pragma unroll
        for (int q = 0; q < FILL_THREADS / 32; ++q) offset += (q < warp) ? warp_count[q] : 0;
        if (hit) cand[offset + __popc(ballot & ((1u << lane) - 1u))] = t;
        if (tid == FILL_THREADS - 1) n_cand = offset + __popc(ballot);
        __syncthreads();
        const int n = n_cand;
        for (int i = 0; i < n; ++i) {
            const TriRec& c = cand[i];
            const unsigned long long key = ((unsigned long long)(unsigned)c.key_hi << 32) | (unsigned)c.key_lo;
            // edge functions are affine in the pixel row: evaluate at py0 and step by 8 rows
            const int dxa = c.bx - px, dxb = c.cx - px, dxc = c.ax - px;
#pragma unroll
            for (int mI = 0; mI < FILL_ROWS; ++mI) {
                const int py = py0 + 8 * mI;
                const int wa = dxa * (c.cy - py) - (c.by - py) * dxb;     // orient(p, b, c)
                const int wb = dxb * (c.ay - py) - (c.cy - py) * dxc;     // orient(p, c, a)
                const int wc = dxc * (c.by - py) - (c.ay - py) * dxa;     // orient(p, a, b)
                if ((wa | wb | wc) >= 0 && key < best[mI]) {
                    const double det = (double)(wa + wb + wc);
                    const double ca = (double)wa / det, cb = (double)wb / det, cc = (double)wc / det;
                    res[mI] = (float)(ca * (double)c.va + cb * (double)c.vb + cc * (double)c.vc);
                    best[mI] = key;
                }
            }
        }
    }
    if (px < W) {
#pragma unroll
        for (int mI = 0; mI < FILL_ROWS; ++mI) {
            const int py = py0 + 8 * mI;
            if (py < H) out[((size_t)frame * H + py) * W + px] = res[mI];
        }
    }
}

}  // namespace riders

extern "C" size_t riders_interp_workspace_bytes(int B, int max_knots, int H, int W) {
    (void)H; (void)W;
    if (B <= 0 || max_knots <= 0) return 256;
    return (size_t)B * riders::ws_layout(max_knots).total + 256;
}

extern "C" int riders_extract_knots(const float* pred, const float* sparse, const uint8_t* valid,
                                    int B, int H, int W, int max_knots,
                                    int32_t* knot_xy, float* knot_scale, float* knot_shift, int32_t* knot_count,
                                    void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0 && max_knots > 0, "extract_knots: bad shape");
    RIDERS_REQUIRE((long long)H * W < (1ll << 30), "extract_knots: frame too large");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(pred && sparse && valid && knot_xy && knot_scale && knot_shift && knot_count, "extract_knots: null pointer");
    count_launches(1);
    dim3 grid(ceil_div(H * W, OC_SLICE), B);
    extract_knots_kernel<<<grid, OC_THREADS, 0, (cudaStream_t)stream>>>(pred, sparse, valid, H, W, max_knots,
                                                                     knot_xy, knot_scale, knot_shift, knot_count);
    return check_launch("extract_knots");
}

extern "C" int riders_interp_triangulate(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                                         int B, int max_knots, int32_t* status,
                                         int32_t* tri_out, int32_t* tri_count, void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0, "interp_triangulate: bad shape");
    RIDERS_REQUIRE(max_knots > 0 && max_knots <= 32768, "interp_triangulate: max_knots=%d out of range (1..32768)", max_knots);
    RIDERS_REQUIRE(B <= 65535, "interp_triangulate: at most 65535 frames per call");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(knot_xy && knot_val && knot_count && status && workspace, "interp_triangulate: null pointer");
    RIDERS_REQUIRE((((uintptr_t)workspace) & 255) == 0, "interp_triangulate: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const WsLayout L = ws_layout(max_knots);
    const SmemMesh S = smem_mesh_layout(max_knots, L);
    const int mesh_in_smem = S.total <= 200 * 1024 ? 1 : 0;               // ~680 knots; larger frames chase pointers in L2
    if (mesh_in_smem) {
        static size_t configured[64] = {0};
        if (int rc_smem = ensure_dynamic_smem(delaunay_kernel, (size_t)(S.total), configured, "interp_triangulate: smem attribute")) return rc_smem;
    }
    count_launches(1);
    delaunay_kernel<<<B, DT_THREADS, mesh_in_smem ? S.total : 0, st>>>(knot_xy, knot_val, knot_count, max_knots, (uint8_t*)workspace, L,
                                                                       status, tri_out, tri_count, mesh_in_smem);
    return check_launch("interp_triangulate");
}

extern "C" int riders_interp_fill(int B, int max_knots, int H, int W, float fill, float* out, void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0, "interp_fill: bad shape");
    RIDERS_REQUIRE(H <= 8192 && W <= 8192, "interp_fill: exact predicates need H, W <= 8192");
    RIDERS_REQUIRE(max_knots > 0 && max_knots <= 32768 && B <= 65535, "interp_fill: max_knots / B out of range");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(out && workspace, "interp_fill: null pointer");
    const WsLayout L = ws_layout(max_knots);
    dim3 grid(ceil_div(W, FT_W), ceil_div(H, FT_H), B);
    count_launches(1);
    interp_fill_kernel<<<grid, FILL_THREADS, 0, (cudaStream_t)stream>>>((uint8_t*)workspace, L, max_knots, H, W, fill, out);
    return check_launch("interp_fill");
}

extern "C" int riders_interp_linear(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                                    int B, int max_knots, int H, int W, float fill,
                                    float* out, int32_t* status, int32_t* tri_out, int32_t* tri_count,
                                    void* workspace, void* stream) {
    int rc = riders_interp_triangulate(knot_xy, knot_val, knot_count, B, max_knots, status, tri_out, tri_count, workspace, stream);
    if (rc != RIDERS_OK) return rc;
    return riders_interp_fill(B, max_knots, H, W, fill, out, workspace, stream);
}
