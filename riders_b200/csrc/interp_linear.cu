// A8: Interpolator2D / interpolate_knots(method='linear') - modules/interpolator.py:7-50 of the
// reference, which delegates to scipy.interpolate.griddata (Qhull Delaunay + barycentric
// LinearNDInterpolator, fill_value=1.0 outside the hull, float64 math, float32 store).
//
//   extract_knots_kernel : row-major scan of `valid` -> knot (x, y), scale = sparse/pred,
//                          shift = sparse - pred            (interpolator.py:28-32)
//   delaunay_star_kernel : ONE WARP PER KNOT builds that knot's Delaunay star (delaunay_star.cuh: exact int64 predicates,
//                          symbolic perturbation for co-circular quads); candidates are visited outwards in knot index
//                          (= row-major) order, 32 at a time against a read-only star, and the walk stops once the row
//                          distance on both sides exceeds twice the largest circumradius of the closed star (no farther knot
//                          can lie inside any of its circumcircles).  Every triangle is emitted once, by its lowest-index
//                          knot.  The same launch writes the fill value over the output maps (the rasteriser overwrites the
//                          covered pixels), and the last CTA of a frame publishes status / triangle count.
//   raster_kernel        : scatter form of the barycentric fill: 8 triangles per CTA, their bounding-box rows dealt
//                          round-robin to the CTA's warps; a pixel is written iff its three exact integer edge functions are >= 0 (a pixel on
//                          the hull boundary is inside, like SciPy's eps test).  Pixels on an edge or a knot are evaluated
//                          from the edge's two knots in canonical order, so the two triangles sharing the edge write
//                          identical bits (benign race, deterministic image); that path is exact at the knots.  Interior
//                          pixels evaluate the triangle's affine interpolant va + gx (x - ax) + gy (y - ay) in fp64 (the
//                          barycentric sum written as a plane: two fused multiply-adds per pixel instead of a division).
//
// Round 1 ran a divide-and-conquer triangulation with ONE CTA per frame (0.24 ms for one 200-knot frame, 0.66 ms for 16
// frames, a cliff past 680 knots) and a gather fill that tested ~15 bounding-box candidates per pixel; the star form uses
// the whole GPU for a single frame and the scatter fill tests ~2 pixels per covered pixel.
//
// Roofline: algorithmic bytes are the 4 B/pixel of the output map (+ 12 B per knot); the launches write the map twice
// (fill value, then the covered pixels) - 8 B/pixel of traffic.
#include "common.cuh"
#include "delaunay_star.cuh"
#include "ordered_compact.cuh"

namespace riders {

using dt::Pt;

struct TriRec {            // 48 bytes
    int ax, ay, bx, by, cx, cy;
    float va, vb, vc;
    int ia, ib, ic;        // knot indices (row-major rank for extracted knots)
};

// per-frame counters in the workspace
enum { CNT_TRIS = 0, CNT_OVERFLOW = 1, CNT_DONE = 2, CNT_NEXT = 3, CNT_STRIDE = 8 };

struct WsLayout { size_t counters, tris, total; };

__host__ __device__ inline WsLayout ws_layout(int B, int max_knots) {
    WsLayout L;
    L.counters = 0;
    L.tris = ((size_t)B * CNT_STRIDE * sizeof(int) + 255) & ~(size_t)255;
    L.total = L.tris + (size_t)B * 2 * (size_t)max_knots * sizeof(TriRec);
    return L;
}

// ---------------------------------------------------------------------------------------------
// knot extraction
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OC_THREADS)
extract_knots_kernel(const float* __restrict__ pred, const float* __restrict__ sparse,
                     const uint8_t* __restrict__ valid, int H, int W, int max_knots,
                     int32_t* __restrict__ knot_xy, float* __restrict__ knot_scale,
                     float* __restrict__ knot_shift, int32_t* __restrict__ knot_count, unsigned* aggregates) {
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
    }, aggregates ? aggregates + (size_t)frame * gridDim.x : nullptr);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) knot_count[frame] = upto;
}

// ---------------------------------------------------------------------------------------------
// triangulation: one warp per knot
// ---------------------------------------------------------------------------------------------
constexpr int STAR_WARPS = 8;
constexpr int STAR_THREADS = STAR_WARPS * 32;
constexpr int STAR_LANES = 32;              // device stars live in registers, one neighbour per lane: at most 32 neighbours
                                            // (random knot sets stay below 20; more is reported as RIDERS_EINTERNAL)
#define FULL 0xffffffffu

// The warp-parallel form of star::accepts / star::insert (delaunay_star.cuh holds the scalar reference the host tests run):
// lane j < k holds neighbour j (ex, ey relative to p, eid its knot index), counter-clockwise, cyclic.
struct WarpStar {
    int ex, ey, eid, k;

    // every lane tests ITS OWN candidate (ux, uy, rid) against the read-only star
    __device__ __forceinline__ bool accepts_each(int p, int ux, int uy, int rid) const {
        if (k == 0) return true;
        int ax = 0, ay = 0, aid = 0, bx = 0, by = 0, bid = 0;
        bool found = false;
        int cx = __shfl_sync(FULL, ex, 0), cy = __shfl_sync(FULL, ey, 0), cid = __shfl_sync(FULL, eid, 0);
        for (int j = 0; j < k; ++j) {                               // uniform trip count: no divergence around the shuffles
            const int jn = j + 1 == k ? 0 : j + 1;
            const int nx = __shfl_sync(FULL, ex, jn), ny = __shfl_sync(FULL, ey, jn), nid = __shfl_sync(FULL, eid, jn);
            if (!found && (k == 1 || star::in_wedge(cx, cy, nx, ny, ux, uy))) {
                found = true; ax = cx; ay = cy; aid = cid; bx = nx; by = ny; bid = nid;
            }
            cx = nx; cy = ny; cid = nid;
        }
        if (star::cross2(ax, ay, ux, uy) == 0 && star::dot2(ax, ay, ux, uy) > 0)             // same ray: the nearer one only
            return star::dot2(ux, uy, ux, uy) < star::dot2(ax, ay, ax, ay);
        if (k == 1) return true;
        if (star::cross2(ax, ay, bx, by) <= 0) return true;                                  // exterior wedge of a hull knot
        return star::inside_sos(ax, ay, bx, by, ux, uy, rid, p, aid, bid);
    }

    __device__ __forceinline__ void get(int j, int& x, int& y, int& id) const {
        x = __shfl_sync(FULL, ex, j); y = __shfl_sync(FULL, ey, j); id = __shfl_sync(FULL, eid, j);
    }

    // the whole warp inserts ONE candidate (warp-uniform arguments): 0 rejected, 1 inserted, -1 more than STAR_LANES neighbours
    __device__ __forceinline__ int insert(int p, int ux, int uy, int rid, int lane) {
        if (k == 0) { if (lane == 0) { ex = ux; ey = uy; eid = rid; } k = 1; return 1; }
        int j = 0;
        if (k >= 2) {
            const int ln = lane + 1 == k ? 0 : lane + 1;
            const int nx = __shfl_sync(FULL, ex, ln), ny = __shfl_sync(FULL, ey, ln);
            const unsigned hit = __ballot_sync(FULL, lane < k && star::in_wedge(ex, ey, nx, ny, ux, uy));
            j = hit ? __ffs(hit) - 1 : 0;
        }
        int ax, ay, aid;
        get(j, ax, ay, aid);
        const bool same_ray = star::cross2(ax, ay, ux, uy) == 0 && star::dot2(ax, ay, ux, uy) > 0;
        if (same_ray) {
            if (!(star::dot2(ux, uy, ux, uy) < star::dot2(ax, ay, ax, ay))) return 0;
        } else if (k >= 2) {
            int bx, by, bid;
            get(j + 1 == k ? 0 : j + 1, bx, by, bid);
            if (star::cross2(ax, ay, bx, by) > 0 && !star::inside_sos(ax, ay, bx, by, ux, uy, rid, p, aid, bid)) return 0;
        }
        int first = j + 1 == k ? 0 : j + 1;
        int kept = same_ray ? k - 1 : k;
        while (kept >= 2) {                                          // counter-clockwise side of the cavity
            int bx, by, bid, cx, cy, cid;
            get(first, bx, by, bid);
            get(first + 1 == k ? 0 : first + 1, cx, cy, cid);
            if (star::cross2(ux, uy, bx, by) <= 0 || star::cross2(bx, by, cx, cy) <= 0) break;
            if (!star::inside_sos(bx, by, cx, cy, ux, uy, rid, cid, p, bid)) break;
            first = first + 1 == k ? 0 : first + 1; --kept;
        }
        while (kept >= 2) {                                          // clockwise side
            int a = first + kept - 1; if (a >= k) a -= k;
            const int z = a == 0 ? k - 1 : a - 1;
            int qx, qy, qid, zx, zy, zid;
            get(a, qx, qy, qid);
            get(z, zx, zy, zid);
            if (star::cross2(qx, qy, ux, uy) <= 0 || star::cross2(zx, zy, qx, qy) <= 0) break;
            if (!star::inside_sos(zx, zy, qx, qy, ux, uy, rid, zid, p, qid)) break;
            --kept;
        }
        if (kept + 1 > STAR_LANES) return -1;
        int src = first + lane - 1; if (src >= k) src -= k; if (src < 0) src = 0;         // new order: r, then the survivors
        const int nx = __shfl_sync(FULL, ex, src), ny = __shfl_sync(FULL, ey, src), nid = __shfl_sync(FULL, eid, src);
        if (lane == 0) { ex = ux; ey = uy; eid = rid; }
        else if (lane <= kept) { ex = nx; ey = ny; eid = nid; }
        k = kept + 1;
        return 1;
    }
};

__global__ void __launch_bounds__(STAR_THREADS)
delaunay_star_kernel(const int32_t* __restrict__ knot_xy, const float* __restrict__ knot_val,
                     const int32_t* __restrict__ knot_count, int max_knots, uint8_t* ws_base, WsLayout L,
                     int32_t* __restrict__ status, int32_t* __restrict__ tri_out, int32_t* __restrict__ tri_count,
                     float* __restrict__ fill_out, float fill, int npx) {
    extern __shared__ __align__(16) int s_pts[];                 // x | y << 16 of every knot of the frame
    __shared__ int s_unsorted;
    const int frame = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int* counters = (int*)(ws_base + L.counters) + (size_t)frame * CNT_STRIDE;
    TriRec* tris = (TriRec*)(ws_base + L.tris) + (size_t)frame * 2 * max_knots;
    const int n_true = knot_count[frame];
    const int n = (n_true >= 3 && n_true <= max_knots) ? n_true : 0;      // < 3: QhullError upstream; > capacity: reported, never truncated
    const int32_t* xy = knot_xy + (size_t)frame * max_knots * 2;

    // ---- the fill value over this CTA's slice of the map (the rasteriser overwrites the covered pixels) --------------
    if (fill_out) {
        float* o = fill_out + (size_t)frame * npx;
        const int per = ((npx + (int)gridDim.x - 1) / (int)gridDim.x + 3) & ~3;
        const int lo = blockIdx.x * per, hi = min(lo + per, npx);
        if (((((uintptr_t)o) & 15) == 0) && (npx & 3) == 0) {
            const float4 f4 = make_float4(fill, fill, fill, fill);
            for (int i = lo + tid * 4; i < hi; i += STAR_THREADS * 4) __stcs(reinterpret_cast<float4*>(o + i), f4);
        } else {
            for (int i = lo + tid; i < hi; i += STAR_THREADS) o[i] = fill;
        }
    }
    if (tid == 0) s_unsorted = 0;
    __syncthreads();
    int unsorted = 0;
    for (int i = tid; i < n; i += STAR_THREADS) {
        const int x = xy[2 * i], y = xy[2 * i + 1];
        s_pts[i] = (x & 0xffff) | (y << 16);
        if (i > 0 && xy[2 * i - 1] > y) unsorted = 1;            // rows not ascending: the early exit below is off
    }
    if (unsorted) s_unsorted = 1;
    __syncthreads();
    const bool sorted = s_unsorted == 0;

    // Stars are handed out dynamically (one atomic per star): a warp that drew a cheap interior knot takes the next one while
    // a hull knot - whose exterior wedge keeps its scan open to the end of the list - is still running elsewhere.  (With one
    // fixed star per warp the CTA waited for its slowest warp: ncu showed 4.2 barrier-stall cycles per issued instruction.)
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(&counters[CNT_NEXT], 1);
        p = __shfl_sync(FULL, p, 0);
        if (p >= n) break;
        const int px = s_pts[p] & 0xffff, py = s_pts[p] >> 16;
        WarpStar st;
        st.ex = st.ey = st.eid = 0; st.k = 0;
        // Reach of the star = bounding box (relative to p) of the union of its fan circumcircles, in float with a margin:
        // a knot outside every circumcircle and outside the exterior wedge cannot enter the star.  Knots are visited
        // outwards in knot order, rows ascending, so a side (up / down) is finished once its next row is beyond the
        // reach in y; batches with no knot inside the box are skipped.  An exterior wedge (hull knot) makes the sides it
        // opens onto unbounded.
        float bx0 = -3e38f, bx1 = 3e38f, by0 = -3e38f, by1 = 3e38f;
        bool have_reach = false, open = false;
        int oax = 0, oay = 0, obx = 0, oby = 0;                    // the exterior wedge [a, b) of a hull knot
        bool up_done = false, dn_done = false;
        const int dmax = max(p, n - 1 - p);
        for (int base = 0; base < dmax; base += 16) {
            // candidates: lane l looks at distance d = base + l/2 + 1 on side l & 1 (1 = up, towards lower indices)
            const int d = base + (lane >> 1) + 1;
            const bool up = lane & 1;
            const int c = up ? p - d : p + d;
            bool ok = c >= 0 && c < n && !(up ? up_done : dn_done);
            int ux = 0, uy = 0;
            if (ok) { ux = (s_pts[c] & 0xffff) - px; uy = (s_pts[c] >> 16) - py; }
            if (have_reach) {
                const bool inside_box = (float)ux >= bx0 && (float)ux <= bx1 && (float)uy >= by0 && (float)uy <= by1;
                ok = ok && (inside_box || (open && star::in_wedge(oax, oay, obx, oby, ux, uy)));
            }
            bool changed = false;
            if (__any_sync(FULL, ok)) {
                // Candidates of a batch are near p in knot order, not in space: inserting them in arrival order makes most of
                // them transient neighbours (measured: 37 insertions per knot, 32 of them in the first batch).  Instead the
                // nearest accepted candidate goes first and the others are re-tested against the star it changed.
                const bool want = st.accepts_each(p, ux, uy, c);          // every lane must call it: it shuffles (no `ok && ...` short-circuit)
                unsigned m = __ballot_sync(FULL, ok && want);
                const int d2 = ux * ux + uy * uy;
                while (m) {
                    int best = (m >> lane) & 1 ? d2 : 0x7fffffff, who = lane;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const int ob = __shfl_xor_sync(FULL, best, o), ow = __shfl_xor_sync(FULL, who, o);
                        if (ob < best || (ob == best && ow < who)) { best = ob; who = ow; }
                    }
                    m &= ~(1u << who);
                    const int rc = st.insert(p, __shfl_sync(FULL, ux, who), __shfl_sync(FULL, uy, who), __shfl_sync(FULL, c, who), lane);
                    if (rc < 0) { if (lane == 0) atomicExch(&counters[CNT_OVERFLOW], 1); }
                    else if (rc > 0) {
                        changed = true;
                        if (m) m &= __ballot_sync(FULL, st.accepts_each(p, ux, uy, c));
                    }
                }
            }
            if (st.k >= 2 && (changed || !have_reach)) {
                const int ln = lane + 1 == st.k ? 0 : lane + 1;
                const int nx = __shfl_sync(FULL, st.ex, ln), ny = __shfl_sync(FULL, st.ey, ln);
                float x0 = 3e38f, x1 = -3e38f, y0 = 3e38f, y1 = -3e38f;
                bool mine_open = false;
                if (lane < st.k) {
                    const int cr = star::cross2(st.ex, st.ey, nx, ny);
                    if (cr <= 0) mine_open = true;
                    else {
                        const float a2 = (float)(st.ex * st.ex + st.ey * st.ey), b2 = (float)(nx * nx + ny * ny);
                        const float inv = 0.5f / (float)cr;
                        const float ccx = ((float)ny * a2 - (float)st.ey * b2) * inv, ccy = ((float)st.ex * b2 - (float)nx * a2) * inv;
                        const float r = sqrtf(ccx * ccx + ccy * ccy) * 1.002f + 2.0f;    // float rounding stays far inside this margin
                        x0 = ccx - r; x1 = ccx + r; y0 = ccy - r; y1 = ccy + r;
                    }
                }
                const unsigned om = __ballot_sync(FULL, mine_open);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    x0 = fminf(x0, __shfl_xor_sync(FULL, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, o));
                    y0 = fminf(y0, __shfl_xor_sync(FULL, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, o));
                }
                bx0 = x0; bx1 = x1; by0 = y0; by1 = y1;
                open = om != 0;
                have_reach = __popc(om) <= 1;                      // two straight wedges (collinear star): no filter yet
                if (open) {
                    const int j = __ffs(om) - 1;
                    oax = __shfl_sync(FULL, st.ex, j); oay = __shfl_sync(FULL, st.ey, j);
                    obx = __shfl_sync(FULL, nx, j); oby = __shfl_sync(FULL, ny, j);
                }
            }
            // ---- a side is finished when its next row lies beyond the reach (and the exterior wedge does not open onto it) ----
            if (sorted && have_reach) {
                const int nd = base + 17;                            // next distance to look at
                bool open_dn = false, open_up = false;
                if (open) {
                    open_dn = oay > 0 || oby > 0 || star::in_wedge(oax, oay, obx, oby, 0, 1);
                    open_up = oay < 0 || oby < 0 || star::in_wedge(oax, oay, obx, oby, 0, -1);
                }
                if (!dn_done) dn_done = p + nd >= n || (!open_dn && (float)((s_pts[p + nd] >> 16) - py) > by1);
                if (!up_done) up_done = p - nd < 0 || (!open_up && (float)((s_pts[p - nd] >> 16) - py) < by0);
                if (up_done && dn_done) break;
            }
        }
        // ---- every fan triangle (p, a, b) whose lowest knot index is p --------------------------------------------------
        const int k = st.k;
        const int ln = lane + 1 == k ? 0 : lane + 1;
        const int nx = __shfl_sync(FULL, st.ex, ln), ny = __shfl_sync(FULL, st.ey, ln), nid = __shfl_sync(FULL, st.eid, ln);
        if (lane < k && k >= 2 && star::cross2(st.ex, st.ey, nx, ny) > 0 && p < st.eid && p < nid) {
            const int slot = atomicAdd(&counters[CNT_TRIS], 1);
            if (slot < 2 * max_knots) {
                const float* val = knot_val + (size_t)frame * max_knots;
                TriRec t;
                t.ax = px; t.ay = py; t.bx = px + st.ex; t.by = py + st.ey; t.cx = px + nx; t.cy = py + ny;
                t.va = val[p]; t.vb = val[st.eid]; t.vc = val[nid];
                t.ia = p; t.ib = st.eid; t.ic = nid;
                tris[slot] = t;
                if (tri_out) {
                    int32_t* o = tri_out + ((size_t)frame * 2 * max_knots + slot) * 3;
                    o[0] = p; o[1] = st.eid; o[2] = nid;
                }
            } else {
                atomicExch(&counters[CNT_OVERFLOW], 1);
            }
        }
    }
    // ---- the last CTA of the frame publishes status and triangle count ---------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&counters[CNT_DONE], 1) == (int)gridDim.x - 1) {
            __threadfence();
            const int nt = min(atomicAdd(&counters[CNT_TRIS], 0), 2 * max_knots);
            const int ovf = atomicAdd(&counters[CNT_OVERFLOW], 0);
            int st = RIDERS_OK;
            if (n_true > max_knots || ovf) st = RIDERS_EINTERNAL;
            else if (n_true < 3 || nt == 0) st = RIDERS_EQHULL;
            status[frame] = st;
            if (st != RIDERS_OK) counters[CNT_TRIS] = 0;          // the frame keeps the fill value
            if (tri_count) tri_count[frame] = st == RIDERS_OK ? nt : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fill value (only when riders_interp_fill is called on its own) and the scatter rasteriser
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fill_value_kernel(float* __restrict__ out, size_t total, float fill) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) out[i] = fill;
}

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_TRIS = 8;              // triangles per CTA
constexpr int RASTER_WIDE = 64;             // bounding boxes wider than this (hull slivers) take the row path

// Pixels exactly ON an edge (or at a knot) belong to both triangles that share it: their value is computed from the edge alone, in
// a canonical form - endpoints in lexicographic order, weights w0 = |p1 - p| and w1 = |p - p0| along the axis the edge advances on,
// v = (w0 v0 + w1 v1) * (1 / (w0 + w1)) in double - so both triangles write the same bits whichever comes last.
struct EdgeCanon { int x0, y0, x1, y1; float v0, v1; double inv_len; };
__device__ __forceinline__ EdgeCanon canon_edge(int x0, int y0, float v0, int x1, int y1, float v1) {
    EdgeCanon e;
    const bool swap = x0 > x1 || (x0 == x1 && y0 > y1);
    e.x0 = swap ? x1 : x0; e.y0 = swap ? y1 : y0; e.v0 = swap ? v1 : v0;
    e.x1 = swap ? x0 : x1; e.y1 = swap ? y0 : y1; e.v1 = swap ? v0 : v1;
    e.inv_len = 1.0 / (double)(e.x1 != e.x0 ? e.x1 - e.x0 : e.y1 - e.y0);
    return e;
}
__device__ __forceinline__ float edge_value(int x0, int y0, float v0, int x1, int y1, float v1, double inv_len, int px, int py) {
    int w0, w1;
    if (x1 != x0) { w0 = x1 - px; w1 = px - x0; } else { w0 = y1 - py; w1 = py - y0; }
    return (float)(fma((double)w0, (double)v0, (double)w1 * (double)v1) * inv_len);
}

// per-triangle constants of the rasteriser, set up once by one thread and then shared by the CTA (16-byte groups: LDS.128).
// Edge functions as planes: wa = orient(p, b, c) = Aa x + Ba y + Ca, wb = orient(p, c, a), wc = orient(p, a, b); all exact in
// int32 (coordinates < 2^13).  The interpolant is the affine function v = va + gx (x - ax) + gy (y - ay) (= the barycentric sum).
struct __align__(16) TriSetup {
    int Aa, Ba, Ca, Ab;
    int Bb, Cb, Ac, Bc;
    int Cc, x0, y0, bw;             // clipped bounding box: origin and width
    float inv_bw, va; int ax, ay;
    double gx, gy;
    int x1, rows, pad0, pad1;
    EdgeCanon ebc, eca, eab;        // the three edges in canonical form (pixels with wa == 0, wb == 0, wc == 0)
};

__device__ __forceinline__ void shade_pixel(const TriSetup& t, int x, int y, int wa, int wb, int wc, float* __restrict__ o, int W) {
    float v;
    if (wa == 0 || wb == 0 || wc == 0) {                 // on an edge (or at a knot): one evaluation, whichever edge it is
        const EdgeCanon& e = wa == 0 ? t.ebc : (wb == 0 ? t.eca : t.eab);
        v = edge_value(e.x0, e.y0, e.v0, e.x1, e.y1, e.v1, e.inv_len, x, y);
    } else {
        v = (float)fma(t.gx, (double)(x - t.ax), fma(t.gy, (double)(y - t.ay), (double)t.va));
    }
    o[(size_t)y * W + x] = v;
}

// A pixel is written iff its three exact integer edge functions are >= 0.  The CTA's RASTER_TRIS triangles are flattened:
//  * bounding boxes up to RASTER_WIDE pixels wide into one list of bounding-box PIXELS dealt to the 256 threads - every lane
//    tests a pixel of some box (round 2's first version dealt box ROWS to warps: a 25-pixel-wide triangle used 12 of the 32
//    lanes of every row and paid the row set-up 25 times);
//  * wider ones (hull slivers, frames with a handful of knots) into a list of box rows dealt to the warps, each row clipped
//    to its conservative span first, so a 600-pixel sliver costs its area and not its box.
__global__ void __launch_bounds__(RASTER_THREADS, 4)
raster_kernel(uint8_t* ws_base, WsLayout L, int max_knots, int H, int W, float* __restrict__ out) {
    __shared__ TriSetup s_tri[RASTER_TRIS];
    __shared__ int s_slot0[RASTER_TRIS + 1];
    __shared__ int s_row0[RASTER_TRIS + 1];
    const int frame = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int* counters = (const int*)(ws_base + L.counters) + (size_t)frame * CNT_STRIDE;
    const TriRec* tris = (const TriRec*)(ws_base + L.tris) + (size_t)frame * 2 * max_knots;
    const int nt = min(counters[CNT_TRIS], 2 * max_knots);
    if ((int)blockIdx.x * RASTER_TRIS >= nt) return;
    float* o = out + (size_t)frame * H * W;
    if (warp == 0) {
        const int ti = blockIdx.x * RASTER_TRIS + lane;
        int slots = 0, wrows = 0;
        if (lane < RASTER_TRIS) {
            TriSetup u;
            u.rows = 0; u.bw = 1; u.inv_bw = 1.f; u.pad0 = u.pad1 = 0;
            if (ti < nt) {
                const TriRec t = tris[ti];
                u.ax = t.ax; u.ay = t.ay; u.va = t.va;
                u.ebc = canon_edge(t.bx, t.by, t.vb, t.cx, t.cy, t.vc);
                u.eca = canon_edge(t.cx, t.cy, t.vc, t.ax, t.ay, t.va);
                u.eab = canon_edge(t.ax, t.ay, t.va, t.bx, t.by, t.vb);
                u.Aa = t.by - t.cy; u.Ba = t.cx - t.bx; u.Ca = t.bx * t.cy - t.cx * t.by;
                u.Ab = t.cy - t.ay; u.Bb = t.ax - t.cx; u.Cb = t.cx * t.ay - t.ax * t.cy;
                u.Ac = t.ay - t.by; u.Bc = t.bx - t.ax; u.Cc = t.ax * t.by - t.bx * t.ay;
                u.x0 = max(min(t.ax, min(t.bx, t.cx)), 0); u.x1 = min(max(t.ax, max(t.bx, t.cx)), W - 1);
                u.y0 = max(min(t.ay, min(t.by, t.cy)), 0);
                u.rows = max(min(max(t.ay, max(t.by, t.cy)), H - 1) - u.y0 + 1, 0);
                u.bw = max(u.x1 - u.x0 + 1, 1);
                u.inv_bw = 1.0f / (float)u.bw;
                const double det = (double)((t.bx - t.ax) * (t.cy - t.ay) - (t.by - t.ay) * (t.cx - t.ax));
                const double dvb = (double)t.vb - (double)t.va, dvc = (double)t.vc - (double)t.va;
                u.gx = (dvb * (double)(t.cy - t.ay) - dvc * (double)(t.by - t.ay)) / det;
                u.gy = ((double)(t.bx - t.ax) * dvc - (double)(t.cx - t.ax) * dvb) / det;
                if (u.bw <= RASTER_WIDE) slots = u.bw * u.rows; else wrows = u.rows;
            }
            s_tri[lane] = u;
        }
        int ps = slots, pr = wrows;                      // inclusive scans over the first RASTER_TRIS lanes
#pragma unroll
        for (int d = 1; d < RASTER_TRIS; d <<= 1) {
            const int a = __shfl_up_sync(FULL, ps, d), b2 = __shfl_up_sync(FULL, pr, d);
            if (lane >= d) { ps += a; pr += b2; }
        }
        if (lane < RASTER_TRIS) { s_slot0[lane + 1] = ps; s_row0[lane + 1] = pr; }
        if (lane == 0) { s_slot0[0] = 0; s_row0[0] = 0; }
    }
    __syncthreads();
    // ---- narrow boxes: one pixel of some box per thread and step ----------------------------------------------------------
    const int total = s_slot0[RASTER_TRIS];
    for (int s = threadIdx.x; s < total; s += RASTER_THREADS) {
        int cur = 0;
#pragma unroll
        for (int step = RASTER_TRIS / 2; step > 0; step >>= 1)
            if (s >= s_slot0[cur + step]) cur += step;
        const TriSetup& t = s_tri[cur];
        const int4 q0 = *reinterpret_cast<const int4*>(&t.Aa), q1 = *reinterpret_cast<const int4*>(&t.Bb), q2 = *reinterpret_cast<const int4*>(&t.Cc);
        const int local = s - s_slot0[cur], bw = q2.w;
        int r = (int)((float)local * t.inv_bw);          // local < 2^24: exact in float; the quotient is off by at most one
        int c = local - r * bw;
        if (c < 0) { --r; c += bw; } else if (c >= bw) { ++r; c -= bw; }
        const int x = q2.y + c, y = q2.z + r;
        const int wa = q0.x * x + q0.y * y + q0.z, wb = q0.w * x + q1.x * y + q1.y, wc = q1.z * x + q1.w * y + q2.x;
        if ((wa | wb | wc) >= 0) shade_pixel(t, x, y, wa, wb, wc, o, W);
    }
    // ---- wide boxes: their rows, dealt round-robin to the warps ------------------------------------------------------------
    const int total_rows = s_row0[RASTER_TRIS];
    int cur = 0;
    for (int r = warp; r < total_rows; r += RASTER_THREADS / 32) {
        while (r >= s_row0[cur + 1]) ++cur;
        const TriSetup& t = s_tri[cur];
        const int y = t.y0 + (r - s_row0[cur]);
        const int Aa = t.Aa, Ab = t.Ab, Ac = t.Ac;
        const int Ca = t.Ba * y + t.Ca, Cb = t.Bb * y + t.Cb, Cc = t.Bc * y + t.Cc;
        // the row's conservative x span from the three half-planes, in float; the exact integer test below then only visits
        // the chunks near the triangle
        float lo = (float)t.x0, hi = (float)t.x1;
        if (Aa > 0) lo = fmaxf(lo, -(float)Ca / (float)Aa - 0.02f); else if (Aa < 0) hi = fminf(hi, (float)Ca / (float)(-Aa) + 0.02f); else if (Ca < 0) continue;
        if (Ab > 0) lo = fmaxf(lo, -(float)Cb / (float)Ab - 0.02f); else if (Ab < 0) hi = fminf(hi, (float)Cb / (float)(-Ab) + 0.02f); else if (Cb < 0) continue;
        if (Ac > 0) lo = fmaxf(lo, -(float)Cc / (float)Ac - 0.02f); else if (Ac < 0) hi = fminf(hi, (float)Cc / (float)(-Ac) + 0.02f); else if (Cc < 0) continue;
        const int xs = max(t.x0, (int)floorf(lo)), xe = min(t.x1, (int)ceilf(hi));
        for (int xb = xs; xb <= xe; xb += 32) {
            const int x = xb + lane;
            const int wa = Aa * x + Ca, wb = Ab * x + Cb, wc = Ac * x + Cc;
            if (x <= xe && (wa | wb | wc) >= 0) shade_pixel(t, x, y, wa, wb, wc, o, W);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// griddata(method='nearest'): value of the nearest knot (exact integer distances; among equidistant knots the lowest index -
// SciPy's cKDTree leaves that choice undefined).  Brute force over the frame's knots staged in shared memory: 640 x 512 pixels
// x 512 knots = 168 M distance evaluations per frame.
// ---------------------------------------------------------------------------------------------
constexpr int NEAR_CHUNK = 1024;
__global__ void __launch_bounds__(256)
interp_nearest_kernel(const int32_t* __restrict__ knot_xy, const float* __restrict__ knot_val, const int32_t* __restrict__ knot_count,
                      int max_knots, int H, int W, float fill, float* __restrict__ out) {
    __shared__ int sx[NEAR_CHUNK], sy[NEAR_CHUNK];
    const int frame = blockIdx.y;
    const int n = min(knot_count[frame], max_knots);
    const int px_i = blockIdx.x * 256 + threadIdx.x;
    const int x = px_i % W, y = px_i / W;
    int best = 0x7fffffff, arg = -1;
    for (int base = 0; base < n; base += NEAR_CHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < min(NEAR_CHUNK, n - base); i += 256) {
            sx[i] = knot_xy[((size_t)frame * max_knots + base + i) * 2];
            sy[i] = knot_xy[((size_t)frame * max_knots + base + i) * 2 + 1];
        }
        __syncthreads();
        const int m = min(NEAR_CHUNK, n - base);
        for (int i = 0; i < m; ++i) {
            const int dx = sx[i] - x, dy = sy[i] - y, d = dx * dx + dy * dy;
            if (d < best) { best = d; arg = base + i; }
        }
    }
    if (px_i < H * W) out[(size_t)frame * H * W + px_i] = arg >= 0 ? knot_val[(size_t)frame * max_knots + arg] : fill;
}

}  // namespace riders

extern "C" int riders_interp_nearest(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                                     int B, int max_knots, int H, int W, float fill, float* out, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0 && H <= 8192 && W <= 8192 && max_knots > 0 && B <= 65535, "interp_nearest: bad shape");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(knot_xy && knot_val && knot_count && out, "interp_nearest: null pointer");
    count_launches(1);
    interp_nearest_kernel<<<dim3(ceil_div(H * W, 256), B), 256, 0, (cudaStream_t)stream>>>(knot_xy, knot_val, knot_count, max_knots, H, W, fill, out);
    return check_launch("interp_nearest");
}

extern "C" size_t riders_interp_workspace_bytes(int B, int max_knots, int H, int W) {
    (void)H; (void)W;
    if (B <= 0 || max_knots <= 0) return 256;
    return riders::ws_layout(B, max_knots).total + 256;
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
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* agg = oc_scratch(B, grid.x, st);
    extract_knots_kernel<<<grid, OC_THREADS, 0, st>>>(pred, sparse, valid, H, W, max_knots,
                                                      knot_xy, knot_scale, knot_shift, knot_count, agg);
    if (agg) cudaFreeAsync(agg, st);
    return check_launch("extract_knots");
}

static int interp_triangulate_impl(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                                   int B, int max_knots, int32_t* status, int32_t* tri_out, int32_t* tri_count,
                                   float* fill_out, float fill, int npx, void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0, "interp_triangulate: bad shape");
    RIDERS_REQUIRE(max_knots > 0 && max_knots <= 32768, "interp_triangulate: max_knots=%d out of range (1..32768)", max_knots);
    RIDERS_REQUIRE(B <= 65535, "interp_triangulate: at most 65535 frames per call");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(knot_xy && knot_val && knot_count && status && workspace, "interp_triangulate: null pointer");
    RIDERS_REQUIRE((((uintptr_t)workspace) & 255) == 0, "interp_triangulate: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const WsLayout L = ws_layout(B, max_knots);
    if (cudaMemsetAsync((uint8_t*)workspace + L.counters, 0, (size_t)B * CNT_STRIDE * sizeof(int), st) != cudaSuccess)
        return check_launch("interp_triangulate: counter reset");
    const size_t smem = (size_t)max_knots * sizeof(int);
    static size_t configured[64] = {0};
    if (smem > 48 * 1024)
        if (int rc_smem = ensure_dynamic_smem(delaunay_star_kernel, smem, configured, "interp_triangulate: smem attribute")) return rc_smem;
    count_launches(1);
    // CTAs per frame: the resident capacity of the device shared among the frames (one wave; the warps then draw stars until
    // the frame's list is exhausted), at most one warp per knot
    static int capacity[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (capacity[dev] == 0) {
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, delaunay_star_kernel, STAR_THREADS, smem);
        capacity[dev] = max(sms * max(per_sm, 1), 1);
        cudaGetLastError();
    }
    const int per_frame = min(max(capacity[dev] / B, 1), ceil_div(max_knots, STAR_WARPS));
    dim3 grid(per_frame, B);
    delaunay_star_kernel<<<grid, STAR_THREADS, smem, st>>>(knot_xy, knot_val, knot_count, max_knots, (uint8_t*)workspace, L,
                                                           status, tri_out, tri_count, fill_out, fill, npx);
    return check_launch("interp_triangulate");
}

extern "C" int riders_interp_triangulate(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                                         int B, int max_knots, int32_t* status,
                                         int32_t* tri_out, int32_t* tri_count, void* workspace, void* stream) {
    return interp_triangulate_impl(knot_xy, knot_val, knot_count, B, max_knots, status, tri_out, tri_count, nullptr, 0.f, 0, workspace, stream);
}

static int interp_raster_impl(int B, int max_knots, int H, int W, float fill, bool prefill, float* out, void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0, "interp_fill: bad shape");
    RIDERS_REQUIRE(H <= 8192 && W <= 8192, "interp_fill: exact predicates need H, W <= 8192");
    RIDERS_REQUIRE(max_knots > 0 && max_knots <= 32768 && B <= 65535, "interp_fill: max_knots / B out of range");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(out && workspace, "interp_fill: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const WsLayout L = ws_layout(B, max_knots);
    if (prefill) {
        count_launches(1);
        const size_t total = (size_t)B * H * W;
        fill_value_kernel<<<(unsigned)min((size_t)2048, (total + 1023) / 1024), 256, 0, st>>>(out, total, fill);
    }
    count_launches(1);
    dim3 grid(ceil_div(2 * max_knots, RASTER_TRIS), B);
    raster_kernel<<<grid, RASTER_THREADS, 0, st>>>((uint8_t*)workspace, L, max_knots, H, W, out);
    return check_launch("interp_fill");
}

extern "C" int riders_interp_fill(int B, int max_knots, int H, int W, float fill, float* out, void* workspace, void* stream) {
    return interp_raster_impl(B, max_knots, H, W, fill, true, out, workspace, stream);
}

extern "C" int riders_interp_linear(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                                    int B, int max_knots, int H, int W, float fill,
                                    float* out, int32_t* status, int32_t* tri_out, int32_t* tri_count,
                                    void* workspace, void* stream) {
    using namespace riders;
    RIDERS_REQUIRE(B >= 0 && H > 0 && W > 0 && H <= 8192 && W <= 8192, "interp_linear: bad shape");
    if (B == 0) return RIDERS_OK;
    RIDERS_REQUIRE(out, "interp_linear: null pointer");
    // two launches: stars (+ the fill value over the maps), then the rasteriser
    int rc = interp_triangulate_impl(knot_xy, knot_val, knot_count, B, max_knots, status, tri_out, tri_count, out, fill, H * W, workspace, stream);
    if (rc != RIDERS_OK) return rc;
    return interp_raster_impl(B, max_knots, H, W, fill, false, out, workspace, stream);
}
