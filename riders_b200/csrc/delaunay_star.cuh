// Exact Delaunay triangulation of integer pixel knots as INDEPENDENT per-site stars (A8: modules/interpolator.py:7-17
// delegates to scipy.interpolate.griddata -> Qhull Delaunay; here the triangulation runs on the device).
//
// Why stars: the round-1 kernel was a Guibas-Stolfi divide and conquer, one CTA per frame, whose top-level merges are a
// serial pointer chase by ONE thread (238 us for 200 knots, a cliff at >680 knots when the mesh left shared memory).  The
// star of a site p - its Delaunay neighbours in counter-clockwise order - depends only on the point set, so the n stars
// of a frame are n independent tasks (one warp each): the whole GPU works on one frame, and B frames are B*n warps.
//
// Star of p, incrementally (Bowyer-Watson restricted to the triangles around p):  the processed sites T define the fan
// (p, q_j, q_j+1).  A new site r is a neighbour of p in DT(T + r) iff it lies inside the circumcircle of the fan triangle
// whose angular wedge contains the direction p->r (the Voronoi vertex extreme in that direction), or the wedge is the
// exterior one (angle >= 180 deg: p is on the hull and sees r).  A rejection is final: the Voronoi cell of p only shrinks.
// After r is accepted the neighbours next to it are removed while r is inside the circumcircle of the next fan triangle
// on that side (the cavity of r, walked from p).  Processing order is irrelevant, so candidates are tested 32 at a time
// against a read-only star and only the few accepted ones are inserted serially.
//
// Predicates are exact (coordinates < 2^13: orientation < 2^28, in-circle < 2^57 in int64).  Co-circular quadruples
// (in-circle == 0) are decided by a symbolic perturbation ordered by knot index (the diagonal through the lowest-index
// knot of the four wins), which is globally consistent: every star belongs to the same valid triangulation.  Which
// diagonal Qhull picks there is outside the reference's definition (the parity tests mask those quads).
//
// Everything is __host__ __device__: tests/ compiles this header with g++ and compares with scipy.spatial.Delaunay.
#pragma once
#include "delaunay_core.cuh"

namespace riders {
namespace star {

using dt::i64;

constexpr int STAR_MAX = 64;            // neighbours of one site (typical 6, random sets stay below 20); overflow is reported

struct Star {
    int x[STAR_MAX], y[STAR_MAX];       // neighbour coordinates relative to p, counter-clockwise, cyclic
    int id[STAR_MAX];                   // knot index of the neighbour
    int k;
};

__host__ __device__ inline int cross2(int ax, int ay, int bx, int by) { return ax * by - ay * bx; }     // |coord| < 2^13: fits int32
__host__ __device__ inline int dot2(int ax, int ay, int bx, int by) { return ax * bx + ay * by; }

// direction u in the half-open wedge [dir a, dir b) swept counter-clockwise from a to b (a, b distinct directions)
__host__ __device__ inline bool in_wedge(int ax, int ay, int bx, int by, int ux, int uy) {
    const int ca = cross2(ax, ay, ux, uy), cb = cross2(ux, uy, bx, by), cab = cross2(ax, ay, bx, by);
    const bool at_a = (ca == 0) && dot2(ax, ay, ux, uy) > 0;
    if (cab > 0) return (ca > 0 && cb > 0) || at_a;             // wedge < 180 deg
    if (cab < 0) return ca >= 0 || cb > 0;                       // wedge > 180 deg (ca == 0: along a, or opposite a - both inside)
    return ca > 0 || at_a;                                       // a, b opposite: wedge = exactly 180 deg
}

// r strictly inside the circle through p (the origin), a, b (counter-clockwise), ties by the symbolic perturbation:
// `partner` is the knot opposite r in the quadrilateral the four form, `other0/1` the remaining two.
__host__ __device__ inline bool inside_sos(int ax, int ay, int bx, int by, int rx, int ry,
                                           int id_r, int id_partner, int id_o0, int id_o1) {
    const dt::Pt o = {0, 0}, a = {ax, ay}, b = {bx, by}, r = {rx, ry};
    const i64 det = dt::incircle(o, a, b, r);
    if (det != 0) return det > 0;
    const int m0 = id_r < id_partner ? id_r : id_partner, m1 = id_o0 < id_o1 ? id_o0 : id_o1;
    return m0 < m1;
}

// index of the wedge (star[j], star[j+1]) that contains direction u; k >= 2
__host__ __device__ inline int find_wedge(const Star& s, int ux, int uy) {
    for (int j = 0; j < s.k; ++j) {
        const int n = j + 1 == s.k ? 0 : j + 1;
        if (in_wedge(s.x[j], s.y[j], s.x[n], s.y[n], ux, uy)) return j;
    }
    return 0;   // unreachable for distinct directions
}

// Would site r (relative coordinates u, knot index id_r) enter the star of p (knot index id_p)?  Read-only.
__host__ __device__ inline bool accepts(const Star& s, int id_p, int ux, int uy, int id_r) {
    if (s.k == 0) return true;
    int j = 0;
    if (s.k >= 2) j = find_wedge(s, ux, uy);
    const int ax = s.x[j], ay = s.y[j];
    if (cross2(ax, ay, ux, uy) == 0 && dot2(ax, ay, ux, uy) > 0)                 // same ray as a: only the nearer one is a neighbour
        return dot2(ux, uy, ux, uy) < dot2(ax, ay, ax, ay);
    if (s.k == 1) return true;
    const int n = j + 1 == s.k ? 0 : j + 1;
    const int bx = s.x[n], by = s.y[n];
    if (cross2(ax, ay, bx, by) <= 0) return true;                                 // exterior wedge of a hull site
    return inside_sos(ax, ay, bx, by, ux, uy, id_r, id_p, s.id[j], s.id[n]);
}

// Insert r if it belongs to the star (re-evaluated against the current star `s`): 0 = rejected (s unchanged), 1 = the new
// star is in `out` (s unchanged; the caller swaps the two buffers), -1 = more than STAR_MAX neighbours.
__host__ __device__ inline int insert(const Star& s, Star& out, int id_p, int ux, int uy, int id_r) {
    if (!accepts(s, id_p, ux, uy, id_r)) return 0;
    const int k = s.k;
    if (k == 0) { out.x[0] = ux; out.y[0] = uy; out.id[0] = id_r; out.k = 1; return 1; }
    const int j = k >= 2 ? find_wedge(s, ux, uy) : 0;
    // survivors: `kept` entries starting (cyclically) at `first`.  If r sits on the ray of neighbour a = s[j] it is the
    // nearer one (accepts() said so) and a leaves; otherwise all k entries survive so far, a being the last of them.
    const bool same_ray = cross2(s.x[j], s.y[j], ux, uy) == 0 && dot2(s.x[j], s.y[j], ux, uy) > 0;
    int first = j + 1 == k ? 0 : j + 1;
    int kept = same_ray ? k - 1 : k;
    // Walk counter-clockwise from b = s[first]: a neighbour leaves while both fan triangles next to its edge exist and r is
    // inside the circumcircle of the far one (the cavity of r seen from p).
    while (kept >= 2) {
        const int b = first, c = b + 1 == k ? 0 : b + 1;
        if (cross2(ux, uy, s.x[b], s.y[b]) <= 0) break;                           // no triangle (p, r, b)
        if (cross2(s.x[b], s.y[b], s.x[c], s.y[c]) <= 0) break;                   // no triangle (p, b, c)
        if (!inside_sos(s.x[b], s.y[b], s.x[c], s.y[c], ux, uy, id_r, s.id[c], id_p, s.id[b])) break;
        first = c; --kept;
    }
    // ... and clockwise from the last survivor a, with z the one before it
    while (kept >= 2) {
        int a = first + kept - 1; if (a >= k) a -= k;
        const int z = a == 0 ? k - 1 : a - 1;
        if (cross2(s.x[a], s.y[a], ux, uy) <= 0) break;                           // no triangle (p, a, r)
        if (cross2(s.x[z], s.y[z], s.x[a], s.y[a]) <= 0) break;                   // no triangle (p, z, a)
        if (!inside_sos(s.x[z], s.y[z], s.x[a], s.y[a], ux, uy, id_r, s.id[z], id_p, s.id[a])) break;
        --kept;
    }
    if (kept + 1 > STAR_MAX) return -1;
    out.x[0] = ux; out.y[0] = uy; out.id[0] = id_r;                               // new cyclic order: r, then the survivors
    for (int t = 0; t < kept; ++t) {
        int i = first + t; if (i >= k) i -= k;
        out.x[t + 1] = s.x[i]; out.y[t + 1] = s.y[i]; out.id[t + 1] = s.id[i];
    }
    out.k = kept + 1;
    return 1;
}

// Whole star of knot p over knots [0, n) in index order (host reference of what one warp does on the device).
__host__ __device__ inline int build_star(const dt::Pt* pts, int n, int p, Star* buf /* two Stars */) {
    int cur = 0;
    buf[0].k = 0;
    for (int r = 0; r < n; ++r) {
        if (r == p) continue;
        const int rc = insert(buf[cur], buf[cur ^ 1], p, pts[r].x - pts[p].x, pts[r].y - pts[p].y, r);
        if (rc < 0) return -1;
        cur ^= rc;
    }
    return cur;
}

}  // namespace star
}  // namespace riders
