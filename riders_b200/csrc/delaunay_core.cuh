// Exact-arithmetic Delaunay triangulation of integer pixel knots (Guibas-Stolfi divide and
// conquer on a quad-edge mesh).  Core of kernel 5 (A8: modules/interpolator.py:7-17 delegates to
// scipy.interpolate.griddata -> Qhull Delaunay; here the triangulation runs on the device).
//
// Everything is __host__ __device__ so tests/ can compile this header with g++ and compare the
// triangulation with scipy.spatial.Delaunay on the CPU before any GPU time is spent.
//
// Predicates are exact: coordinates are integers < 2^13, so orientation fits in 2^28 and the
// in-circle determinant in 2^57 (int64).  Co-circular quadruples (in-circle == 0) are NOT
// flipped; which diagonal Qhull picks there is outside the reference's definition.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace riders {
namespace dt {

struct Pt { int x, y; };

typedef long long i64;

__host__ __device__ inline i64 orient(const Pt& a, const Pt& b, const Pt& c) {
    return (i64)(b.x - a.x) * (c.y - a.y) - (i64)(b.y - a.y) * (c.x - a.x);
}
// > 0  <=>  d strictly inside the circle through a, b, c (given counter-clockwise)
__host__ __device__ inline i64 incircle(const Pt& a, const Pt& b, const Pt& c, const Pt& d) {
    const i64 ax = a.x - d.x, ay = a.y - d.y, bx = b.x - d.x, by = b.y - d.y, cx = c.x - d.x, cy = c.y - d.y;
    return (ax * ax + ay * ay) * (bx * cy - cx * by) - (bx * bx + by * by) * (ax * cy - cx * ay) +
           (cx * cx + cy * cy) * (ax * by - bx * ay);
}

// ---------------------------------------------------------------------------------------------
// quad-edge mesh.  Edge reference e = 4*quad + rot.  next[e] = Onext(e); org[e] valid for rot 0, 2.
// A free quad has org[4q] == -1 and next[4q] = next free quad (or -1).
// ---------------------------------------------------------------------------------------------
struct Mesh {
    int* next;
    int* org;
    int* fresh;        // shared bump counter of never-used quads
    int capacity;      // quads
    const Pt* pts;     // sorted by (x, y)
    int* overflow;     // set to 1 when the pool is exhausted
};

struct Task {          // per-thread allocation state of one merge / leaf
    int free_head;
};

__host__ __device__ inline int rot(int e) { return (e & ~3) | ((e + 1) & 3); }
__host__ __device__ inline int sym(int e) { return (e & ~3) | ((e + 2) & 3); }
__host__ __device__ inline int irot(int e) { return (e & ~3) | ((e + 3) & 3); }
__host__ __device__ inline int onext(const Mesh& m, int e) { return m.next[e]; }
__host__ __device__ inline int oprev(const Mesh& m, int e) { return rot(m.next[rot(e)]); }
__host__ __device__ inline int lnext(const Mesh& m, int e) { return rot(m.next[irot(e)]); }
__host__ __device__ inline int rprev(const Mesh& m, int e) { return m.next[sym(e)]; }
__host__ __device__ inline int orgv(const Mesh& m, int e) { return m.org[e]; }
__host__ __device__ inline int dstv(const Mesh& m, int e) { return m.org[sym(e)]; }

__host__ __device__ inline int alloc_quad(Mesh& m, Task& t) {
    int q;
    if (t.free_head >= 0) {
        q = t.free_head;
        t.free_head = m.next[4 * q];
    } else {
#ifdef __CUDA_ARCH__
        q = atomicAdd(m.fresh, 1);
#else
        q = (*m.fresh)++;
#endif
        if (q >= m.capacity) { *m.overflow = 1; q = m.capacity - 1; }
    }
    return q;
}

__host__ __device__ inline int make_edge(Mesh& m, Task& t, int o, int d) {
    const int e = 4 * alloc_quad(m, t);
    m.next[e] = e; m.next[e + 1] = e + 3; m.next[e + 2] = e + 2; m.next[e + 3] = e + 1;
    m.org[e] = o; m.org[e + 2] = d;
    return e;
}

__host__ __device__ inline void splice(Mesh& m, int a, int b) {
    const int alpha = rot(m.next[a]), beta = rot(m.next[b]);
    const int t1 = m.next[b], t2 = m.next[a], t3 = m.next[beta], t4 = m.next[alpha];
    m.next[a] = t1; m.next[b] = t2; m.next[alpha] = t3; m.next[beta] = t4;
}

__host__ __device__ inline int connect(Mesh& m, Task& t, int a, int b) {
    const int e = make_edge(m, t, dstv(m, a), orgv(m, b));
    splice(m, e, lnext(m, a));
    splice(m, sym(e), b);
    return e;
}

__host__ __device__ inline void delete_edge(Mesh& m, Task& t, int e) {
    splice(m, e, oprev(m, e));
    splice(m, sym(e), oprev(m, sym(e)));
    const int q = e >> 2;
    m.org[4 * q] = -1; m.org[4 * q + 2] = -1;
    m.next[4 * q] = t.free_head;
    t.free_head = q;
}

__host__ __device__ inline bool left_of(const Mesh& m, int p, int e) {
    return orient(m.pts[p], m.pts[orgv(m, e)], m.pts[dstv(m, e)]) > 0;
}
__host__ __device__ inline bool right_of(const Mesh& m, int p, int e) {
    return orient(m.pts[p], m.pts[dstv(m, e)], m.pts[orgv(m, e)]) > 0;
}

struct Hull { int le, re; };     // ccw hull edge out of the leftmost point, cw hull edge out of the rightmost

// 2 or 3 points [lo, lo+n)
__host__ __device__ inline Hull build_leaf(Mesh& m, Task& t, int lo, int n) {
    Hull h;
    if (n == 2) {
        const int a = make_edge(m, t, lo, lo + 1);
        h.le = a; h.re = sym(a);
        return h;
    }
    const int a = make_edge(m, t, lo, lo + 1);
    const int b = make_edge(m, t, lo + 1, lo + 2);
    splice(m, sym(a), b);
    const i64 o = orient(m.pts[lo], m.pts[lo + 1], m.pts[lo + 2]);
    if (o > 0) {
        connect(m, t, b, a);
        h.le = a; h.re = sym(b);
    } else if (o < 0) {
        const int c = connect(m, t, b, a);
        h.le = sym(c); h.re = c;
    } else {
        h.le = a; h.re = sym(b);
    }
    return h;
}

__host__ __device__ inline bool valid_cand(const Mesh& m, int e, int basel) { return right_of(m, dstv(m, e), basel); }

__host__ __device__ inline Hull merge_hulls(Mesh& m, Task& t, Hull L, Hull R) {
    int ldo = L.le, ldi = L.re, rdi = R.le, rdo = R.re;
    // lower common tangent
    for (;;) {
        if (left_of(m, orgv(m, rdi), ldi)) ldi = lnext(m, ldi);
        else if (right_of(m, orgv(m, ldi), rdi)) rdi = rprev(m, rdi);
        else break;
    }
    int basel = connect(m, t, sym(rdi), ldi);
    if (orgv(m, ldi) == orgv(m, ldo)) ldo = sym(basel);
    if (orgv(m, rdi) == orgv(m, rdo)) rdo = basel;
    for (;;) {
        int lcand = onext(m, sym(basel));
        if (valid_cand(m, lcand, basel)) {
            while (incircle(m.pts[dstv(m, basel)], m.pts[orgv(m, basel)], m.pts[dstv(m, lcand)],
                            m.pts[dstv(m, onext(m, lcand))]) > 0) {
                const int nx = onext(m, lcand);
                delete_edge(m, t, lcand);
                lcand = nx;
            }
        }
        int rcand = oprev(m, basel);
        if (valid_cand(m, rcand, basel)) {
            while (incircle(m.pts[dstv(m, basel)], m.pts[orgv(m, basel)], m.pts[dstv(m, rcand)],
                            m.pts[dstv(m, oprev(m, rcand))]) > 0) {
                const int nx = oprev(m, rcand);
                delete_edge(m, t, rcand);
                rcand = nx;
            }
        }
        const bool lv = valid_cand(m, lcand, basel), rv = valid_cand(m, rcand, basel);
        if (!lv && !rv) break;
        if (!lv || (rv && incircle(m.pts[dstv(m, lcand)], m.pts[orgv(m, lcand)], m.pts[orgv(m, rcand)],
                                   m.pts[dstv(m, rcand)]) > 0))
            basel = connect(m, t, rcand, sym(basel));
        else
            basel = connect(m, t, sym(basel), sym(lcand));
    }
    Hull h;
    h.le = ldo; h.re = rdo;
    return h;
}

// ---------------------------------------------------------------------------------------------
// The recursion tree of the divide and conquer, addressed by (depth, index):
// node (0,0) = [0, n); children of [lo, hi) are [lo, mid) and [mid, hi), mid = lo + (hi-lo)/2;
// a node with <= 3 points is a leaf.  Nodes of one depth are independent -> one thread each.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline bool node_range(int n, int depth, int j, int& lo, int& hi) {
    lo = 0; hi = n;
    for (int b = depth - 1; b >= 0; --b) {
        if (hi - lo <= 3) return false;            // an ancestor is already a leaf
        const int mid = lo + (hi - lo) / 2;
        if ((j >> b) & 1) lo = mid; else hi = mid;
    }
    return true;
}

__host__ __device__ inline int tree_depth(int n) {
    int d = 0;
    // deepest level that still holds a node: sizes shrink as ceil(n / 2^d)
    while (((n + (1 << d) - 1) >> d) > 3) ++d;
    return d;
}

// processes node (depth, j); hulls is indexed by heap id (1 << depth) + j
__host__ __device__ inline void process_node(Mesh& m, int n, int depth, int j, Hull* hulls) {
    int lo, hi;
    if (!node_range(n, depth, j, lo, hi)) return;
    Task t;
    t.free_head = -1;
    const int id = (1 << depth) + j;
    if (hi - lo <= 3) {
        hulls[id] = build_leaf(m, t, lo, hi - lo);
    } else {
        hulls[id] = merge_hulls(m, t, hulls[2 * id], hulls[2 * id + 1]);
    }
}

// Triangle to the left of directed primal edge e, or false (outer face / not the canonical edge).
__host__ __device__ inline bool left_triangle(const Mesh& m, int e, int& a, int& b, int& c) {
    if (m.org[e & ~3] < 0) return false;                       // freed quad
    const int e2 = lnext(m, e), e3 = lnext(m, e2);
    if (lnext(m, e3) != e) return false;                       // face is not a triangle
    if (e2 < e || e3 < e) return false;                        // emit once, from its smallest edge id
    a = orgv(m, e); b = orgv(m, e2); c = orgv(m, e3);
    return orient(m.pts[a], m.pts[b], m.pts[c]) > 0;           // the unbounded face runs clockwise
}

}  // namespace dt
}  // namespace riders
