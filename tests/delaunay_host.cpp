// Host build of riders_b200/csrc/delaunay_core.cuh for CPU unit tests (compiled by
// tests/test_delaunay_host.py with g++; not part of the product library).
#include <algorithm>
#include <vector>
#include <cstring>
#include "../riders_b200/csrc/delaunay_core.cuh"

using namespace riders::dt;

extern "C" int delaunay_host(int n, const int* xy, int* tri /* 2n*3 */, int* info /* [overflow, quads_used] */) {
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        return xy[2 * a] != xy[2 * b] ? xy[2 * a] < xy[2 * b] : xy[2 * a + 1] < xy[2 * b + 1];
    });
    std::vector<Pt> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * order[i]]; pts[i].y = xy[2 * order[i] + 1]; }
    const int cap = 8 * n + 64;
    std::vector<int> next(4 * cap, 0), org(4 * cap, -1);
    int fresh = 0, overflow = 0;
    Mesh m{next.data(), org.data(), &fresh, cap, pts.data(), &overflow};
    const int D = tree_depth(n);
    std::vector<Hull> hulls((size_t)2 << D);
    for (int d = D; d >= 0; --d)
        for (int j = 0; j < (1 << d); ++j) process_node(m, n, d, j, hulls.data());
    int nt = 0;
    for (int q = 0; q < fresh && q < cap; ++q)
        for (int r = 0; r < 4; r += 2) {
            int a, b, c;
            if (left_triangle(m, 4 * q + r, a, b, c)) {
                tri[3 * nt] = order[a]; tri[3 * nt + 1] = order[b]; tri[3 * nt + 2] = order[c];
                ++nt;
            }
        }
    info[0] = overflow; info[1] = fresh;
    return nt;
}
