// Host build of the per-site star Delaunay core (riders_b200/csrc/delaunay_star.cuh) for tests/test_delaunay_host.py.
#include "../riders_b200/csrc/delaunay_star.cuh"
#include <vector>
using namespace riders;

// xy: n x 2 int32 (x, y); tri: up to 2n x 3 int32 knot indices; info[0] = star overflow, info[1] = max star size.
extern "C" int delaunay_star_host(int n, const int* xy, int* tri, int* info) {
    std::vector<dt::Pt> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * i]; pts[i].y = xy[2 * i + 1]; }
    info[0] = 0; info[1] = 0;
    int nt = 0;
    star::Star buf[2];
    for (int p = 0; p < n; ++p) {
        const int cur = star::build_star(pts.data(), n, p, buf);
        if (cur < 0) { info[0] = 1; continue; }
        const star::Star& s = buf[cur];
        if (s.k > info[1]) info[1] = s.k;
        for (int j = 0; j < s.k && s.k >= 2; ++j) {
            const int m = j + 1 == s.k ? 0 : j + 1;
            if (star::cross2(s.x[j], s.y[j], s.x[m], s.y[m]) > 0 && p < s.id[j] && p < s.id[m]) {
                if (nt < 2 * n) { tri[3 * nt] = p; tri[3 * nt + 1] = s.id[j]; tri[3 * nt + 2] = s.id[m]; }
                ++nt;
            }
        }
    }
    return nt;
}
