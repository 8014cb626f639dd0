"""The exact-arithmetic Delaunay cores compiled for the host and checked against scipy.spatial.Delaunay (Qhull, what the
reference's griddata uses): the per-site star construction the device kernel runs (riders_b200/csrc/delaunay_star.cuh, one
warp per knot) and the round-1 divide-and-conquer core it replaced (delaunay_core.cuh, kept for its exact predicates)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.interpolator_ref import incircle_exact

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["star", "dnc"])
def dt_lib(request, tmp_path_factory):
    src, sym = {"star": ("delaunay_star_host.cpp", "delaunay_star_host"), "dnc": ("delaunay_host.cpp", "delaunay_host")}[request.param]
    out = str(tmp_path_factory.mktemp("dt") / ("lib%s.so" % sym))
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", src), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    fn = getattr(lib, sym)
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.triangulate, lib.kind = fn, request.param
    return lib


def triangulate(lib, xy):
    xy = np.ascontiguousarray(xy, dtype=np.int32)
    n = xy.shape[0]
    tri = np.full((2 * n, 3), -1, np.int32)
    info = np.zeros(2, np.int32)
    nt = lib.triangulate(n, xy.ctypes.data, tri.ctypes.data, info.ctypes.data)
    assert info[0] == 0, "edge pool / star overflow"
    return tri[:nt], info


def canon(tris):
    return set(tuple(sorted(int(v) for v in t)) for t in tris)


def unique_pixels(rng, n, H, W):
    flat = rng.choice(H * W, n, replace=False)
    return np.stack([flat % W, flat // W], 1).astype(np.int32)


def is_delaunay(xy, tris):
    """every triangle ccw and no other vertex strictly inside its circumcircle (exact)."""
    for a, b, c in tris:
        A, B, Cc = xy[a], xy[b], xy[c]
        if (int(B[0]) - int(A[0])) * (int(Cc[1]) - int(A[1])) - (int(B[1]) - int(A[1])) * (int(Cc[0]) - int(A[0])) <= 0:
            return False
        for d in range(len(xy)):
            if d in (a, b, c):
                continue
            if incircle_exact(A, B, Cc, xy[d]) > 0:
                return False
    return True


@pytest.mark.parametrize("n,H,W,seed", [(3, 24, 32, 0), (4, 16, 16, 1), (5, 64, 64, 2), (17, 64, 80, 3),
                                         (200, 512, 640, 4), (512, 512, 640, 5), (4096, 1024, 1280, 6)])
def test_matches_qhull_where_unique(dt_lib, n, H, W, seed):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    xy = unique_pixels(rng, n, H, W)
    tris, info = triangulate(dt_lib, xy)
    ref = Delaunay(xy.astype(np.float64))
    mine, theirs = canon(tris), canon(ref.simplices)
    # Euler: same triangle count even when diagonals of co-circular quads differ
    assert len(mine) == len(theirs)
    diff = mine ^ theirs
    # every disagreement must involve an exactly co-circular quad (non-unique triangulation)
    for t in diff:
        src = mine if t in mine else theirs
        other = theirs if t in mine else mine
        a, b, c = t
        A, B, Cc = xy[a], xy[b], xy[c]
        if (int(B[0]) - int(A[0])) * (int(Cc[1]) - int(A[1])) - (int(B[1]) - int(A[1])) * (int(Cc[0]) - int(A[0])) < 0:
            B, Cc = Cc, B
        on_circle = [d for d in range(n) if d not in t and incircle_exact(A, B, Cc, xy[d]) == 0]
        assert on_circle, "triangle %s differs from Qhull without a co-circular knot" % (t,)
    assert len(diff) <= max(8, n // 20)
    if dt_lib.kind == "dnc":
        assert info[1] <= 4 * n + 8        # pool of 8n quads is ample
    else:
        assert info[1] <= 24               # largest star of a random set stays far below STAR_MAX = 64


@pytest.mark.parametrize("seed", range(5))
def test_small_sets_are_delaunay_exactly(dt_lib, seed):
    rng = np.random.default_rng(100 + seed)
    xy = unique_pixels(rng, 40, 12, 12)       # dense grid: many collinear / co-circular configurations
    tris, _ = triangulate(dt_lib, xy)
    assert is_delaunay(xy, tris)
    from scipy.spatial import ConvexHull
    area2 = sum(abs((int(xy[b][0]) - int(xy[a][0])) * (int(xy[c][1]) - int(xy[a][1])) -
                    (int(xy[b][1]) - int(xy[a][1])) * (int(xy[c][0]) - int(xy[a][0]))) for a, b, c in tris)
    assert abs(area2 / 2.0 - ConvexHull(xy.astype(float)).volume) < 1e-6      # triangles tile the hull


def test_degenerate_inputs(dt_lib):
    line = np.stack([np.arange(7) * 3, np.arange(7) * 2], 1)
    tris, _ = triangulate(dt_lib, line)
    assert len(tris) == 0                                        # all collinear -> no triangle (QhullError upstream)
    grid = np.stack(np.meshgrid(np.arange(6), np.arange(5)), -1).reshape(-1, 2)
    tris, _ = triangulate(dt_lib, grid)
    assert len(tris) == 2 * 5 * 4 and is_delaunay(grid, tris)   # regular grid: every quad is co-circular
    two = np.array([[1, 1], [5, 9]])
    assert len(triangulate(dt_lib, two)[0]) == 0
