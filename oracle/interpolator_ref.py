"""Oracle for SURVEY 8(a) row A8: sparse-to-dense scale-map interpolation.

Follows modules/interpolator.py of the reference:
  * ``Interpolator2D.__init__``   lines 21-41: knots in row-major scan order of ``valid``,
    knot_scales = sparse[valid] / pred[valid], knot_shifts = sparse[valid] - pred[valid]
  * ``interpolate_knots``         lines 7-17: scipy.interpolate.griddata(points=(x,y),
    values, xi=(col grid, row grid), method, fill_value=1.0)
  * ``generate_interpolated_scale_map`` lines 43-50: cast to float32

The triangulation itself lives in SciPy/Qhull (third-party, installed here:
reference pins scipy 1.10.1, this image has 1.18); the oracle calls it with the
reference's exact arguments.  ``delaunay_simplices`` and ``degenerate_pixel_mask``
expose what the parity tests need to mask co-circular quads.
Test infrastructure only.
"""
import numpy as np


def extract_knots_ref(pred_inv, sparse_inv, valid):
    ys, xs = np.nonzero(valid)
    coords = np.stack((xs, ys))                      # (2, N): row 0 = x (col), row 1 = y (row)
    scales = sparse_inv[valid] / pred_inv[valid]
    shifts = sparse_inv[valid] - pred_inv[valid]
    return coords, scales, shifts


def interpolate_knots_ref(map_size, knot_coords, knot_values, method="linear", fill=1.0):
    from scipy.interpolate import griddata
    rows, cols = np.mgrid[0:map_size[0], 0:map_size[1]]
    return griddata(points=knot_coords.T, values=knot_values, xi=(cols, rows),
                    method=method, fill_value=fill)


def interpolated_scale_map_ref(pred_inv, sparse_inv, valid, method="linear"):
    coords, scales, _ = extract_knots_ref(pred_inv, sparse_inv, valid)
    return interpolate_knots_ref(pred_inv.shape, coords, scales, method).astype(np.float32)


def delaunay_simplices(knot_xy):
    from scipy.spatial import Delaunay
    return Delaunay(np.asarray(knot_xy, dtype=np.float64)).simplices


def incircle_exact(a, b, c, d):
    """Exact integer in-circle determinant (>0: d strictly inside circumcircle of ccw abc)."""
    ax, ay = int(a[0]) - int(d[0]), int(a[1]) - int(d[1])
    bx, by = int(b[0]) - int(d[0]), int(b[1]) - int(d[1])
    cx, cy = int(c[0]) - int(d[0]), int(c[1]) - int(d[1])
    return ((ax * ax + ay * ay) * (bx * cy - cx * by)
            - (bx * bx + by * by) * (ax * cy - cx * ay)
            + (cx * cx + cy * cy) * (ax * by - bx * ay))


def degenerate_pixel_mask(knot_xy, H, W):
    """Pixels inside triangle pairs whose 4 corners are exactly co-circular: the
    Delaunay triangulation is not unique there and Qhull's choice is not defined
    by the reference, so parity tests exclude them (and report the fraction)."""
    from scipy.spatial import Delaunay
    pts = np.asarray(knot_xy, dtype=np.int64)
    tri = Delaunay(pts.astype(np.float64))
    mask = np.zeros((H, W), bool)
    for t, nbrs in enumerate(tri.neighbors):
        for j, n in enumerate(nbrs):
            if n < 0 or n < t:
                continue
            a, b, c = pts[tri.simplices[t]]
            opp = [v for v in tri.simplices[n] if v not in tri.simplices[t]]
            if not opp:
                continue
            d = pts[opp[0]]
            orient = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
            if orient < 0:
                b, c = c, b
            if incircle_exact(a, b, c, d) == 0:
                quad = np.array([a, b, c, d])
                x0, y0 = quad.min(0)
                x1, y1 = quad.max(0)
                x0, y0 = max(x0, 0), max(y0, 0)
                x1, y1 = min(x1, W - 1), min(y1, H - 1)
                mask[y0:y1 + 1, x0:x1 + 1] = True
    return mask
