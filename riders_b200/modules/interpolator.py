"""Sparse-to-dense scale-map interpolation, host side (mirror of the reference's
``modules/interpolator.py``: ``interpolate_knots`` :7-17, ``Interpolator2D`` :20-50).

``method='linear'`` - the only branch VI-Depth/RIDERS configure - runs on the device: knot
extraction, exact-predicate Delaunay triangulation and the per-pixel barycentric fill (fp64
math, float32 store, exactly ``fill_value=1.0`` outside the hull, like ``griddata``).
``'nearest'`` runs on the device too (exact integer distances; SciPy's choice among equidistant
knots is undefined, here the lowest knot index).  ``'cubic'`` (Clough-Tocher) is a SciPy feature
the reference never selects and raises ``NotImplementedError``.  Where SciPy raises
``QhullError`` (fewer than 3 knots, all collinear) the same exception type is raised.
"""
import numpy as np
import torch

from .. import _lib, ops

np.set_printoptions(suppress=True)

try:                                     # same exception type as the reference's call into SciPy
    from scipy.spatial import QhullError
except Exception:                        # pragma: no cover - scipy is in the image
    class QhullError(RuntimeError):
        pass


def _cuda(a, dtype):
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_cuda else t.pin_memory().cuda(non_blocking=True)


def _raise_for_status(status):
    st = status.cpu().numpy()
    if (st == _lib.RIDERS_EQHULL).any():
        raise QhullError("QH6214/QH6154 qhull input error: not enough points or initial simplex is flat "
                         "(frames %s)" % np.nonzero(st == _lib.RIDERS_EQHULL)[0].tolist())
    if (st != 0).any():
        raise _lib.RidersError("riders_interp_linear: internal status %s" % st.tolist())


def interpolate_knots(map_size, knot_coords, knot_values, interpolate, fill_corners):
    """knot_coords (2,N) = (x row, y row); returns the float64 (H,W) map like griddata."""
    if interpolate not in ('linear', 'nearest'):
        raise NotImplementedError("griddata method %r is not provided on the device path (linear, nearest)" % (interpolate,))
    coords = np.asarray(knot_coords)
    n = coords.shape[1]
    xy = _cuda(np.ascontiguousarray(coords.T.astype(np.int32))[None], torch.int32)
    val = _cuda(np.asarray(knot_values, dtype=np.float32)[None], torch.float32)
    cnt = torch.tensor([n], dtype=torch.int32, device=xy.device)
    if n == 0:
        raise QhullError("QH6214 qhull input error: not enough points(0) to construct initial simplex")
    if interpolate == 'nearest':
        return ops.interp_nearest(xy, val, cnt, int(map_size[0]), int(map_size[1]), 1.0)[0].cpu().numpy().astype(np.float64)
    out, status = ops.interp_linear(xy, val, cnt, int(map_size[0]), int(map_size[1]), 1.0)
    _raise_for_status(status)
    return out[0].cpu().numpy().astype(np.float64)


class Interpolator2D(object):
    def __init__(self, pred_inv, sparse_depth_inv, valid):
        self.pred_inv = pred_inv
        self.sparse_depth_inv = sparse_depth_inv
        self.valid = valid

        self.map_size = tuple(pred_inv.shape)
        p = _cuda(pred_inv, torch.float32)[None]
        s = _cuda(sparse_depth_inv, torch.float32)[None]
        v = _cuda(valid, torch.bool)[None]
        self.num_knots = int(v.sum().item())
        cap = max(self.num_knots, 1)
        self._xy, self._scales, self._shifts, self._cnt = ops.extract_knots(p, s, v, cap)
        xy = self._xy[0].cpu().numpy()
        self.knot_coords = np.stack((xy[:self.num_knots, 0], xy[:self.num_knots, 1])).astype(np.int64)
        self.knot_scales = self._scales[0, :self.num_knots].cpu().numpy()
        self.knot_shifts = self._shifts[0, :self.num_knots].cpu().numpy()
        self.knot_list = [(int(x), int(y)) for x, y in xy[:self.num_knots]]

        # to be computed
        self.interpolated_map = None
        self.confidence_map = None
        self.output = None

    def generate_interpolated_scale_map(self, interpolate_method, fill_corners=False):
        if interpolate_method not in ('linear', 'nearest'):
            raise NotImplementedError("griddata method %r is not provided on the device path (linear, nearest)" % (interpolate_method,))
        if interpolate_method == 'nearest':
            if self.num_knots == 0:
                raise ValueError("griddata(method='nearest') needs at least one knot")
            out = ops.interp_nearest(self._xy, self._scales, self._cnt, self.map_size[0], self.map_size[1], 1.0)
            self.interpolated_scale_map = out[0].cpu().numpy().astype(np.float32)
            return
        if self.num_knots == 0:
            raise QhullError("QH6214 qhull input error: not enough points(0) to construct initial simplex")
        out, status = ops.interp_linear(self._xy, self._scales, self._cnt, self.map_size[0], self.map_size[1], 1.0)
        _raise_for_status(status)
        self.interpolated_scale_map = out[0].cpu().numpy().astype(np.float32)
