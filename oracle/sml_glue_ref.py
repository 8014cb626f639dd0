"""Oracle for the 'next' rows N2/N3/N4 of SURVEY 8(f) (NumPy float32, statement for statement).

  * ``png_roundtrip_ref``        data/data_utils.py:128-143 save_depth (uint32(z*256)) -> :94-116 load_depth (/256, <=0 -> 0)
  * ``scale_scaffold_ref``       val_zju.py:135-138, 171-179 + modules/midas/utils.py:223-237 normalize_unit_range
  * ``depth_metrics_ref``        val_zju.py:212-231 + utils/eval_utils.py:17-117
Test infrastructure only.
"""
import numpy as np


def png_roundtrip_ref(z):
    q = np.uint32(np.asarray(z, np.float32) * 256.0).astype(np.float32) / np.float32(256.0)
    q[q <= 0] = 0.0
    return q


def normalize_unit_range_ref(data):
    if np.max(data) - np.min(data) > np.finfo("float").eps:
        return (data - np.min(data)) / (np.max(data) - np.min(data))
    raise ValueError("cannot normalize array, max-min range is 0")


def scale_scaffold_ref(int_depth, rcnet_depth, sparse_inv, sparse_valid, min_depth=0.0, max_depth=100.0, png=True):
    int_depth = np.asarray(int_depth, np.float32)
    int_scales = np.ones_like(int_depth)
    if rcnet_depth is not None:
        rc = png_roundtrip_ref(rcnet_depth) if png else np.asarray(rcnet_depth, np.float32).copy()
        rc_valid = (rc < np.float32(max_depth)) & (rc > np.float32(min_depth))
        rc[~rc_valid] = np.inf
        with np.errstate(divide="ignore"):
            rc = (np.float32(1.0) / rc).astype(np.float32)
        int_scales[rc_valid] = rc[rc_valid] / int_depth[rc_valid]
    sv = np.asarray(sparse_valid, bool)
    int_scales[sv] = np.asarray(sparse_inv, np.float32)[sv] / int_depth[sv]
    return normalize_unit_range_ref(int_scales.astype(np.float32))


def depth_metrics_ref(pred, gt, min_val=0.0, max_val=50.0):
    """Returns [mae, rmse, imae, irmse, abs_rel, sq_rel, delta1, n] like val_zju.py:212-231."""
    pred, gt = np.asarray(pred, np.float32), np.asarray(gt, np.float32)
    mask = np.where((gt > 0) & (gt > min_val) & (gt < max_val))
    o, g = pred[mask], gt[mask]
    src, tgt = 1000.0 * o, 1000.0 * g
    isrc, itgt = 0.001 * o, 0.001 * g
    return np.array([
        np.mean(np.abs(tgt - src)), np.sqrt(np.mean((tgt - src) ** 2)),
        np.mean(np.abs((1.0 / itgt) - (1.0 / isrc))), np.sqrt(np.mean(((1.0 / itgt) - (1.0 / isrc)) ** 2)),
        np.mean(np.abs(src - tgt) / tgt), np.mean(((src - tgt) ** 2) / tgt),
        np.mean(np.maximum((g / o), (o / g)) < 1.25), float(o.size)], dtype=np.float64)
