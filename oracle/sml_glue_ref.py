"""Oracle for the 'next' rows N2/N3/N4 of SURVEY 8(f) (NumPy float32, statement for statement).

  * ``png_roundtrip_ref``        data/data_utils.py:128-143 save_depth (uint32(z*256)) -> :94-116 load_depth (/256, <=0 -> 0)
  * ``scale_scaffold_ref``       val_zju.py:135-138, 171-179 + modules/midas/utils.py:223-237 normalize_unit_range
  * ``depth_metrics_ref``        val_zju.py:212-231 + utils/eval_utils.py:17-117
  * ``sml_resize_size_ref`` / ``sml_input_ref``   val_zju.py:181-196 + modules/midas/transforms.py:70-139 (Resize, 'minimal',
                                 cv2.INTER_NEAREST on every item), :185-193 (NormalizeIntermediate), :196-217 (PrepareForNet)
  * ``rcnet_prologue_ref``       run_rcnet_zju.py:214-240 + RCNet/rcnet_transforms.py:100-103, 258-261 + rcnet_main.py:441-444
Pinned by tests/golden/sml_glue.npz (outputs of the reference's own transforms / utils, oracle/gen_golden.py).
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


def sml_resize_size_ref(width, height, net_w=288, net_h=288, multiple_of=32):
    """transforms.py Resize.get_size, keep_aspect_ratio=True, resize_method='minimal' (:70-125)."""
    scale_height, scale_width = net_h / height, net_w / width
    if abs(1 - scale_width) < abs(1 - scale_height):
        scale_height = scale_width
    else:
        scale_width = scale_height
    new_height = (np.round(scale_height * height / multiple_of) * multiple_of).astype(int)
    new_width = (np.round(scale_width * width / multiple_of) * multiple_of).astype(int)
    return int(new_width), int(new_height)


def nearest_indices_ref(dst, src):
    """cv2.resize(..., INTER_NEAREST): inv_scale = dst / src (double); resizeNN: ofs[x] = min(floor(x * (1 / inv_scale)), src - 1)."""
    ifx = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst) * ifx).astype(np.int64), src - 1)


def sml_input_ref(image_hwc, int_depth, int_scales, net=(288, 288),
                  mean=(0.729, 0.404), std=(0.210, 0.117)):
    """One frame: image (H,W,3) float32, int_depth / int_scales (H,W) float32 -> x (3,oh,ow), d (1,oh,ow) float32."""
    image_hwc, int_depth, int_scales = (np.asarray(a, np.float32) for a in (image_hwc, int_depth, int_scales))
    H, W = int_depth.shape
    ow, oh = sml_resize_size_ref(W, H, net[0], net[1])
    iy, ix = nearest_indices_ref(oh, H)[:, None], nearest_indices_ref(ow, W)[None, :]
    img, dep, sc = image_hwc[iy, ix], int_depth[iy, ix], int_scales[iy, ix]
    dep_n = ((dep - np.float32(mean[0])) / np.float32(std[0])).astype(np.float32)
    sc_n = ((sc - np.float32(mean[1])) / np.float32(std[1])).astype(np.float32)
    chw = np.ascontiguousarray(np.transpose(img, (2, 0, 1))).astype(np.float32)
    gray = (chw[0] * np.float32(0.299) + chw[1] * np.float32(0.587)) + chw[2] * np.float32(0.114)
    return np.stack([dep_n, sc_n, gray.astype(np.float32)]), dep[None].astype(np.float32)


def rcnet_prologue_ref(image_nchw, points, patch):
    """One frame: image (1,C,H,W) float32 as loaded, points (K,3) unshifted -> (padded (1,C,H+ph,W+pw), points (K,3), boxes (K,4))."""
    image = np.asarray(image_nchw, np.float32)
    ph, pw = int(patch[0]), int(patch[1])
    px, py = pw // 2, ph // 2
    if image.max() > 1.0:
        image = image.astype(np.int32).astype(np.float32)              # Tensor.int() truncates toward zero
    image = (image / np.float32(255.0)).astype(np.float32)
    padded = np.pad(image, ((0, 0), (0, 0), (py, py), (px, px)), mode="edge")
    pts = np.asarray(points, np.float32).copy()
    pts[:, 0] = pts[:, 0] + np.float32(px)
    pts[:, 1] = pts[:, 1] + np.float32(py)
    boxes = np.stack([pts[:, 0] - np.float32(px), pts[:, 1] - np.float32(py),
                      pts[:, 0] + np.float32(px), pts[:, 1] + np.float32(py)], axis=1).astype(np.float32)
    return padded, pts, boxes
