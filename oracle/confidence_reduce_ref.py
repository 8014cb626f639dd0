"""Oracle for SURVEY 8(a) row A5: thresholded confidence -> quasi-dense depth.

Follows the tail of ``forward_output`` (RCNet/rcnet_main.py:460-487):
  :462  c = where(c < thr, 0, c)            (strict <, so c == thr survives)
  :463  paste the (ph, pw) crop into a zero canvas of the PADDED frame at rows
        int(y)-ph/2 .. int(y)+ph/2, cols int(x)-pw/2 .. int(x)+pw/2 where (x, y)
        are the pad-shifted point coordinates
  :466  crop the pad away  -> tile k covers unpadded rows y0-ph/2 .. y0+ph/2
  :468  response, select = max over k  (first maximum wins)
  :478  depth = sum_k tile_k * z_k / sum_k tile_k
  :482  depth = 0 where response == 0

Vectorised per point instead of materialising the (K, H', W') tensor; sums
run in point order in float32 (float64 variant available for tolerance
budgeting).  Test infrastructure only.
"""
import numpy as np


def confidence_reduce_ref(conf, points, H, W, thr, acc_dtype=np.float32):
    """conf (K, ph, pw) f32 confidences; points (K,3) pad-shifted (x, y, z).

    Returns depth (H,W) f32, response (H,W) f32, select (H,W) int32.
    """
    conf = np.asarray(conf, dtype=np.float32)
    points = np.asarray(points, dtype=np.float32)
    K, ph, pw = conf.shape
    pad_y, pad_x = ph // 2, pw // 2
    response = np.zeros((H, W), np.float32)
    select = np.zeros((H, W), np.int32)
    sum_c = np.zeros((H, W), acc_dtype)
    sum_cz = np.zeros((H, W), acc_dtype)
    for k in range(K):
        t = np.where(conf[k] < np.float32(thr), np.float32(0), conf[k])
        # unpadded top-left corner of the pasted crop
        r0 = int(points[k, 1]) - 2 * pad_y
        c0 = int(points[k, 0]) - 2 * pad_x
        ra, rb = max(r0, 0), min(r0 + ph, H)
        ca, cb = max(c0, 0), min(c0 + pw, W)
        if ra >= rb or ca >= cb:
            continue
        tt = t[ra - r0:rb - r0, ca - c0:cb - c0]
        view = (slice(ra, rb), slice(ca, cb))
        better = tt > response[view]
        select[view] = np.where(better, k, select[view])
        response[view] = np.where(better, tt, response[view])
        sum_c[view] += tt.astype(acc_dtype)
        sum_cz[view] += (tt * points[k, 2]).astype(acc_dtype) if acc_dtype == np.float32 \
            else tt.astype(acc_dtype) * acc_dtype(points[k, 2])
    with np.errstate(invalid="ignore", divide="ignore"):
        depth = (sum_cz / sum_c).astype(np.float32)
    depth = np.where(response == 0, np.float32(0), depth).astype(np.float32)
    return depth, response, select


def top2_margin(conf, points, H, W, thr):
    """Gap between best and second-best thresholded confidence per pixel (for the
    'select is bit-exact where margin > 1e-6' rule of BASELINE.json)."""
    conf = np.asarray(conf, dtype=np.float32)
    K, ph, pw = conf.shape
    pad_y, pad_x = ph // 2, pw // 2
    best = np.zeros((H, W), np.float32)
    second = np.zeros((H, W), np.float32)
    for k in range(K):
        t = np.where(conf[k] < np.float32(thr), np.float32(0), conf[k])
        r0 = int(points[k, 1]) - 2 * pad_y
        c0 = int(points[k, 0]) - 2 * pad_x
        ra, rb = max(r0, 0), min(r0 + ph, H)
        ca, cb = max(c0, 0), min(c0 + pw, W)
        if ra >= rb or ca >= cb:
            continue
        tt = t[ra - r0:rb - r0, ca - c0:cb - c0]
        v = (slice(ra, rb), slice(ca, cb))
        nb = np.maximum(best[v], tt)
        second[v] = np.maximum(second[v], np.minimum(best[v], tt))
        best[v] = nb
    return best - second
