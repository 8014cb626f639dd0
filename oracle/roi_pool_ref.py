"""Oracle for SURVEY 8(a) row A1: the per-radar-point feature gather.

The reference gathers with ``torchvision.ops.roi_pool`` called five times from
``RCNet/networks.py:418-433`` (latent map at scale 1/32 and four skip maps at
1/2 .. 1/16).  torchvision is a third-party dependency (pinned 0.14.0 in the
reference's environment.yaml:9); its RoIPool algorithm restated here:

  for RoI (b, x1, y1, x2, y2) on a map (C, Hl, Wl) with spatial scale s and
  pooled size (oh, ow):
      sw = round(x1*s), sh = round(y1*s), ew = round(x2*s), eh = round(y2*s)
          (C ``round``: half away from zero; product in float32)
      rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1)
      bh = float32(rh) / float32(oh), bw = float32(rw) / float32(ow)
      bin (i, j): rows [floor(i*bh) + sh, ceil((i+1)*bh) + sh) clipped to [0, Hl],
                  cols likewise clipped to [0, Wl]; empty bin -> 0, else max,
                  first maximum in row-major scan order wins for argmax.

All bin arithmetic is float32 exactly as in the C++ kernel; the restatement
is verified bit-exact against the installed torchvision op in
tests/test_oracle_golden.py.  Test infrastructure only.
"""
import numpy as np

F32 = np.float32


def _round_half_away(v):
    # C round() on a float32 value
    v = F32(v)
    return int(np.floor(v + F32(0.5))) if v >= 0 else -int(np.floor(-v + F32(0.5)))


def roi_bins(lo_coord, hi_coord, scale, pooled, limit):
    """Start/end (clipped) input indices of every pooled bin along one axis."""
    s = F32(scale)
    start = _round_half_away(F32(lo_coord) * s)
    end = _round_half_away(F32(hi_coord) * s)
    extent = max(end - start + 1, 1)
    bin_size = F32(extent) / F32(pooled)
    idx = np.arange(pooled, dtype=np.float32)
    b0 = np.floor(idx * bin_size).astype(np.int64) + start
    b1 = np.ceil((idx + F32(1)) * bin_size).astype(np.int64) + start
    return np.clip(b0, 0, limit), np.clip(b1, 0, limit)


def roi_pool_ref(feat, rois, spatial_scale, output_size, return_argmax=False):
    """feat (B,C,H,W) float32; rois (R,5) = (batch, x1, y1, x2, y2)."""
    feat = np.asarray(feat, dtype=np.float32)
    rois = np.asarray(rois, dtype=np.float32)
    B, C, H, W = feat.shape
    oh, ow = output_size
    R = rois.shape[0]
    out = np.zeros((R, C, oh, ow), dtype=np.float32)
    arg = np.full((R, C, oh, ow), -1, dtype=np.int32)
    for r in range(R):
        b = int(rois[r, 0])
        w0, w1 = roi_bins(rois[r, 1], rois[r, 3], spatial_scale, ow, W)
        h0, h1 = roi_bins(rois[r, 2], rois[r, 4], spatial_scale, oh, H)
        for i in range(oh):
            if h1[i] <= h0[i]:
                continue
            for j in range(ow):
                if w1[j] <= w0[j]:
                    continue
                win = feat[b, :, h0[i]:h1[i], w0[j]:w1[j]]
                flat = win.reshape(C, -1)
                k = np.argmax(flat, axis=1)           # first max wins
                out[r, :, i, j] = flat[np.arange(C), k]
                ww = w1[j] - w0[j]
                arg[r, :, i, j] = (h0[i] + k // ww) * W + (w0[j] + k % ww)
    return (out, arg) if return_argmax else out


def boxes_list_to_rois(boxes_list):
    """torchvision's convert_boxes_to_roi_format: list position = batch index."""
    rows = []
    for b, boxes in enumerate(boxes_list):
        boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
        rows.append(np.concatenate([np.full((boxes.shape[0], 1), b, np.float32), boxes], 1))
    return np.concatenate(rows, 0) if rows else np.zeros((0, 5), np.float32)


def roi_pool_torchvision(feat, rois, spatial_scale, output_size):
    """The third-party op itself on the CPU (checker for the restatement)."""
    import torch
    import torchvision
    out = torchvision.ops.roi_pool(torch.as_tensor(np.asarray(feat), dtype=torch.float32),
                                   torch.as_tensor(np.asarray(rois), dtype=torch.float32),
                                   output_size=tuple(output_size), spatial_scale=float(spatial_scale))
    return out.numpy()


# RC-Net pyramid (RCNet/networks.py:404-409): scales and pooled sizes for a patch (ph, pw)
def rcnet_levels(patch_hw):
    ph, pw = patch_hw
    scales = [1 / 2.0, 1 / 4.0, 1 / 8.0, 1 / 16.0, 1 / 32.0]
    sizes = [(int(ph * s), int(pw * s)) for s in scales[:4]] + [(int(ph // 32.0), int(pw // 32.0))]
    return scales, sizes
