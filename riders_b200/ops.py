"""Tensor-level wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function launches on ``torch.cuda.current_stream()`` and never synchronises.
CPU tensors are rejected - the product has no CPU path (the oracle lives in ``oracle/``
and is test infrastructure).
"""
import ctypes as C

import numpy as np

import weakref

import torch

from . import _lib
from ._lib import LAYOUT_NCHW, LAYOUT_NHWC, LOFTR_SIMT, LOFTR_TCGEN05  # noqa: F401


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class StageTimer(object):
    """Per-stage device timing for bench.py: while installed (``with StageTimer() as t``) every C-ABI
    call below is bracketed by CUDA events recorded on the launching (current) stream.  Nothing
    synchronises until ``summary()``; outside the context the wrappers cost one ``is None`` test."""

    def __init__(self):
        self.spans = {}

    def __enter__(self):
        global _timer
        self._prev, _timer = _timer, self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = self._prev

    def summary(self):
        """name -> (calls, total_ms, total algorithmic bytes or flops reported by the wrappers); synchronises."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b, _ in v), sum(w for _, _, w in v)) for k, v in self.spans.items()}


_timer = None


class _span(object):
    """``work`` = algorithmic bytes (HBM kernels) or flops (LoFTR) of the call, for bench.py's roofline."""

    def __init__(self, name, work=0):
        self.name = name
        self.work = work

    def __enter__(self):
        if _timer is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if _timer is not None and exc[0] is None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _timer.spans.setdefault(self.name, []).append((self.a, b, self.work))


def debug_set(key, value):
    """Tuning / test knobs of the library (riders_debug_set): 'lsq_fused' 1|0, 'reduce_rows', 'gather_smem_kb', ..."""
    _lib.check(_lib.lib().riders_debug_set(key.encode(), int(value)), "riders_debug_set")


def launch_count():
    """Kernels launched by libriders_b200.so in this process so far (bench.py's gpu_launches)."""
    return int(_lib.lib().riders_launch_count())


def _need_cuda(*tensors):
    """CUDA tensors only (no CPU fallback), all on the CURRENT device: the wrappers launch on
    ``torch.cuda.current_stream()`` of the current device, so a tensor living on another GPU would be addressed from the
    wrong context - refuse it loudly (wrap the call in ``with torch.cuda.device(t.device):``)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.RidersError("riders_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.RidersError("riders_b200 ops launch on the current device (cuda:%d) but got a tensor on %s; "
                                   "use `with torch.cuda.device(tensor.device):`" % (cur, t.device))


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else 0 for t in tensors])


# --------------------------------------------------------------------------------------------
# A1 gather
# --------------------------------------------------------------------------------------------
def boxes_to_rois(boxes, device):
    """list[Tensor(K,4)] (batch index = list position) or Tensor(R,5) -> (R,5) float32 on device.
    Same convention as torchvision's convert_boxes_to_roi_format."""
    if isinstance(boxes, (list, tuple)):
        rows = []
        for b, bx in enumerate(boxes):
            bx = torch.as_tensor(bx, dtype=torch.float32).reshape(-1, 4).to(device, non_blocking=True)
            rows.append(torch.cat([torch.full((bx.shape[0], 1), float(b), device=device), bx], dim=1))
        return torch.cat(rows, dim=0).contiguous() if rows else torch.zeros((0, 5), device=device)
    rois = torch.as_tensor(boxes, dtype=torch.float32).to(device)
    if rois.dim() != 2 or rois.shape[1] != 5:
        raise ValueError("rois must be (R,5) or a list of (K,4) boxes")
    return rois.contiguous()


def _is_channels_last(t):
    # ambiguous cases (C == 1 or H == W == 1) count as NCHW
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def placed_skip_buffers(R, lead_channels, skip_channels, out_sizes, device):
    """Concat buffers of the decoder, allocated BEFORE the gather: buffer l is (R, lead + C, oh, ow) channels_last
    and the returned view ``buf[:, lead:]`` is where roi_gather writes the pooled skip of level l.  The decoder
    block later finds the buffer behind the view (``placed_parent``) and writes its own ``lead`` channels in
    front - the torch.cat of utils/net_utils.py:474-569 without ever copying the skip."""
    views = []
    for lead, c, (oh, ow) in zip(lead_channels, skip_channels, out_sizes):
        buf = torch.empty((R, oh, ow, lead + c), device=device, dtype=torch.float32).permute(0, 3, 1, 2)
        views.append(buf[:, lead:])
    return views


def placed_parent(skip, lead):
    """If ``skip`` (N,C,H,W) is the trailing-channel view made by ``placed_skip_buffers`` with ``lead`` channels in
    front, return the full (N, lead + C, H, W) channels_last buffer (same storage), else None."""
    if skip is None or skip.dim() != 4 or lead <= 0 or lead % 4 or skip.dtype != torch.float32 or not skip.is_cuda:
        return None
    N, Cs, H, W = skip.shape
    ld = lead + Cs
    if skip.stride() != (H * W * ld, 1, W * ld, ld) or skip.storage_offset() < lead:
        return None
    if (skip.storage_offset() - lead) % 4 or skip.untyped_storage().nbytes() < ((skip.storage_offset() - lead) + N * H * W * ld) * 4:
        return None
    return torch.as_strided(skip, (N, ld, H, W), (H * W * ld, 1, W * ld, ld), skip.storage_offset() - lead)


def roi_gather(feats, boxes, scales, out_sizes, max_roi=None, want_argmax=False, chunk=32768, outs=None):
    """All pyramid levels of torchvision.ops.roi_pool in one call (RCNet/networks.py:418-433).

    feats: list of (B,C,H,W) float32 CUDA maps, all NCHW-contiguous or all channels_last.
    Returns list of (R,C,oh,ow) tensors in the same memory format (plus argmax list if asked).
    max_roi=(w,h): upper bound of the RoI size in image pixels (enables the TMA path).
    outs: optional pre-allocated destinations (channels_last inputs only): (R,C,oh,ow) views with channel stride 1
    and any pixel stride >= C, e.g. from ``placed_skip_buffers``.
    """
    feats = list(feats)
    _need_cuda(*feats)
    dev = feats[0].device
    nhwc = all(_is_channels_last(f) for f in feats)
    if not nhwc:
        feats = [_f32c(f) for f in feats]
    else:
        feats = [f if f.dtype == torch.float32 else f.float() for f in feats]
    rois = boxes_to_rois(boxes, dev)
    R, B = rois.shape[0], feats[0].shape[0]
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    lds = None
    if outs is not None:
        if not nhwc or want_argmax:
            raise _lib.RidersError("roi_gather(outs=...) needs channels_last maps and no argmax")
        lds = []
        for o, f, (oh, ow) in zip(outs, feats, out_sizes):
            ld = o.stride(3)
            if (tuple(o.shape) != (R, f.shape[1], oh, ow) or o.dtype != torch.float32 or o.device != dev or o.stride(1) != 1
                    or o.stride(2) != ow * ld or o.stride(0) != oh * ow * ld or ld < f.shape[1] or ld % 4 or o.data_ptr() % 16):
                raise _lib.RidersError("roi_gather(outs=...): destination %s / strides %s do not describe a channel slice of a "
                                       "channels-last map" % (tuple(o.shape), o.stride()))
            lds.append(ld)
        outs = list(outs)
    else:
        outs = [torch.empty((R, f.shape[1], oh, ow), device=dev, dtype=torch.float32, memory_format=fmt)
                for f, (oh, ow) in zip(feats, out_sizes)]
    args = [torch.empty(o.shape, device=dev, dtype=torch.int32, memory_format=fmt) for o in outs] if want_argmax else None
    ld_arr = _int_array(lds) if lds is not None else None
    n = len(feats)
    Cs = _int_array([f.shape[1] for f in feats])
    Hs = _int_array([f.shape[2] for f in feats])
    Ws = _int_array([f.shape[3] for f in feats])
    ohs = _int_array([o[0] for o in out_sizes])
    ows = _int_array([o[1] for o in out_sizes])
    sc = (C.c_float * n)(*[float(s) for s in scales])
    mw, mh = (int(max_roi[0]), int(max_roi[1])) if max_roi else (0, 0)
    L = _lib.lib()
    for r0 in range(0, R, chunk):
        r1 = min(R, r0 + chunk)
        o_ptrs = _ptr_array([o[r0:r1] for o in outs])
        a_ptrs = _ptr_array([a[r0:r1] for a in args]) if want_argmax else None
        nb = sum(o[r0:r1].numel() for o in outs) * 4 + (r1 - r0) * 20          # written patches + RoIs (maps: see bench.py)
        with _span("roi_gather", nb):
            rc = L.riders_roi_gather_fwd_placed(n, _ptr_array(feats), Cs, Hs, Ws, sc, ohs, ows,
                                                LAYOUT_NHWC if nhwc else LAYOUT_NCHW, B,
                                                _ptr(rois[r0:r1]), r1 - r0, mw, mh, o_ptrs, ld_arr, a_ptrs, _stream())
        _lib.check(rc, "riders_roi_gather_fwd_placed")
    return (outs, args) if want_argmax else outs


def roi_gather_bwd(grad_outs, argmaxes, feat_shapes, boxes, scales, nhwc):
    dev = grad_outs[0].device
    rois = boxes_to_rois(boxes, dev)
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    grad_outs = [g.contiguous(memory_format=fmt).float() for g in grad_outs]
    grads = [torch.zeros(s, device=dev, dtype=torch.float32).contiguous(memory_format=fmt) for s in feat_shapes]
    n = len(grads)
    rc = _lib.lib().riders_roi_gather_bwd(
        n, _ptr_array(grad_outs), _ptr_array(argmaxes),
        _int_array([s[1] for s in feat_shapes]), _int_array([s[2] for s in feat_shapes]),
        _int_array([s[3] for s in feat_shapes]),
        _int_array([g.shape[2] for g in grad_outs]), _int_array([g.shape[3] for g in grad_outs]),
        (C.c_float * n)(*[float(s) for s in scales]),
        LAYOUT_NHWC if nhwc else LAYOUT_NCHW, feat_shapes[0][0], _ptr(rois), rois.shape[0],
        _ptr_array(grads), _stream())
    _lib.check(rc, "riders_roi_gather_bwd")
    return grads


class RoiGatherFunction(torch.autograd.Function):
    """Differentiable multi-level gather for RC-Net training (rcnet_main.py:342-358): forward records the
    argmax of every bin, backward is the deterministic (atomic-free) scatter of riders_roi_gather_bwd."""

    @staticmethod
    def forward(ctx, rois, scales, out_sizes, *feats):
        outs, args = roi_gather(list(feats), rois, scales, out_sizes, want_argmax=True)
        ctx.save_for_backward(rois, *args)
        ctx.meta = (tuple(scales), [tuple(f.shape) for f in feats], all(_is_channels_last(f) for f in feats))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        rois, args = ctx.saved_tensors[0], list(ctx.saved_tensors[1:])
        scales, shapes, nhwc = ctx.meta
        grads = roi_gather_bwd(list(grad_outs), args, shapes, rois, scales, nhwc)
        return (None, None, None) + tuple(grads)


def roi_gather_autograd(feats, boxes, scales, out_sizes):
    """roi_gather with gradients w.r.t. the feature maps (training path)."""
    rois = boxes_to_rois(boxes, feats[0].device)
    return list(RoiGatherFunction.apply(rois, tuple(float(s) for s in scales), tuple(tuple(o) for o in out_sizes), *feats))


def roi_pool(input, boxes, output_size, spatial_scale=1.0):
    """Drop-in for ``torchvision.ops.roi_pool`` (single level)."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    return roi_gather([input], boxes, [spatial_scale], [tuple(output_size)])[0]


def upsample_nearest(x, size):
    """F.interpolate(x, size=size) (nearest) for channels-last float32 CUDA maps, bit-exact."""
    _need_cuda(x)
    if not (_is_channels_last(x) and x.dtype == torch.float32 and x.shape[1] % 4 == 0):
        raise _lib.RidersError("upsample_nearest needs a channels_last float32 map with C % 4 == 0")
    N, Cc, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((N, Cc, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    with _span("upsample_nearest", (x.numel() + out.numel()) * 4):
        rc = _lib.lib().riders_upsample_nearest_nhwc(_ptr(x), N, Cc, h, w, H, W, _ptr(out), _stream())
    _lib.check(rc, "riders_upsample_nearest_nhwc")
    return out


def bias_act_(x, bias, slope):
    """In place x = act(x + bias[c]) for a channels-last (or C == 1) float32 CUDA map; slope 0.2 leaky, 0 relu, 1 none."""
    _need_cuda(x, bias)
    N, Cc, H, W = x.shape
    if not (x.dtype == torch.float32 and (x.is_contiguous(memory_format=torch.channels_last) or (Cc == 1 and x.is_contiguous()))):
        raise _lib.RidersError("bias_act_ needs a channels_last float32 map")
    b = _f32c(bias) if bias is not None else None
    with _span("bias_act", 2 * x.numel() * 4):
        rc = _lib.lib().riders_bias_act_nhwc(_ptr(x), _ptr(b), Cc, N * H * W, float(slope), _ptr(x), _stream())
    _lib.check(rc, "riders_bias_act_nhwc")
    return x


def bias_act_pad(x, bias, slope, sy, sx):
    """act(x + bias[c]) written into a fresh map replicate-padded by one row on top (sy) / one column on the left (sx):
    (N, C, h, w) channels_last raw convolution output -> (N, C, h+sy, w+sx) channels_last; ``out[:, :, sy:, sx:]`` is the
    activated map itself.  Same bytes as ``bias_act_`` plus the duplicated row / column."""
    _need_cuda(x, bias)
    N, Cc, h, w = x.shape
    if not (x.dtype == torch.float32 and _is_channels_last(x) and Cc % 4 == 0):
        raise _lib.RidersError("bias_act_pad needs a channels_last float32 map with C % 4 == 0")
    out = torch.empty((N, Cc, h + sy, w + sx), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    b = _f32c(bias) if bias is not None else None
    with _span("bias_act", (x.numel() + out.numel()) * 4):
        rc = _lib.lib().riders_bias_act_pad_nhwc(_ptr(x), _ptr(b), Cc, N, h, w, int(sy), int(sx), float(slope), _ptr(out), _stream())
    _lib.check(rc, "riders_bias_act_pad_nhwc")
    return out


def d2s_bias_act_concat(y4, bias, skip, slope=0.2, taps=3):
    """Depth-to-space (2x2, phase-major channels) + bias + LeakyReLU (+ concat of ``skip``) in one pass.
    y4 (N, 4*Ca, h, w) channels_last [``taps=2``: (N, 4*Ca, h+1, w+1), the 2x2 phase kernels with padding 1]
    -> (N, Ca [+ Cb], 2h, 2w) channels_last.  A ``skip`` that already sits behind Ca free channels of a wider buffer
    (``placed_skip_buffers``) is not copied: only the Ca channels are written."""
    _need_cuda(y4, bias, skip)
    t = 1 if taps == 2 else 0
    N, C4 = y4.shape[:2]
    h, w = y4.shape[2] - t, y4.shape[3] - t
    Ca = C4 // 4
    if not (_is_channels_last(y4) and y4.dtype == torch.float32 and Ca % 4 == 0 and h > 0 and w > 0):
        raise _lib.RidersError("d2s_bias_act_concat needs a channels_last float32 y4 with Ca % 4 == 0")
    if skip is not None and (tuple(skip.shape[2:]) != (2 * h, 2 * w) or skip.shape[0] != N):
        raise ValueError("skip must be (N, Cb, 2h, 2w)")
    b = _f32c(bias) if bias is not None else None
    parent = placed_parent(skip, Ca)
    if parent is not None:
        with _span("d2s_bias_act_concat", (y4.numel() + N * 4 * h * w * Ca) * 4):
            rc = _lib.lib().riders_subpixel_d2s_place_nhwc(_ptr(y4), int(taps), _ptr(b), Ca, None, 0, N, h, w, float(slope),
                                                           _ptr(parent), parent.shape[1], _stream())
        _lib.check(rc, "riders_subpixel_d2s_place_nhwc")
        return parent
    Cb = skip.shape[1] if skip is not None else 0
    if skip is not None and not (_is_channels_last(skip) and skip.dtype == torch.float32 and Cb % 4 == 0):
        raise _lib.RidersError("d2s_bias_act_concat needs channels_last float32 maps with C % 4 == 0")
    out = torch.empty((N, Ca + Cb, 2 * h, 2 * w), device=y4.device, dtype=torch.float32, memory_format=torch.channels_last)
    with _span("d2s_bias_act_concat", (y4.numel() + (skip.numel() if skip is not None else 0) + out.numel()) * 4):
        rc = _lib.lib().riders_subpixel_d2s_place_nhwc(_ptr(y4), int(taps), _ptr(b), Ca, _ptr(skip), Cb, N, h, w, float(slope),
                                                       _ptr(out), Ca + Cb, _stream())
    _lib.check(rc, "riders_subpixel_d2s_place_nhwc")
    return out


def d2s_shift_bias_act_concat(y4, yfirst, bias, skip, sy, sx, slope=0.2, taps=3):
    """Sub-pixel consumer for a 2n+1 target along one axis (see upsample.cu): y4 (N,4Ca,hp,wp) [``taps=2``: (N,4Ca,hp+1,wp+1)],
    yfirst (N,2Ca,h,1) [sx] or (N,2Ca,1,w) [sy] -> (N, Ca [+ Cb], 2hp-sy, 2wp-sx) channels_last with ``skip`` concatenated
    (a placed skip is not copied)."""
    _need_cuda(y4, yfirst, bias, skip)
    t = 1 if taps == 2 else 0
    N, C4 = y4.shape[:2]
    hp, wp = y4.shape[2] - t, y4.shape[3] - t
    Ca = C4 // 4
    H, W = 2 * hp - sy, 2 * wp - sx
    for m in (y4, yfirst):
        if not (m.dtype == torch.float32 and m.is_contiguous(memory_format=torch.channels_last)):
            raise _lib.RidersError("d2s_shift needs channels_last float32 maps")
    parent = placed_parent(skip, Ca) if skip is not None else None
    if skip is not None and (tuple(skip.shape[2:]) != (H, W) or skip.shape[0] != N):
        raise ValueError("skip must be (N, Cb, %d, %d)" % (H, W))
    if parent is None:
        Cb = skip.shape[1] if skip is not None else 0
        parent = torch.empty((N, H, W, Ca + Cb), device=y4.device, dtype=torch.float32).permute(0, 3, 1, 2)
        if skip is not None:
            parent[:, Ca:] = skip
    b = _f32c(bias) if bias is not None else None
    with _span("d2s_bias_act_concat", (y4.numel() + N * H * W * Ca) * 4):
        rc = _lib.lib().riders_subpixel_d2s_shift_place_nhwc(_ptr(y4), int(taps), _ptr(yfirst), _ptr(b), Ca, N, hp, wp, int(sy), int(sx),
                                                             float(slope), _ptr(parent), parent.shape[1], _stream())
    _lib.check(rc, "riders_subpixel_d2s_shift_place_nhwc")
    return parent


_HOST_CACHE = {}


def _host_floats(t):
    """Small parameter tensor -> cached contiguous float32 host copy (one D2H per parameter version)."""
    if t is None:
        return None
    return _cached(_HOST_CACHE, t, lambda: t.detach().to("cpu", torch.float32).contiguous())


def _cached(cache, t, make):
    """Value derived from tensor ``t``, rebuilt when ``t`` is another tensor object (a recycled address does not count) or
    was modified in place (``_version``)."""
    ent = cache.get(id(t))
    if ent is not None and ent[0]() is t and ent[1] == t._version and ent[2] == t.data_ptr():
        return ent[3]
    if len(cache) > 64:
        cache.clear()
    val = make()
    cache[id(t)] = (weakref.ref(t), t._version, t.data_ptr(), val)
    return val


def conv3x3_head(x_raw, bias, slope, weight, sigmoid=False):
    """logits (N,1,H,W) = conv2d(act(x_raw + bias), weight (1,C,3,3), padding=1) for a channels-last C=16 map;
    ``sigmoid=True`` writes torch.sigmoid(logits), the confidences of RCNetModel.forward(return_logits=False)."""
    _need_cuda(x_raw)
    if not (_is_channels_last(x_raw) and x_raw.dtype == torch.float32):
        raise _lib.RidersError("conv3x3_head needs a channels_last float32 map")
    N, Cc, H, W = x_raw.shape
    hb, hw = _host_floats(bias), _host_floats(weight)
    out = torch.empty((N, 1, H, W), device=x_raw.device, dtype=torch.float32)
    with _span("conv3x3_head", (x_raw.numel() + out.numel()) * 4):
        rc = _lib.lib().riders_conv3x3_head_sigmoid_nhwc(_ptr(x_raw), N, Cc, H, W, _ptr(hb), float(slope), _ptr(hw),
                                                         1 if sigmoid else 0, _ptr(out), _stream())
    _lib.check(rc, "riders_conv3x3_head_sigmoid_nhwc")
    return out


_FRAG_CACHE = {}


def _tail_weight_fragments(weight):
    """(16,16,3,3) convolution weight -> mma.m16n8k8 B fragments [K step][lane][N half][2] on the device, rounded to TF32
    (cvt.rna), cached per parameter version: b0 = W[co = 8 nt + g][ci = 8 j + t][ky][kx], b1 the same at ci + 4,
    lane = 4 g + t, K step = (ky * 3 + kx) * 2 + j."""
    def make():
        dev = weight.device
        s = torch.arange(18, device=dev).view(18, 1, 1, 1)
        lane = torch.arange(32, device=dev).view(1, 32, 1, 1)
        nt = torch.arange(2, device=dev).view(1, 1, 2, 1)
        e = torch.arange(2, device=dev).view(1, 1, 1, 2)
        tap, j = s // 2, s % 2
        f = weight.detach().float()[nt * 8 + lane // 4, 8 * j + lane % 4 + 4 * e, tap // 3, tap % 3].contiguous()
        return ((f.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return _cached(_FRAG_CACHE, weight, make)


def _subpixel_weight_fragments(w2):
    """(64,32,2,2) 2x2-tap phase kernels -> mma.m16n8k8 B fragments [K step][N tile pair][lane][4] = {b0, b1 of N tile 2 np,
    b0, b1 of N tile 2 np + 1}, b0 = W2[n = 8 nt + g][c = 8 (ks % 4) + t][u][v], b1 the same at c + 4, tap (u, v) = ks // 4,
    TF32-rounded; cached per parameter version."""
    def make():
        dev = w2.device
        ks = torch.arange(16, device=dev).view(16, 1, 1, 1)
        npair = torch.arange(4, device=dev).view(1, 4, 1, 1)
        lane = torch.arange(32, device=dev).view(1, 1, 32, 1)
        e = torch.arange(4, device=dev).view(1, 1, 1, 4)
        tap = ks // 4
        f = w2.detach().float()[(2 * npair + e // 2) * 8 + lane // 4, (ks % 4) * 8 + lane % 4 + 4 * (e % 2), tap // 2, tap % 2].contiguous()
        return ((f.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return _cached(_FRAG_CACHE, w2, make)


_PACK_CACHE = {}


def _subpixel_tc_weights(w2):
    """(64,32,2,2) phase kernels -> four K-major SWIZZLE_128B B slabs for tcgen05 [tap = u*2+v][n][c]: 128-byte rows (32 tf32),
    8-row groups of 1024 B, 16-byte chunk index XOR (n & 7); TF32-rounded; cached per parameter version."""
    def make():
        dev = w2.device
        tap = torch.arange(4, device=dev).view(4, 1, 1)
        n = torch.arange(64, device=dev).view(1, 64, 1)
        k = torch.arange(32, device=dev).view(1, 1, 32)
        val = w2.detach().float()[n, k, tap // 2, tap % 2]
        val = ((val.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        off = tap * 2048 + (n >> 3) * 256 + (n & 7) * 32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3)        # in floats
        out = torch.empty(4 * 64 * 32, device=dev, dtype=torch.float32)
        out[off.reshape(-1)] = val.reshape(-1)
        return out
    return _cached(_PACK_CACHE, w2, make)


def subpixel_conv_tc(x, w2):
    """y4 (N,64,h+1,w+1) channels_last = conv2d(x, w2 (64,32,2,2), padding=1) in TF32 on the 5th-generation tensor cores
    (TMA-fed tcgen05.mma, accumulator in TMEM); x (N,32,h,w) channels_last, already activated, w <= 54."""
    _need_cuda(x, w2)
    N, Cc, h, w = x.shape
    if not (_is_channels_last(x) and x.dtype == torch.float32 and Cc == 32 and tuple(w2.shape) == (64, 32, 2, 2) and w <= 54):
        raise _lib.RidersError("subpixel_conv_tc: needs a channels_last float32 (N,32,h,w<=54) map and (64,32,2,2) kernels")
    y4 = torch.empty((N, 64, h + 1, w + 1), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    with _span("subpixel_conv", N * (h + 1) * (w + 1) * 128 * 64 * 2):
        rc = _lib.lib().riders_subpixel_conv_tc_nhwc(_ptr(x), _ptr(_subpixel_tc_weights(w2)), N, h, w, _ptr(y4), _stream())
    _lib.check(rc, "riders_subpixel_conv_tc_nhwc")
    return y4


def subpixel_conv_act(x_raw, bias, slope, w2):
    """y4 (N,64,h+1,w+1) channels_last = conv2d(act(x_raw + bias), w2 (64,32,2,2), padding=1) in TF32 (warp-level mma.sync),
    the bias + LeakyReLU(slope) of the convolution that produced ``x_raw`` (N,32,h,w) fused into the operand load."""
    _need_cuda(x_raw, w2)
    N, Cc, h, w = x_raw.shape
    if not (_is_channels_last(x_raw) and x_raw.dtype == torch.float32 and Cc == 32 and tuple(w2.shape) == (64, 32, 2, 2)):
        raise _lib.RidersError("subpixel_conv_act: needs a channels_last float32 (N,32,h,w) map and (64,32,2,2) kernels")
    hb = _host_floats(bias)
    y4 = torch.empty((N, 64, h + 1, w + 1), device=x_raw.device, dtype=torch.float32, memory_format=torch.channels_last)
    with _span("subpixel_conv", N * (h + 1) * (w + 1) * 128 * 64 * 2):
        rc = _lib.lib().riders_subpixel_conv_act_nhwc(_ptr(x_raw), _ptr(hb), float(slope), _ptr(_subpixel_weight_fragments(w2)),
                                                      N, h, w, _ptr(y4), _stream())
    _lib.check(rc, "riders_subpixel_conv_act_nhwc")
    return y4


def decoder_tail(y4, bias_a, weight_b, bias_b, weight_head, slope=0.2, taps=3, sigmoid=False):
    """Fused last decoder block + head (TF32): logits (N,1,2h,2w) = conv3x3(act(conv3x3(act(d2s(y4) + bias_a); weight_b)
    + bias_b); weight_head).  y4 (N,64,h,w) [taps=2: (N,64,h+1,w+1)] channels_last phase-convolution output."""
    _need_cuda(y4, weight_b)
    t = 1 if taps == 2 else 0
    N, C4 = y4.shape[:2]
    h, w = y4.shape[2] - t, y4.shape[3] - t
    if not (_is_channels_last(y4) and y4.dtype == torch.float32 and C4 == 64 and tuple(weight_b.shape) == (16, 16, 3, 3)
            and tuple(weight_head.shape) == (1, 16, 3, 3) and h > 0 and w > 0):
        raise _lib.RidersError("decoder_tail: needs a channels_last float32 (N,64,h,w) map, a (16,16,3,3) and a (1,16,3,3) weight")
    frag = _tail_weight_fragments(weight_b)
    ha, hb, hh = _host_floats(bias_a), _host_floats(bias_b), _host_floats(weight_head)
    out = torch.empty((N, 1, 2 * h, 2 * w), device=y4.device, dtype=torch.float32)
    # tensor-bound stage: the span carries FLOPs (16 -> 16 and 16 -> 1 3x3 convolutions); HBM bytes are y4 + out
    with _span("decoder_tail", N * 4 * h * w * (16 * 16 * 9 + 16 * 9) * 2):
        rc = _lib.lib().riders_decoder_tail_nhwc(_ptr(y4), int(taps), N, h, w, _ptr(ha), _ptr(frag), _ptr(hb), _ptr(hh),
                                                 float(slope), 1 if sigmoid else 0, _ptr(out), _stream())
    _lib.check(rc, "riders_decoder_tail_nhwc")
    return out


_TAILTC_CACHE = {}


def _tail_tc_weights(weight_b, weight_head):
    """(16,16,3,3) + (1,16,3,3) -> the 6144-byte fp16 B-operand image of riders_decoder_tail_tc_nhwc: per horizontal tap dx a
    K-major, unswizzled matrix [row group n // 8][K chunk c // 8][n % 8][c % 8]; rows n = 16 ky + co of Wb (48), then rows
    n = ky of Wh (16, 3 used).  Cached per version of both parameters."""
    def make():
        dev = weight_b.device
        wb, wh = weight_b.detach().float(), weight_head.detach().float()
        dx = torch.arange(3, device=dev).view(3, 1, 1, 1, 1)
        kc = torch.arange(2, device=dev).view(1, 1, 2, 1, 1)
        nr = torch.arange(8, device=dev).view(1, 1, 1, 8, 1)
        ke = torch.arange(8, device=dev).view(1, 1, 1, 1, 8)
        n = torch.arange(6, device=dev).view(1, 6, 1, 1, 1) * 8 + nr
        b = wb[n % 16, kc * 8 + ke, n // 16, dx]                                   # (3, 6, 2, 8, 8)
        n = torch.arange(2, device=dev).view(1, 2, 1, 1, 1) * 8 + nr
        hd = torch.where(n < 3, wh[0, kc * 8 + ke, n.clamp(max=2), dx], torch.zeros((), device=dev))    # (3, 2, 2, 8, 8)
        return torch.cat([b.reshape(-1), hd.reshape(-1)]).to(torch.float16).contiguous()
    ent = _TAILTC_CACHE.get(id(weight_b))
    if (ent is not None and ent[0]() is weight_b and ent[1]() is weight_head
            and ent[2] == (weight_b._version, weight_b.data_ptr(), weight_head._version, weight_head.data_ptr())):
        return ent[3]
    if len(_TAILTC_CACHE) > 64:
        _TAILTC_CACHE.clear()
    val = make()
    _TAILTC_CACHE[id(weight_b)] = (weakref.ref(weight_b), weakref.ref(weight_head),
                                   (weight_b._version, weight_b.data_ptr(), weight_head._version, weight_head.data_ptr()), val)
    return val


TAIL_MMA_SYNC, TAIL_TCGEN05 = 0, 1


class TailOverflow(object):
    """fp16-range guard of the tcgen05 decoder tail (``decoder_tail(impl=TAIL_TCGEN05)``): the kernel ORs 1 into ``flag`` when
    a staged activation left the fp16 range; the caller then re-runs the TF32 mma.sync kernel.  Same protocol as
    ``LoftrOverflow``: checked immediately, or collected in ``LoftrOverflow.pending`` with ``defer=True``."""

    def __init__(self, flag, rerun):
        self.flag, self.rerun = flag, rerun

    def overflowed(self):
        return bool(int(self.flag.item()))


def decoder_tail_tc(y4, bias_a, weight_b, bias_b, weight_head, slope=0.2, taps=3, sigmoid=False, defer=False):
    """``decoder_tail`` on tcgen05 / TMEM (fp16 operands, fp32 accumulation, csrc/decoder_tail_tc.cu).  A value outside the
    fp16 range makes the call fall back to the TF32 mma.sync kernel (immediately, or at the caller's sync point with
    ``defer=True``: the guard is appended to ``LoftrOverflow.pending``)."""
    _need_cuda(y4, weight_b)
    t = 1 if taps == 2 else 0
    N, C4 = y4.shape[:2]
    h, w = y4.shape[2] - t, y4.shape[3] - t
    if not (_is_channels_last(y4) and y4.dtype == torch.float32 and C4 == 64 and tuple(weight_b.shape) == (16, 16, 3, 3)
            and tuple(weight_head.shape) == (1, 16, 3, 3) and h > 0 and w > 0):
        raise _lib.RidersError("decoder_tail_tc: needs a channels_last float32 (N,64,h,w) map, a (16,16,3,3) and a (1,16,3,3) weight")
    pack = _tail_tc_weights(weight_b, weight_head)
    ha, hb = _host_floats(bias_a), _host_floats(bias_b)
    out = torch.empty((N, 1, 2 * h, 2 * w), device=y4.device, dtype=torch.float32)
    flag = torch.zeros(1, dtype=torch.int32, device=y4.device)
    with _span("decoder_tail", N * 4 * h * w * (16 * 16 * 9 + 16 * 9) * 2):
        rc = _lib.lib().riders_decoder_tail_tc_nhwc(_ptr(y4), int(taps), N, h, w, _ptr(ha), _ptr(pack), _ptr(hb),
                                                    float(slope), 1 if sigmoid else 0, _ptr(out), _ptr(flag), _stream())
    _lib.check(rc, "riders_decoder_tail_tc_nhwc")
    if N > 0:
        guard = TailOverflow(flag, lambda: decoder_tail(y4, bias_a, weight_b, bias_b, weight_head, slope, taps, sigmoid))
        if defer:
            LoftrOverflow.pending.append((guard, (out,)))
        elif guard.overflowed():
            return guard.rerun()
    return out


def subpixel_conv_tc_half(x, w2, bias, slope=0.2, flag=None):
    """The last block's up-convolution finished in one kernel (tcgen05 phase convolution + depth-to-space + bias + LeakyReLU
    + fp16): x (N,32,h,w) channels_last, activated -> yh (N, 2h, 2, 2w, 8) float16, the operand image ``decoder_tail_tc_half``
    reads.  ``flag``: int32 device scalar receiving the fp16-range violation bit (created when None).  Returns (yh, flag)."""
    _need_cuda(x, w2)
    N, Cc, h, w = x.shape
    if not (_is_channels_last(x) and x.dtype == torch.float32 and Cc == 32 and tuple(w2.shape) == (64, 32, 2, 2) and w <= 54):
        raise _lib.RidersError("subpixel_conv_tc_half: needs a channels_last float32 (N,32,h,w<=54) map and (64,32,2,2) kernels")
    yh = torch.empty((N, 2 * h, 2, 2 * w, 8), device=x.device, dtype=torch.float16)
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _span("subpixel_conv", N * (h + 1) * (w + 1) * 128 * 64 * 2):
        rc = _lib.lib().riders_subpixel_conv_tc_half_nhwc(_ptr(x), _ptr(_subpixel_tc_weights(w2)), N, h, w, _ptr(_host_floats(bias)),
                                                          float(slope), _ptr(yh), _ptr(flag), _stream())
    _lib.check(rc, "riders_subpixel_conv_tc_half_nhwc")
    return yh, flag


def _upconv_tc_weights(w2):
    """(4 Cout, Cin, 2, 2) 2x2-tap phase kernels -> the fp16 B-operand image of riders_upconv_tc_*_nhwc: for tap t = 2u + v and K
    step kh (16 input channels) a K-major unswizzled (4 Cout) x 16 matrix [n // 8][k // 8][n % 8][k % 8] of W2[n][16 kh + k][u][v]
    ((64,32,2,2): 16 KB; (128,64,2,2): 64 KB); cached per version."""
    def make():
        dev = w2.device
        t = torch.arange(4, device=dev).view(4, 1, 1, 1, 1, 1)
        kh = torch.arange(w2.shape[1] // 16, device=dev).view(1, -1, 1, 1, 1, 1)
        ng = torch.arange(w2.shape[0] // 8, device=dev).view(1, 1, -1, 1, 1, 1)
        kc = torch.arange(2, device=dev).view(1, 1, 1, 2, 1, 1)
        nr = torch.arange(8, device=dev).view(1, 1, 1, 1, 8, 1)
        ke = torch.arange(8, device=dev).view(1, 1, 1, 1, 1, 8)
        return w2.detach().float()[ng * 8 + nr, kh * 16 + kc * 8 + ke, t // 2, t % 2].reshape(-1).to(torch.float16).contiguous()
    return _cached(_UPCONV_CACHE, w2, make)          # (its own cache: _FRAG_CACHE holds the mma.sync fragments of the same w2)


_UPCONV_CACHE = {}


def upconv_tc_half(x, bias_p, slope_p, w2, bias_a, slope=0.2, flag=None):
    """The last block's up-convolution from the RAW output ``x`` (N,32,h,w<=54, channels_last) of the convolution in front of it:
    act_p(x + bias_p) -> tcgen05 phase convolution -> depth-to-space + bias_a + LeakyReLU -> fp16 operand map (N,2h,2,2w,8).
    ``bias_p=None, slope_p=1``: x is already activated.  Returns (yh, flag)."""
    _need_cuda(x, w2)
    N, Cc, h, w = x.shape
    if not (_is_channels_last(x) and x.dtype == torch.float32 and Cc == 32 and tuple(w2.shape) == (64, 32, 2, 2) and w <= 54):
        raise _lib.RidersError("upconv_tc_half: needs a channels_last float32 (N,32,h,w<=54) map and (64,32,2,2) kernels")
    yh = torch.empty((N, 2 * h, 2, 2 * w, 8), device=x.device, dtype=torch.float16)
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    # algorithmic bytes: the 32-channel map in, the fp16 map out
    with _span("upconv", x.numel() * 4 + yh.numel() * 2):
        rc = _lib.lib().riders_upconv_tc_half_nhwc(_ptr(x), _ptr(_host_floats(bias_p)), float(slope_p), _ptr(_upconv_tc_weights(w2)),
                                                   N, h, w, _ptr(_host_floats(bias_a)), float(slope), _ptr(yh), _ptr(flag), _stream())
    _lib.check(rc, "riders_upconv_tc_half_nhwc")
    return yh, flag


def upconv_tc_f32(x, bias_p, slope_p, w2, bias_a, slope=0.2, skip=None, flag=None):
    """deconv1 of RC-Net's decoder (64 -> 32 channels) as ONE tcgen05 kernel: act_p(x + bias_p) -> 2x2-tap phase convolution ->
    depth-to-space + bias_a + LeakyReLU, float32, written into the leading 32 channels of the block's concat buffer.
    x (N,64,h,w<=26) channels_last; ``skip``: None, or a skip placed by ``placed_skip_buffers`` (its parent buffer is the
    destination), or any (N,Cs,2h,2w) map (copied behind the result).  Returns (z (N, 32 + Cs, 2h, 2w) channels_last, flag)."""
    _need_cuda(x, w2)
    N, Cc, h, w = x.shape
    if not (_is_channels_last(x) and x.dtype == torch.float32 and Cc == 64 and tuple(w2.shape) == (128, 64, 2, 2) and w <= 26):
        raise _lib.RidersError("upconv_tc_f32: needs a channels_last float32 (N,64,h,w<=26) map and (128,64,2,2) kernels")
    parent = placed_parent(skip, 32) if skip is not None else None
    if parent is None:
        Cs = 0 if skip is None else skip.shape[1]
        if Cs % 4:
            raise _lib.RidersError("upconv_tc_f32: the skip needs a multiple of 4 channels")
        parent = torch.empty((N, 32 + Cs, 2 * h, 2 * w), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        if Cs:
            parent[:, 32:].copy_(skip)
    if tuple(parent.shape[2:]) != (2 * h, 2 * w) or parent.shape[0] != N:
        raise _lib.RidersError("upconv_tc_f32: the skip does not match the up-sampled map")
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _span("upconv", x.numel() * 4 + N * 4 * h * w * 32 * 4):
        rc = _lib.lib().riders_upconv_tc_f32_nhwc(_ptr(x), _ptr(_host_floats(bias_p)), float(slope_p), _ptr(_upconv_tc_weights(w2)),
                                                  N, h, w, _ptr(_host_floats(bias_a)), float(slope), _ptr(parent), parent.shape[1],
                                                  _ptr(flag), _stream())
    _lib.check(rc, "riders_upconv_tc_f32_nhwc")
    return parent, flag


_CONV3_CACHE = {}


def _conv3x3_tc_weights(weight):
    """(32,64,3,3) convolution weight -> the 36 KB fp16 B-operand image of riders_conv3x3_tc_f32_nhwc: for tap t = 3 ky + kx and K
    step kh a K-major unswizzled 32 x 16 matrix [n // 8][k // 8][n % 8][k % 8] of W[n][16 kh + k][ky][kx]; cached per version."""
    def make():
        dev = weight.device
        t = torch.arange(9, device=dev).view(9, 1, 1, 1, 1, 1)
        kh = torch.arange(weight.shape[1] // 16, device=dev).view(1, -1, 1, 1, 1, 1)
        ng = torch.arange(weight.shape[0] // 8, device=dev).view(1, 1, -1, 1, 1, 1)
        kc = torch.arange(2, device=dev).view(1, 1, 1, 2, 1, 1)
        nr = torch.arange(8, device=dev).view(1, 1, 1, 1, 8, 1)
        ke = torch.arange(8, device=dev).view(1, 1, 1, 1, 1, 8)
        return weight.detach().float()[ng * 8 + nr, kh * 16 + kc * 8 + ke, t // 3, t % 3].reshape(-1).to(torch.float16).contiguous()
    return _cached(_CONV3_CACHE, weight, make)


def conv3x3_tc_f32(x, weight, bias_p=None, slope_p=1.0, bias=None, slope=1.0, flag=None):
    """3x3 convolution (stride 1, padding 1) 64 -> 32 channels on tcgen05 (fp16 operands staged once, nine taps as shifted views
    of the staged rows, fp32 accumulation): out = act(conv(act_p(x + bias_p), weight) + bias), float32 channels_last; with the
    defaults the RAW convolution output of an activated map.  x (N,64,H,W<=54) channels_last.  Returns (out, flag)."""
    _need_cuda(x, weight)
    N, Cc, H, W = x.shape
    if not (_is_channels_last(x) and x.dtype == torch.float32 and Cc == 64 and tuple(weight.shape) == (32, 64, 3, 3) and W <= 54):
        raise _lib.RidersError("conv3x3_tc_f32: needs a channels_last float32 (N,64,H,W<=54) map and a (32,64,3,3) weight")
    out = torch.empty((N, 32, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _span("conv3x3_tc", N * H * W * 64 * 32 * 9 * 2):
        rc = _lib.lib().riders_conv3x3_tc_f32_nhwc(_ptr(x), _ptr(_host_floats(bias_p)), float(slope_p), _ptr(_conv3x3_tc_weights(weight)),
                                                   N, H, W, _ptr(_host_floats(bias)), float(slope), _ptr(out), 32, _ptr(flag), _stream())
    _lib.check(rc, "riders_conv3x3_tc_f32_nhwc")
    return out, flag


def decoder_tail_tc_half(yh, weight_b, bias_b, weight_head, slope=0.2, sigmoid=False, flag=None):
    """The tcgen05 decoder tail on the fp16 operand map of ``subpixel_conv_tc_half``: logits (N,1,2h,2w), flag."""
    _need_cuda(yh, weight_b)
    if not (yh.dim() == 5 and yh.dtype == torch.float16 and yh.is_contiguous() and yh.shape[2] == 2 and yh.shape[4] == 8
            and yh.shape[1] % 2 == 0 and yh.shape[3] % 2 == 0 and tuple(weight_b.shape) == (16, 16, 3, 3)
            and tuple(weight_head.shape) == (1, 16, 3, 3)):
        raise _lib.RidersError("decoder_tail_tc_half: needs a contiguous float16 (N,2h,2,2w,8) map, a (16,16,3,3) and a (1,16,3,3) weight")
    N, H, _, W, _ = yh.shape
    out = torch.empty((N, 1, H, W), device=yh.device, dtype=torch.float32)
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=yh.device)
    with _span("decoder_tail", N * H * W * (16 * 16 * 9 + 16 * 9) * 2):
        rc = _lib.lib().riders_decoder_tail_tc_half_nhwc(_ptr(yh), N, H // 2, W // 2, _ptr(_tail_tc_weights(weight_b, weight_head)),
                                                         _ptr(_host_floats(bias_b)), float(slope), 1 if sigmoid else 0, _ptr(out),
                                                         _ptr(flag), _stream())
    _lib.check(rc, "riders_decoder_tail_tc_half_nhwc")
    return out, flag


def decoder_upconv_tail_tc(x, w2, bias_a, weight_b, bias_b, weight_head, slope=0.2, sigmoid=False, defer=False,
                           bias_p=None, slope_p=1.0, producer="upconv"):
    """Last decoder block + head from the 32-channel map ``x`` (N,32,h,w) on: up-convolution kernel -> ``decoder_tail_tc_half``
    (both tcgen05 / TMEM, fp16 map between them).  ``producer``: 'upconv' (``upconv_tc_half``: operands staged once, takes the RAW
    output of the previous convolution with its ``bias_p`` / ``slope_p``) or 'tma' (``subpixel_conv_tc_half``: TMA-fed TF32 MMAs on an
    activated map).  One fp16-range flag for both kernels; when raised the result comes from the TF32 flow (cuDNN phase
    convolution + mma.sync tail) - immediately, or at the caller's sync point with ``defer=True``."""
    def activated():
        if bias_p is None and float(slope_p) == 1.0:
            return x
        return torch.nn.functional.leaky_relu(x if bias_p is None else x + bias_p.view(1, -1, 1, 1), float(slope_p))
    if producer == "upconv":
        yh, flag = upconv_tc_half(x, bias_p, slope_p, w2, bias_a, slope)
    else:
        yh, flag = subpixel_conv_tc_half(activated(), w2, bias_a, slope)
    out, _ = decoder_tail_tc_half(yh, weight_b, bias_b, weight_head, slope, sigmoid, flag=flag)
    if x.shape[0] > 0:
        def rerun():
            y4 = torch.nn.functional.conv2d(activated(), w2, None, 1, 1)
            return decoder_tail(y4, bias_a, weight_b, bias_b, weight_head, slope, 2, sigmoid)
        guard = TailOverflow(flag, rerun)
        if defer:
            LoftrOverflow.pending.append((guard, (out,)))
        elif guard.overflowed():
            return guard.rerun()
    return out


def leaky_concat(a, b, slope=0.2, act_on_a=True, bias=None):
    """cat([leaky_relu(a [+ bias]), b], dim=1) for channels-last float32 CUDA maps in one pass (bit-exact)."""
    if bias is not None:
        _need_cuda(a, b, bias)
        N, Ca, H, W = a.shape
        parent = placed_parent(b, Ca)
        if parent is not None and tuple(b.shape[2:]) == (H, W) and b.shape[0] == N and _is_channels_last(a) and Ca % 4 == 0:
            bias = _f32c(bias)
            with _span("leaky_concat", 2 * a.numel() * 4):
                rc = _lib.lib().riders_bias_leaky_place_nhwc(_ptr(a), _ptr(bias), Ca, None, 0, N * H * W, float(slope),
                                                             _ptr(parent), parent.shape[1], _stream())
            _lib.check(rc, "riders_bias_leaky_place_nhwc")
            return parent
        Cb = b.shape[1]
        out = torch.empty((N, Ca + Cb, H, W), device=a.device, dtype=torch.float32, memory_format=torch.channels_last)
        bias = _f32c(bias)
        with _span("leaky_concat", 2 * out.numel() * 4):
            rc = _lib.lib().riders_bias_leaky_concat_nhwc(_ptr(a), _ptr(bias), Ca, _ptr(b), Cb, N * H * W, float(slope),
                                                          _ptr(out), _stream())
        _lib.check(rc, "riders_bias_leaky_concat_nhwc")
        return out
    _need_cuda(a, b)
    for t in (a, b):
        if not (_is_channels_last(t) and t.dtype == torch.float32 and t.shape[1] % 4 == 0):
            raise _lib.RidersError("leaky_concat needs channels_last float32 maps with C % 4 == 0")
    N, Ca, H, W = a.shape
    Cb = b.shape[1]
    out = torch.empty((N, Ca + Cb, H, W), device=a.device, dtype=torch.float32, memory_format=torch.channels_last)
    with _span("leaky_concat", 2 * out.numel() * 4):
        rc = _lib.lib().riders_leaky_concat_nhwc(_ptr(a), Ca, _ptr(b), Cb, N * H * W, float(slope),
                                                 1 if act_on_a else 0, _ptr(out), _stream())
    _lib.check(rc, "riders_leaky_concat_nhwc")
    return out


# --------------------------------------------------------------------------------------------
# A2-A4 LoFTR
# --------------------------------------------------------------------------------------------
def linear_attention(q, k, v, q_mask=None, kv_mask=None, eps=1e-6):
    _need_cuda(q, k, v, q_mask, kv_mask)
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    N, Lq, H, D = q.shape
    S = k.shape[1]
    qm = _f32c(q_mask) if q_mask is not None else None
    km = _f32c(kv_mask) if kv_mask is not None else None
    out = torch.empty_like(q)
    rc = _lib.lib().riders_linear_attention_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(qm), _ptr(km),
                                                N, Lq, S, H, D, float(eps), _ptr(out), _stream())
    _lib.check(rc, "riders_linear_attention_fwd")
    return out


def loftr_impl_available(impl, d_model, nhead, tokens):
    return bool(_lib.lib().riders_loftr_impl_available(int(impl), int(d_model), int(nhead), int(tokens)))


_WEIGHT_ORDER = ("q_proj.weight", "k_proj.weight", "v_proj.weight", "merge.weight", "mlp.0.weight",
                 "mlp.2.weight", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")


def _weights_struct(params, keep):
    """params: dict name->tensor for one LoFTREncoderLayer (reference parameter names)."""
    w = _lib.LoftrWeights()
    for field, key in zip([f for f, _ in _lib.LoftrWeights._fields_], _WEIGHT_ORDER):
        t = _f32c(params[key].detach())
        _need_cuda(t)
        keep.append(t)
        setattr(w, field, t.data_ptr())
    return w


class LoftrOverflow(object):
    """fp16-range guard of the tcgen05 LoFTR path.  Its operands are fp16 hi/lo pairs, so every value must satisfy
    |a| <= 65504 - a limit the fp32 reference does not have.  The kernel raises a device flag when a value leaves the range;
    ``check()`` reads it (one D2H) and, if set, ``rerun()`` recomputes the call on the fp32 SIMT kernels.  ``loftr_layer`` /
    ``loftr_transformer`` check immediately unless ``defer=True`` (the batched pipeline reads all flags at its own sync point)."""
    pending = []

    def __init__(self, flag, rerun):
        self.flag, self.rerun = flag, rerun

    def overflowed(self):
        return bool(int(self.flag.item()))


def loftr_layer(x, source, params, nhead=8, impl=LOFTR_SIMT, attn_eps=1e-6, ln_eps=1e-5, defer=False):
    _need_cuda(x, source)
    x, source = _f32c(x), _f32c(source)
    N, Lx, Cm = x.shape
    S = source.shape[1]
    keep = []
    w = _weights_struct(params, keep)
    L = _lib.lib()
    ws = torch.empty(L.riders_loftr_workspace_bytes(N, max(Lx, S), Cm, impl), dtype=torch.uint8, device=x.device)
    out = torch.empty_like(x)
    rc = L.riders_loftr_layer_fwd(_ptr(x), _ptr(source), C.byref(w), N, Lx, S, Cm, nhead,
                                  float(attn_eps), float(ln_eps), impl, _ptr(out), _ptr(ws), _stream())
    _lib.check(rc, "riders_loftr_layer_fwd")
    if impl == LOFTR_TCGEN05 and N > 0:
        guard = LoftrOverflow(ws[:4].view(torch.int32), lambda: loftr_layer(x, source, params, nhead, LOFTR_SIMT, attn_eps, ln_eps))
        if defer:
            LoftrOverflow.pending.append((guard, (out,)))
        elif guard.overflowed():
            return guard.rerun()
    return out


def loftr_transformer(feat0, feat1, layer_params, layer_names, nhead=8, impl=LOFTR_SIMT,
                      attn_eps=1e-6, ln_eps=1e-5, defer=False):
    """LocalFeatureTransformer.forward (RCNet/linear_attention.py:159-184) in one C call."""
    _need_cuda(feat0, feat1)
    feat0, feat1 = _f32c(feat0), _f32c(feat1)
    N, L0, Cm = feat0.shape
    S = feat1.shape[1]
    keep = []
    n = len(layer_params)
    arr = (_lib.LoftrWeights * n)(*[_weights_struct(p, keep) for p in layer_params])
    kinds = []
    for name in layer_names:
        if name not in ("self", "cross"):
            raise KeyError(name)                      # linear_attention.py:181-182
        kinds.append(0 if name == "self" else 1)
    L = _lib.lib()
    ws = torch.empty(L.riders_loftr_workspace_bytes(N, max(L0, S), Cm, impl), dtype=torch.uint8, device=feat0.device)
    out0, out1 = torch.empty_like(feat0), torch.empty_like(feat1)
    # 7 058 688 flop per point per layer call, two calls per layer
    with _span("loftr_transformer", 2 * n * N * 7058688):
        rc = L.riders_loftr_transformer_fwd(_ptr(feat0), _ptr(feat1), arr, _int_array(kinds), n,
                                            N, L0, S, Cm, nhead, float(attn_eps), float(ln_eps), impl,
                                            _ptr(out0), _ptr(out1), _ptr(ws), _stream())
    _lib.check(rc, "riders_loftr_transformer_fwd")
    if impl == LOFTR_TCGEN05 and N > 0:
        guard = LoftrOverflow(ws[:4].view(torch.int32),
                              lambda: loftr_transformer(feat0, feat1, layer_params, layer_names, nhead, LOFTR_SIMT, attn_eps, ln_eps))
        if defer:
            LoftrOverflow.pending.append((guard, (out0, out1)))
        elif guard.overflowed():
            return guard.rerun()
    return out0, out1


def loftr_overflow_pending():
    """Read the overflow flags of every deferred tcgen05 LoFTR call since the last check (one 4-byte D2H each) and forget
    them.  True = at least one call left the fp16 range: the caller recomputes with the SIMT kernels."""
    pending, LoftrOverflow.pending = LoftrOverflow.pending, []
    if not pending:
        return False
    if len(pending) == 1:
        return pending[0][0].overflowed()
    flags = torch.stack([guard.flag.reshape(-1)[0] for guard, _ in pending])           # one kernel + one D2H for all of them
    return bool(int((flags != 0).any().item()))


_GRAD_FIELDS = ("wq", "wk", "wv", "wm", "w1", "w2", "ln1_g", "ln1_b", "ln2_g", "ln2_b")


class LoftrLayerFunction(torch.autograd.Function):
    """LoFTREncoderLayer with gradients (RC-Net training, RCNet/rcnet_main.py:342-358): forward = riders_loftr_layer_fwd_train
    (keeps q, k, v, msg, the two pre-LayerNorm tensors, m and the pre-ReLU hidden layer), backward = riders_loftr_layer_bwd
    (hand-derived, deterministic).  Tokens are float32 or bfloat16 (``x.dtype``), the ten parameters float32 master weights."""

    @staticmethod
    def forward(ctx, x, source, nhead, attn_eps, ln_eps, *params):
        _need_cuda(x, source, *params)
        dt = _lib.DTYPE_BF16 if x.dtype == torch.bfloat16 else _lib.DTYPE_F32
        tdt = torch.bfloat16 if dt == _lib.DTYPE_BF16 else torch.float32
        # a self layer passes the same tensor twice; autograd hands the Function two distinct wrappers of it, so compare storage
        same = source is x or (source.data_ptr() == x.data_ptr() and source.shape == x.shape and source.stride() == x.stride()
                               and source.dtype == x.dtype)
        x = x.to(tdt).contiguous()
        source = x if same else source.to(tdt).contiguous()
        params = tuple(_f32c(p.detach()) for p in params)
        N, Lx, Cm = x.shape
        S = source.shape[1]
        L = _lib.lib()
        w = _lib.LoftrWeights(*[p.data_ptr() for p in params])
        saved = torch.empty(L.riders_loftr_saved_bytes(N, Lx, S, Cm), dtype=torch.uint8, device=x.device)
        ws = torch.empty(L.riders_loftr_train_workspace_bytes(N, Lx, S, Cm), dtype=torch.uint8, device=x.device)
        out = torch.empty_like(x)
        rc = L.riders_loftr_layer_fwd_train(_ptr(x), _ptr(source), C.byref(w), N, Lx, S, Cm, int(nhead), float(attn_eps),
                                            float(ln_eps), dt, _ptr(out), _ptr(saved), _ptr(ws), _stream())
        _lib.check(rc, "riders_loftr_layer_fwd_train")
        ctx.save_for_backward(x, source, saved, *params)
        ctx.meta = (int(nhead), float(attn_eps), float(ln_eps), dt, same)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, source, saved = ctx.saved_tensors[:3]
        params = ctx.saved_tensors[3:]
        nhead, attn_eps, ln_eps, dt, same = ctx.meta
        N, Lx, Cm = x.shape
        S = source.shape[1]
        L = _lib.lib()
        d_out = d_out.to(x.dtype).contiguous()
        w = _lib.LoftrWeights(*[p.data_ptr() for p in params])
        grads = [torch.zeros_like(p) for p in params]
        dw = _lib.LoftrWeights(*[g.data_ptr() for g in grads])
        d_x = torch.empty_like(x)
        d_s = None if same else torch.empty_like(source)
        ws = torch.empty(L.riders_loftr_train_workspace_bytes(N, Lx, S, Cm), dtype=torch.uint8, device=x.device)
        rc = L.riders_loftr_layer_bwd(_ptr(x), _ptr(x if same else source), C.byref(w), _ptr(saved), _ptr(d_out), N, Lx, S, Cm, nhead,
                                      attn_eps, ln_eps, dt, _ptr(d_x), _ptr(d_s), C.byref(dw), _ptr(ws), _stream())
        _lib.check(rc, "riders_loftr_layer_bwd")
        return (d_x, d_s, None, None, None) + tuple(grads)


def loftr_layer_autograd(x, source, params, nhead=8, attn_eps=1e-6, ln_eps=1e-5):
    """Differentiable LoFTREncoderLayer.forward; ``params`` = dict with the reference parameter names."""
    return LoftrLayerFunction.apply(x, source, nhead, attn_eps, ln_eps, *[params[k] for k in _WEIGHT_ORDER])


# --------------------------------------------------------------------------------------------
# A5 reduce
# --------------------------------------------------------------------------------------------
def confidence_reduce(conf, points, frame_offsets, H, W, thr, logits=False, want_select=True):
    """conf (R,ph,pw) or (R,1,ph,pw); points (R,3) pad-shifted; frame_offsets (B+1) int32 tensor/list.
    Returns depth (B,H,W), response (B,H,W), select (B,H,W) int32 | None."""
    _need_cuda(conf, points)
    if conf.dim() == 4:
        conf = conf.reshape(conf.shape[0], conf.shape[-2], conf.shape[-1])
    conf, points = _f32c(conf), _f32c(points)
    dev = conf.device
    if not torch.is_tensor(frame_offsets):
        frame_offsets = torch.tensor(list(frame_offsets), dtype=torch.int32)
    offs = frame_offsets.to(device=dev, dtype=torch.int32).contiguous()
    B = offs.numel() - 1
    ph, pw = conf.shape[-2], conf.shape[-1]
    depth = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    resp = torch.empty_like(depth)
    sel = torch.empty((B, H, W), device=dev, dtype=torch.int32) if want_select else None
    with _span("confidence_reduce", conf.numel() * 4 + points.numel() * 4 + (3 if want_select else 2) * B * H * W * 4):
        rc = _lib.lib().riders_confidence_reduce(_ptr(conf), 1 if logits else 0, _ptr(points), _ptr(offs),
                                                 B, H, W, ph, pw, float(thr), _ptr(depth), _ptr(resp), _ptr(sel), _stream())
    _lib.check(rc, "riders_confidence_reduce")
    return depth, resp, sel


# --------------------------------------------------------------------------------------------
# A6/A7 alignment
# --------------------------------------------------------------------------------------------
def _binary_mask(mask):
    """The reference multiplies by ``mask`` (modules/estimator.py:10-16), and every caller passes the boolean validity map
    (val_zju.py:130-131,141).  The kernels take a 0/1 byte mask; a floating-point WEIGHT mask would be truncated to 0/1,
    so it is refused instead of being silently changed."""
    if mask.is_floating_point():
        raise _lib.RidersError("alignment kernels take a boolean / integer 0-1 validity mask; floating-point weight masks "
                               "are not supported (they would be truncated)")
    mask = mask.contiguous()
    return mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)


def lsq_align(pred, target, mask, clamp_hi=None, clamp_lo=None, want_out=True):
    """pred/target (B,H,W) float32, mask (B,H,W) bool/uint8 (all CUDA).
    Returns sums (B,5) f64, scale_shift (B,2) f64, out (B,H,W) f32 | None."""
    _need_cuda(pred, target, mask)
    pred, target = _f32c(pred), _f32c(target)
    mask = _binary_mask(mask)
    B, H, W = pred.shape
    dev = pred.device
    L = _lib.lib()
    ws = torch.empty(max(L.riders_lsq_workspace_bytes(B, H, W), 8), dtype=torch.uint8, device=dev)
    sums = torch.empty((B, 5), dtype=torch.float64, device=dev)
    ss = torch.empty((B, 2), dtype=torch.float64, device=dev)
    out = torch.empty_like(pred) if want_out else None
    flags = (1 if clamp_hi is not None else 0) | (2 if clamp_lo is not None else 0)
    # algorithmic bytes: pred + target + mask read once (9 B/pixel) + the output written (4 B/pixel); the
    # kernel's own second read of pred is traffic, not work
    with _span("lsq_align", B * H * W * (9 + (4 if want_out else 0))):
        rc = L.riders_lsq_align(_ptr(pred), _ptr(target), _ptr(mask), B, H, W, _ptr(sums), _ptr(ss), _ptr(out),
                                float(clamp_hi or 0.0), float(clamp_lo or 0.0), flags, _ptr(ws), _stream())
    _lib.check(rc, "riders_lsq_align")
    return sums, ss, out


def l1_scale_align(pred, target, mask, bounds=(0.01, 0.3), clamp_hi=None, clamp_lo=None, want_out=True,
                   max_valid=None, xatol=1e-5, maxiter=500):
    """N1: Optimizer.optimize_scale (+ apply_scale + clamp).  Returns scale (B,) f64, fval (B,) f64,
    nfev (B,) int32, out (B,H,W) f32 | None."""
    _need_cuda(pred, target, mask)
    pred, target = _f32c(pred), _f32c(target)
    mask = _binary_mask(mask)
    B, H, W = pred.shape
    dev = pred.device
    cap = int(max_valid) if max_valid else H * W
    L = _lib.lib()
    ws = torch.empty(L.riders_l1_scale_workspace_bytes(B, cap), dtype=torch.uint8, device=dev)
    scale = torch.empty((B,), dtype=torch.float64, device=dev)
    fval = torch.empty((B,), dtype=torch.float64, device=dev)
    nfev = torch.empty((B,), dtype=torch.int32, device=dev)
    out = torch.empty_like(pred) if want_out else None
    flags = (1 if clamp_hi is not None else 0) | (2 if clamp_lo is not None else 0)
    with _span("l1_scale_align", B * H * W * (9 + (4 if want_out else 0))):
        rc = L.riders_l1_scale_align(_ptr(pred), _ptr(target), _ptr(mask), B, H, W, cap,
                                     float(bounds[0]), float(bounds[1]), float(xatol), int(maxiter),
                                     _ptr(scale), _ptr(fval), _ptr(nfev), _ptr(out),
                                     float(clamp_hi or 0.0), float(clamp_lo or 0.0), flags, _ptr(ws), _stream())
    _lib.check(rc, "riders_l1_scale_align")
    return scale, fval, nfev, out


# --------------------------------------------------------------------------------------------
# N2/N3 scaffold, N4 metrics
# --------------------------------------------------------------------------------------------
def scale_scaffold(int_depth, rcnet_depth, sparse_inv, sparse_valid, min_depth=0.0, max_depth=100.0, png_quantize=True):
    """Scale scaffold of val_zju.py:171-179 (+ the PNG codec of the RC-Net -> SML hand-off).
    Returns int_scales (B,H,W) f32 in [0,1], minmax (B,2) f32, status (B,) int32 (1 = zero range)."""
    _need_cuda(int_depth, rcnet_depth, sparse_inv, sparse_valid)
    int_depth, sparse_inv = _f32c(int_depth), _f32c(sparse_inv)
    rc = _f32c(rcnet_depth) if rcnet_depth is not None else None
    sv = sparse_valid.contiguous()
    sv = sv.view(torch.uint8) if sv.dtype == torch.bool else sv.to(torch.uint8)
    B, H, W = int_depth.shape
    dev = int_depth.device
    L = _lib.lib()
    ws = torch.empty(L.riders_glue_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    out = torch.empty_like(int_depth)
    minmax = torch.empty((B, 2), dtype=torch.float32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    with _span("scale_scaffold", B * H * W * (4 + (4 if rc is not None else 0) + 4 + 1 + 4 + 8)):
        rc_ = L.riders_scale_scaffold(_ptr(int_depth), _ptr(rc), _ptr(sparse_inv), _ptr(sv), B, H, W, float(min_depth),
                                      float(max_depth), 1 if png_quantize else 0, _ptr(out), _ptr(minmax), _ptr(status),
                                      _ptr(ws), _stream())
    _lib.check(rc_, "riders_scale_scaffold")
    return out, minmax, status


def depth_metrics(pred, gt, min_val=0.0, max_val=50.0):
    """Per-frame [MAE, RMSE, iMAE, iRMSE, AbsRel, SqRel, delta1, n] (B,8) float64 (val_zju.py:212-231)."""
    _need_cuda(pred, gt)
    pred, gt = _f32c(pred), _f32c(gt)
    B, H, W = pred.shape
    L = _lib.lib()
    ws = torch.empty(L.riders_glue_workspace_bytes(B, H, W), dtype=torch.uint8, device=pred.device)
    out = torch.empty((B, 8), dtype=torch.float64, device=pred.device)
    with _span("depth_metrics", B * H * W * 8):
        rc = L.riders_depth_metrics(_ptr(pred), _ptr(gt), B, H, W, float(min_val), float(max_val), _ptr(out), _ptr(ws), _stream())
    _lib.check(rc, "riders_depth_metrics")
    return out


# --------------------------------------------------------------------------------------------
# A8 interpolation
# --------------------------------------------------------------------------------------------
def extract_knots(pred, sparse, valid, max_knots):
    """(B,H,W) maps -> knot_xy (B,max_knots,2) int32, scales, shifts (B,max_knots) f32, counts (B) int32."""
    _need_cuda(pred, sparse, valid)
    pred, sparse = _f32c(pred), _f32c(sparse)
    valid = valid.contiguous()
    valid = valid.view(torch.uint8) if valid.dtype == torch.bool else valid.to(torch.uint8)
    B, H, W = pred.shape
    dev = pred.device
    xy = torch.zeros((B, max_knots, 2), dtype=torch.int32, device=dev)
    sc = torch.zeros((B, max_knots), dtype=torch.float32, device=dev)
    sh = torch.zeros((B, max_knots), dtype=torch.float32, device=dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    with _span("extract_knots", B * H * W + 20 * B * max_knots):
        rc = _lib.lib().riders_extract_knots(_ptr(pred), _ptr(sparse), _ptr(valid), B, H, W, max_knots,
                                             _ptr(xy), _ptr(sc), _ptr(sh), _ptr(cnt), _stream())
    _lib.check(rc, "riders_extract_knots")
    return xy, sc, sh, cnt


def interp_linear(knot_xy, knot_val, knot_count, H, W, fill=1.0, want_triangles=False):
    """knot_xy (B,max_knots,2) int32 (x,y); knot_val (B,max_knots) f32; knot_count (B) int32.
    Returns out (B,H,W) f32, status (B) int32 [, triangles (B,2*max_knots,3) int32, tri_count (B)]."""
    _need_cuda(knot_xy, knot_val, knot_count)
    knot_xy = knot_xy.to(torch.int32).contiguous()
    knot_val = _f32c(knot_val)
    knot_count = knot_count.to(torch.int32).contiguous()
    B, max_knots = knot_val.shape
    dev = knot_val.device
    L = _lib.lib()
    ws = torch.empty(L.riders_interp_workspace_bytes(B, max_knots, H, W), dtype=torch.uint8, device=dev)
    out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    tri = torch.full((B, 2 * max_knots, 3), -1, dtype=torch.int32, device=dev) if want_triangles else None
    tcount = torch.zeros((B,), dtype=torch.int32, device=dev) if want_triangles else None
    # one C call = two launches (stars + fill value, rasteriser); algorithmic bytes: the output map + 12 B per knot
    with _span("interp_linear", B * H * W * 4 + B * max_knots * 12):
        rc = L.riders_interp_linear(_ptr(knot_xy), _ptr(knot_val), _ptr(knot_count), B, max_knots, H, W, float(fill),
                                    _ptr(out), _ptr(status), _ptr(tri), _ptr(tcount), _ptr(ws), _stream())
    _lib.check(rc, "riders_interp_linear")
    if want_triangles:
        return out, status, tri, tcount
    return out, status


def interp_nearest(knot_xy, knot_val, knot_count, H, W, fill=1.0):
    """griddata(method='nearest') on the device: (B,max_knots,2) int32 knots -> (B,H,W) float32."""
    _need_cuda(knot_xy, knot_val, knot_count)
    knot_xy = knot_xy.to(torch.int32).contiguous()
    knot_val = _f32c(knot_val)
    knot_count = knot_count.to(torch.int32).contiguous()
    B, max_knots = knot_val.shape
    out = torch.empty((B, H, W), dtype=torch.float32, device=knot_val.device)
    with _span("interp_nearest", B * H * W * 4 + B * max_knots * 12):
        rc = _lib.lib().riders_interp_nearest(_ptr(knot_xy), _ptr(knot_val), _ptr(knot_count), B, max_knots, H, W, float(fill),
                                              _ptr(out), _stream())
    _lib.check(rc, "riders_interp_nearest")
    return out


# --------------------------------------------------------------------------------------------
# N4 prologue / N2 SML input assembly
# --------------------------------------------------------------------------------------------
def rcnet_prologue(images, points, frame_offsets, patch, channels_last=True):
    """run_rcnet_zju.py:214-240 + rcnet_transforms normalisation + rcnet_main.py:441-444 edge pad, one pass.
    images (B,C,H,W) float32 as loaded ([0,255] or [0,1]); points (R,3) UNSHIFTED; patch (ph, pw).
    Returns padded normalised images (B,C,H+ph,W+pw) [channels_last], pad-shifted points (R,3), rois (R,5)."""
    _need_cuda(images, points)
    images, points = _f32c(images), _f32c(points)
    dev = images.device
    B, Cc, H, W = images.shape
    ph, pw = int(patch[0]), int(patch[1])
    px, py = pw // 2, ph // 2
    if not torch.is_tensor(frame_offsets):
        frame_offsets = torch.tensor(list(frame_offsets), dtype=torch.int32)
    offs = frame_offsets.to(device=dev, dtype=torch.int32).contiguous()
    R = points.shape[0]
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    padded = torch.empty((B, Cc, H + 2 * py, W + 2 * px), device=dev, dtype=torch.float32, memory_format=fmt)
    shifted = torch.empty((R, 3), device=dev, dtype=torch.float32)
    rois = torch.empty((R, 5), device=dev, dtype=torch.float32)
    flags = torch.empty(B, device=dev, dtype=torch.int32)
    with _span("rcnet_prologue", (images.numel() * 2 + padded.numel() + R * 11) * 4):
        rc = _lib.lib().riders_rcnet_prologue(_ptr(images), B, Cc, H, W, px, py, LAYOUT_NHWC if channels_last else LAYOUT_NCHW,
                                              _ptr(points), _ptr(offs), R, _ptr(padded), _ptr(shifted), _ptr(rois),
                                              _ptr(flags), _stream())
    _lib.check(rc, "riders_rcnet_prologue")
    return padded, shifted, rois


def sml_resize_size(width, height, net_w=288, net_h=288, multiple_of=32):
    """modules/midas/transforms.py Resize.get_size with keep_aspect_ratio=True, resize_method='minimal' (:70-125)."""
    scale_h, scale_w = net_h / height, net_w / width
    if abs(1 - scale_w) < abs(1 - scale_h):
        scale_h = scale_w
    else:
        scale_w = scale_h
    new_h = int(np.round(scale_h * height / multiple_of) * multiple_of)
    new_w = int(np.round(scale_w * width / multiple_of) * multiple_of)
    return new_w, new_h


SML_MEAN = {"int_depth": 0.729, "int_scales": 0.404}          # modules/midas/transforms.py:333-334
SML_STD = {"int_depth": 0.210, "int_scales": 0.117}


def sml_input_assemble(int_depth, int_scales, image, image_hwc=False, net=(288, 288)):
    """val_zju.py:181-196: the SML network's inputs x (B,3,oh,ow) and d (B,1,oh,ow) from the aligned prior
    (B,H,W), the scale scaffold (B,H,W) and the frame ((B,3,H,W), or (B,H,W,3) with image_hwc)."""
    _need_cuda(int_depth, int_scales, image)
    int_depth, int_scales, image = _f32c(int_depth), _f32c(int_scales), _f32c(image)
    B, H, W = int_depth.shape
    ow, oh = sml_resize_size(W, H, net[0], net[1])
    x = torch.empty((B, 3, oh, ow), device=int_depth.device, dtype=torch.float32)
    d = torch.empty((B, 1, oh, ow), device=int_depth.device, dtype=torch.float32)
    with _span("sml_input_assemble", B * oh * ow * 9 * 4):
        rc = _lib.lib().riders_sml_input_assemble(_ptr(int_depth), _ptr(int_scales), _ptr(image), 1 if image_hwc else 0,
                                                  B, H, W, oh, ow, SML_MEAN["int_depth"], SML_STD["int_depth"],
                                                  SML_MEAN["int_scales"], SML_STD["int_scales"], _ptr(x), _ptr(d), _stream())
    _lib.check(rc, "riders_sml_input_assemble")
    return x, d
