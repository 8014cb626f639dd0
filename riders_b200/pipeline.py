"""Batched radar-to-dense-depth chain: the public entry point of the B200 path.

One call = the reference's two-script flow for a batch of frames, without the PNG round trip:

  run_rcnet_zju.py:204-263   RC-Net quasi-dense depth   -> forward_output_batched (A0-A5)
  val_zju.py:130-133         sparse radar depth -> inverse depth + validity mask
  val_zju.py:146-155         global scale/shift LSQ alignment of the mono prior (A6/A7)
  modules/interpolator.py    scale scaffold: knots = target/aligned, Delaunay + barycentric (A8)

``run_device`` takes device tensors; ``run_host`` is the end-to-end form (pinned host buffers in,
pinned host buffers out, copies on the current stream) that bench.py's ``e2e`` figure times.
The SML network itself (MidasNet_small_videpth) is not constructible offline (torch.hub encoder,
modules/midas/blocks.py:44-51) and stays the reference's torch module - any ``nn.Module`` with
``forward(x, d)`` can be attached by the caller after this front-end.
"""
import contextlib

import numpy as np
import torch

from . import ops
from .rcnet.rcnet_main import forward_output_batched


class RadarDepthPipeline(object):
    def __init__(self, model, response_thr=0.1, min_depth=0.0, max_depth=100.0,
                 clamp_min=0.1, clamp_max=255.0, chunk_points=2048, max_knots=None,
                 raw_inputs=False, retry_empty=False, sml_input=False, overlap_front_end=True,
                 global_alignment='st', mono_type='inv', return_select=False, sml_model=None, host_keys=None):
        """raw_inputs: frames as loaded ([0,255]) and radar points in frame coordinates (the prologue kernel does the
        shift / boxes / normalisation / pad of run_rcnet_zju.py:214-240); retry_empty: the script's threshold-lowering
        retry for frames where nothing fires (:254-265); sml_input: also assemble the SML network's inputs (x, d) of
        val_zju.py:181-196 (resize to 288-multiple-of-32, normalise, gray channel).
        global_alignment: 'st' = least-squares scale + shift (val_zju.py:146-155), 's' = the scripts' default, the bounded-Brent
        L1 scale of ``Optimizer`` (val_zju.py:156-167, bounds by ``mono_type`` 'inv' | 'pos').
        sml_model: any ``nn.Module`` with ``forward(x, d)`` (the reference's MidasNet_small_videpth is not constructible
        offline, modules/midas/blocks.py:44-51): when given it is fed the assembled inputs and its prediction is inverted and
        bicubically resized to the frame like val_zju.py:198-210 (``out['sml_pred']``); without it the chain ends at the
        network's inputs (``sml_x``, ``sml_d``).
        host_keys: which results the host entry points copy back (default: the maps a caller consumes on the host; the SML
        inputs stay on the device unless asked for)."""
        if global_alignment not in ('st', 's'):
            raise NotImplementedError(global_alignment)                 # val_zju.py:168-169
        self.global_alignment, self.mono_type = global_alignment, mono_type
        self.return_select = bool(return_select)
        self.sml_model = sml_model
        if sml_model is not None:
            sml_input = True
        self.host_keys = tuple(host_keys) if host_keys is not None else None
        self.raw_inputs, self.retry_empty, self.sml_input = bool(raw_inputs), bool(retry_empty), bool(sml_input)
        self.overlap_front_end = bool(overlap_front_end)
        self._side = None
        self.model = model
        self.response_thr = float(response_thr)
        self.min_depth, self.max_depth = float(min_depth), float(max_depth)
        self.clamp_hi = float(np.float32(1.0 / clamp_min)) if clamp_min and clamp_min > 0 else None
        self.clamp_lo = float(np.float32(1.0 / clamp_max)) if clamp_max else None
        self.chunk_points = chunk_points
        self.max_knots = max_knots
        self._pinned = {}
        self._copy_streams = None
        self._slot = 0

    # ------------------------------------------------------------------------------ SML front-end
    @torch.no_grad()
    def align_and_interpolate(self, mono_pred, sparse_depth, max_knots):
        """The RC-Net-independent half of the SML front-end for a batch of frames (BASELINE configs[0]): inverse sparse
        depth + validity (val_zju.py:130-133), global alignment (:146-167), knot extraction + Delaunay + barycentric scale
        map (modules/interpolator.py).  Knot capacity: the knots are the valid pixels of ``sparse_depth``; a frame with
        more of them than ``max_knots`` is NOT truncated silently - its ``interp_status`` is RIDERS_EINTERNAL and its map
        the fill value."""
        valid = (sparse_depth < self.max_depth) & (sparse_depth > self.min_depth)
        sparse_inv = torch.where(valid, 1.0 / sparse_depth, torch.zeros_like(sparse_depth))
        if self.global_alignment == 'st':
            _, scale_shift, int_depth = ops.lsq_align(mono_pred, sparse_inv, valid,
                                                      clamp_hi=self.clamp_hi, clamp_lo=self.clamp_lo)
        else:
            bounds = (0.01, 0.3) if self.mono_type == 'inv' else (0.5, 1.6)               # modules/estimator.py:147-150
            scale, _, _, int_depth = ops.l1_scale_align(mono_pred, sparse_inv, valid, bounds=bounds,
                                                        clamp_hi=self.clamp_hi, clamp_lo=self.clamp_lo)
            scale_shift = torch.stack([scale, torch.zeros_like(scale)], dim=1)
        xy, scales, _, cnt = ops.extract_knots(int_depth, sparse_inv, valid, int(max_knots))
        scale_map, status = ops.interp_linear(xy, scales, cnt, int_depth.shape[-2], int_depth.shape[-1], 1.0)
        return {"valid": valid, "sparse_inv": sparse_inv, "scale_shift": scale_shift, "int_depth": int_depth,
                "knot_count": cnt, "scale_map": scale_map, "interp_status": status}

    @torch.no_grad()
    def sml_front_end(self, mono_pred, sparse_depth, rcnet_depth, image, max_knots, image_hwc=False):
        """BASELINE configs[0] on the device: everything val_zju.py does between the data loader and ScaleMapLearner.forward
        (:130-196) for frames whose RC-Net quasi-dense depth ``rcnet_depth`` (B,H,W) is already known (the reference loads it
        from the PNG run_rcnet_zju.py wrote).  Returns the front-end dict + ``int_scales``, ``sml_x``, ``sml_d`` (+ ``sml_pred``
        with an attached ``sml_model``)."""
        fe = self.align_and_interpolate(mono_pred, sparse_depth, max_knots)
        fe["int_scales"], _, fe["scaffold_status"] = ops.scale_scaffold(fe["int_depth"], rcnet_depth, fe["sparse_inv"], fe["valid"],
                                                                       self.min_depth, self.max_depth, png_quantize=True)
        fe["sml_x"], fe["sml_d"] = ops.sml_input_assemble(fe["int_depth"], fe["int_scales"], image, image_hwc=image_hwc)
        if self.sml_model is not None:
            fe["sml_pred"] = self._run_sml(fe["sml_x"], fe["sml_d"], fe["int_depth"].shape[-2:])
        return fe

    def _run_sml(self, x, d, size):
        # val_zju.py:198-210: ScaleMapLearner.forward(x, d) -> 1 / pred -> bicubic resize to the frame
        pred = self.sml_model(x, d)
        if pred.dim() == 3:
            pred = pred[:, None]
        return torch.nn.functional.interpolate(1.0 / pred, size=tuple(size), mode="bicubic", align_corners=False)[:, 0]

    # ------------------------------------------------------------------------------ device
    @torch.no_grad()
    def run_device(self, images, radar_points, frame_offsets, mono_pred, sparse_depth, sml_image=None):
        """images (B,3,H,W) in [0,1]; radar_points (R,3) pad-shifted (x+pw/2, y+ph/2, z), frames
        concatenated (with ``raw_inputs``: [0,255] frames and unshifted points); frame_offsets python list (B+1);
        mono_pred (B,H,W) relative inverse depth; sparse_depth (B,H,W) metric radar depth, 0 where empty;
        sml_image: the frame the SML network sees ((B,3,H,W) or (B,H,W,3) in [0,1]; default: the normalised ``images``).
        Everything on the device."""
        # The alignment / interpolation half of the front-end depends only on the mono prior and the radar map, not on
        # RC-Net: it runs on a side stream underneath the RC-Net forward (the triangulation is one CTA per frame and
        # would otherwise leave most of the 148 SMs idle for ~0.6 ms).  The scale scaffold joins the two.
        main = torch.cuda.current_stream()
        if self.overlap_front_end:
            if self._side is None:
                self._side = torch.cuda.Stream(device=images.device)
            fork = torch.cuda.Event()
            fork.record(main)
            self._side.wait_event(fork)
            ctx = torch.cuda.stream(self._side)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            counts = [frame_offsets[i + 1] - frame_offsets[i] for i in range(len(frame_offsets) - 1)]
            fe = self.align_and_interpolate(mono_pred, sparse_depth, self.max_knots or max(max(counts, default=1), 1))
            valid, sparse_inv, scale_shift, int_depth = fe["valid"], fe["sparse_inv"], fe["scale_shift"], fe["int_depth"]
            cnt, scale_map, status = fe["knot_count"], fe["scale_map"], fe["interp_status"]
            if self.overlap_front_end:
                join = torch.cuda.Event()
                join.record(self._side)
                for t in (valid, sparse_inv, scale_shift, int_depth, cnt, scale_map, status):
                    t.record_stream(main)
        fo = forward_output_batched(self.model, images, radar_points, frame_offsets,
                                    self.response_thr, chunk_points=self.chunk_points,
                                    raw_inputs=self.raw_inputs, retry_empty=self.retry_empty,
                                    return_select=self.return_select)
        depth_q, response = fo[0], fo[1]
        if self.overlap_front_end:
            main.wait_event(join)
        # val_zju.py:171-179 (interp='rcnet_*'): the scaffold the scripts actually feed to the SML network, with the
        # 16-bit PNG hand-off between run_rcnet_zju.py and val_zju.py applied in memory
        int_scales, _, scaffold_status = ops.scale_scaffold(int_depth, depth_q, sparse_inv, valid,
                                                            self.min_depth, self.max_depth, png_quantize=True)
        out = {"depth": depth_q, "response": response, "int_depth": int_depth, "scale_shift": scale_shift,
               "scale_map": scale_map, "interp_status": status, "knot_count": cnt,
               "int_scales": int_scales, "scaffold_status": scaffold_status}
        if self.return_select:
            out["select"] = fo[2]
        if self.sml_input:
            # val_zju.py:181-196: what ScaleMapLearner.forward(x, d) is fed
            img = sml_image if sml_image is not None else (images / 255.0 if self.raw_inputs else images)
            hwc = img.dim() == 4 and img.shape[-1] == 3 and img.shape[1] != 3
            out["sml_x"], out["sml_d"] = ops.sml_input_assemble(int_depth, int_scales, img, image_hwc=hwc)
            if self.sml_model is not None:
                out["sml_pred"] = self._run_sml(out["sml_x"], out["sml_d"], int_depth.shape[-2:])
        return out

    def _host_keys(self):
        if self.host_keys is not None:
            return self.host_keys
        keys = ("depth", "int_depth", "scale_map", "scale_shift", "interp_status")
        return keys + (("sml_pred",) if self.sml_model is not None else ())

    # ------------------------------------------------------------------------------ host
    def _pin(self, name, shape, dtype):
        t = self._pinned.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            self._pinned[name] = t
        return t

    @torch.no_grad()
    def run_host(self, images, radar_points, frame_offsets, mono_pred, sparse_depth, device=None):
        """Same chain with HOST inputs (numpy arrays or CPU tensors, ideally pinned) and pinned
        host outputs.  Returns (outputs dict of pinned CPU tensors, h2d_bytes, d2h_bytes).  The
        copies run on the current stream; the final ``synchronize`` is the caller's read point."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        h2d = 0
        dev = []
        for a in (images, radar_points, mono_pred, sparse_depth):
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            if not t.is_pinned():
                t = t.pin_memory()
            h2d += t.numel() * t.element_size()
            dev.append(t.to(device, non_blocking=True))
        out = self.run_device(dev[0], dev[1], list(frame_offsets), dev[2], dev[3])
        host, d2h = {}, 0
        for k in self._host_keys():
            h = self._pin(k, out[k].shape, out[k].dtype)
            h.copy_(out[k], non_blocking=True)
            d2h += h.numel() * h.element_size()
            host[k] = h
        return host, h2d, d2h

    @torch.no_grad()
    def submit_host(self, images, radar_points, frame_offsets, mono_pred, sparse_depth, device=None):
        """Pipelined form of ``run_host``: the host->device copies run on a copy stream, the chain on the current stream, the
        device->host copies on a second copy stream, each through one of two sets of persistent device / pinned buffers, so
        that the copies of step k + 1 / step k - 1 overlap the compute of step k.  Returns a handle; ``handle.wait()`` blocks
        until this step's results are in pinned host memory and returns ``(outputs, h2d_bytes, d2h_bytes)`` like
        ``run_host``.  The host inputs must stay untouched until then; at most two steps may be in flight (the buffers
        alternate).  No allocator interplay: every cross-stream buffer is persistent and guarded by events."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        main = torch.cuda.current_stream(device)
        if self._copy_streams is None:
            self._copy_streams = (torch.cuda.Stream(device), torch.cuda.Stream(device))
            self._slots = [dict(inp={}, stage={}, computed=None, drained=None) for _ in range(2)]
        s_in, s_out = self._copy_streams
        self._slot ^= 1
        slot = self._slots[self._slot]
        # ---- host -> device into this slot's persistent input buffers (free once the step that last used them has run)
        h2d, dev = 0, []
        if slot["computed"] is not None:
            s_in.wait_event(slot["computed"])
        with torch.cuda.stream(s_in):
            for name, a in (("images", images), ("points", radar_points), ("mono", mono_pred), ("sparse", sparse_depth)):
                t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
                if not t.is_pinned():
                    t = t.pin_memory()
                d = slot["inp"].get(name)
                if d is None or d.shape != t.shape or d.dtype != t.dtype:
                    d = torch.empty(t.shape, dtype=t.dtype, device=device)
                    slot["inp"][name] = d
                d.copy_(t, non_blocking=True)
                h2d += t.numel() * t.element_size()
                dev.append(d)
            loaded = torch.cuda.Event()
            loaded.record(s_in)
        main.wait_event(loaded)
        out = self.run_device(dev[0], dev[1], list(frame_offsets), dev[2], dev[3])
        # ---- results -> this slot's device staging buffers (on the compute stream), then device -> host on the copy stream
        keys = self._host_keys()
        if slot["drained"] is not None:
            main.wait_event(slot["drained"])                     # the copy-out of the step that last used the staging buffers
        for k in keys:
            g = slot["stage"].get(k)
            if g is None or g.shape != out[k].shape or g.dtype != out[k].dtype:
                g = torch.empty_like(out[k], memory_format=torch.contiguous_format)
                slot["stage"][k] = g
            g.copy_(out[k])
        slot["computed"] = torch.cuda.Event()
        slot["computed"].record(main)
        s_out.wait_event(slot["computed"])
        host, d2h = {}, 0
        with torch.cuda.stream(s_out):
            for k in keys:
                h = self._pin("%s#%d" % (k, self._slot), slot["stage"][k].shape, slot["stage"][k].dtype)
                h.copy_(slot["stage"][k], non_blocking=True)
                d2h += h.numel() * h.element_size()
                host[k] = h
            slot["drained"] = torch.cuda.Event()
            slot["drained"].record(s_out)
        return _HostResult(host, h2d, d2h, slot["drained"])


class _HostResult(object):
    def __init__(self, host, h2d, d2h, done):
        self.host, self.h2d, self.d2h, self.done = host, h2d, d2h, done

    def wait(self):
        self.done.synchronize()
        return self.host, self.h2d, self.d2h
