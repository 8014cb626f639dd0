"""``forward_output`` (row A5 of SURVEY 8(a)), host side: mirror of
``RCNet/rcnet_main.py:435-487`` of the reference with the same signature and return values.

Reference flow: edge-pad -> model.forward(sigmoid) -> K zero canvases -> cat -> max/argmax ->
confidence-weighted mean.  Here: edge-pad -> model.forward(LOGITS) -> one launch of the
reduce kernel, which applies the sigmoid + threshold on the fly and never materialises the
(K, H', W') tensor.  ``forward_output_batched`` is the B200 entry point for many frames.
"""
import torch

from .. import ops
from . import networks


def _pad_edge(image, px, py):
    # torchvision.transforms.functional.pad(..., padding_mode='edge') == replicate padding
    return torch.nn.functional.pad(image, (px, px, py, py), mode='replicate')


def _model_logits(model, image, radar_points, bounding_boxes_list, chunk_points):
    """model.forward(return_logits=True), optionally in chunks of points to bound the decoder's
    activation memory (the gathered skips are 1.39 MB per point)."""
    n = radar_points.shape[0]
    if chunk_points is None or n <= chunk_points or len(bounding_boxes_list) != 1:
        return model.forward(image=image, point=radar_points, bounding_boxes=bounding_boxes_list,
                             return_logits=True)
    outs = []
    boxes = bounding_boxes_list[0]
    for s in range(0, n, chunk_points):
        e = min(n, s + chunk_points)
        outs.append(model.forward(image=image, point=radar_points[s:e], bounding_boxes=[boxes[s:e]],
                                  return_logits=True))
    return torch.cat(outs, dim=0)


def forward_output(model, image, radar_points, bounding_boxes_list, response_thr=0.5,
                   device=torch.device('cuda'), return_select=False):
    """image (1,3,H,W); radar_points (K,3) or (1,K,3), pad-shifted; bounding_boxes_list [Tensor(K,4)].
    Returns (output_depth (1,H,W), output_response (1,H,W)) like the reference (+ select if asked)."""
    ph, pw = model.input_patch_size_image
    px, py = pw // 2, ph // 2
    H, W = image.shape[-2], image.shape[-1]
    image = _pad_edge(image, px, py)
    if radar_points.dim() == 3:
        radar_points = torch.squeeze(radar_points, dim=0)
    logits = model.forward(image=image, point=radar_points, bounding_boxes=bounding_boxes_list,
                           return_logits=True)
    depth, response, select = ops.confidence_reduce(
        logits, radar_points, [0, radar_points.shape[0]], H, W, response_thr, logits=True,
        want_select=return_select)
    if return_select:
        return depth, response, select
    return depth, response


def _latent_to_maps(model, enc, dec, latent_image, skips_image, radar_points, rois, frame_offsets, H, W, ph, pw, response_thr,
                    chunk_points, return_select):
    """Latent path (1/32-level gather + point MLP + LoFTR) for every point of the batch at once, then the skip gather + decoder per
    chunk of points and the reduce kernel.  Returns (reduce outputs, logits / confidences, is_logits)."""
    images_device = radar_points.device
    R = radar_points.shape[0]
    logits = torch.empty((R, 1, ph, pw), device=images_device, dtype=torch.float32)
    # latent path (1/32-level gather + point MLP + LoFTR) for every point of the batch at once: 1366 tiles of
    # 6 points fill the 148 SMs in ~9 full waves; the skip gather + decoder run per chunk of points.
    latent = enc.encode_latent(latent_image, radar_points, rois)
    # each pooled skip is gathered straight into the concat buffer of the decoder block that consumes it
    lead = dec.skip_lead_channels(len(skips_image)) if hasattr(dec, 'skip_lead_channels') else None
    sigmoid_in_head = None if isinstance(dec, networks.MultiScaleDecoder) else False
    for s in range(0, R, chunk_points):
        e = min(R, s + chunk_points)
        skips = enc.gather_skips(skips_image, rois[s:e], lead_channels=lead)
        # the fused head kernel writes sigmoid(logit) directly when it runs (SFU work under an HBM-bound kernel), so the
        # reduce streams confidences; otherwise the reduce applies the sigmoid itself
        if sigmoid_in_head is None or sigmoid_in_head:
            logits[s:e] = dec(x=latent[s:e], skips=skips, shape=(ph, pw), head_sigmoid=True)[-1]
            done = bool(getattr(dec, 'last_head_sigmoid', False))
            if sigmoid_in_head is None:
                sigmoid_in_head = done
            elif not done:
                logits[s:e] = torch.sigmoid(logits[s:e])
        else:
            logits[s:e] = dec(x=latent[s:e], skips=skips, shape=(ph, pw))[-1]
    is_logits = not sigmoid_in_head
    out = list(ops.confidence_reduce(logits, radar_points, frame_offsets, H, W, response_thr, logits=is_logits,
                                     want_select=return_select))
    return out, logits, is_logits


def forward_output_batched(model, images, radar_points, frame_offsets, response_thr=0.5,
                           chunk_points=1024, return_select=False, raw_inputs=False, retry_empty=False):
    """Many frames per call (BASELINE config 2/3).

    images (B,3,H,W) unpadded; radar_points (R,3), frames concatenated; frame_offsets: python list (B+1).
    ``raw_inputs=False``: images already normalised to [0,1] and points pad-shifted (what ``forward_output`` takes).
    ``raw_inputs=True``: images as loaded ([0,255]) and points in frame coordinates - the prologue of
    run_rcnet_zju.py:214-240 (shift, boxes, int()/255 normalisation, edge pad) runs as one kernel.
    ``retry_empty``: the caller's retry loop of run_rcnet_zju.py:254-265 - a frame whose depth map is all zero is reduced
    again with the threshold lowered by 0.05 until something fires (one tiny D2H per call; the logits are reused).
    The backbone runs once on the padded batch, the encoder / decoder run per chunk of points (each chunk within one
    frame), the reduce kernel once.  Returns depth (B,H,W), response (B,H,W) [, select]."""
    ph, pw = model.input_patch_size_image
    px, py = pw // 2, ph // 2
    B, _, H, W = images.shape
    enc = model.encoder.module if hasattr(model.encoder, 'module') else model.encoder
    dec = model.decoder.module if hasattr(model.decoder, 'module') else model.decoder
    if raw_inputs:
        cl = next(enc.encoder_image.parameters()).is_contiguous(memory_format=torch.channels_last) and \
            next(enc.encoder_image.parameters()).dim() == 4 and not next(enc.encoder_image.parameters()).is_contiguous()
        padded, radar_points, rois = ops.rcnet_prologue(images, radar_points, frame_offsets, (ph, pw), channels_last=cl)
    else:
        padded = _pad_edge(images, px, py)
        x, y = radar_points[:, 0:1], radar_points[:, 1:2]
        boxes = torch.cat([x - px, y - py, x + px, y + py], dim=1)      # run_rcnet_zju.py:227-232
        frame_of = torch.repeat_interleave(
            torch.arange(B, device=images.device, dtype=torch.float32),
            torch.tensor([frame_offsets[i + 1] - frame_offsets[i] for i in range(B)], device=images.device))
        rois = torch.cat([frame_of[:, None], boxes], dim=1).contiguous()
    latent_image, skips_image = enc.encoder_image(padded)
    R = radar_points.shape[0]
    tf = getattr(enc, 'attention', None)
    for attempt in (0, 1):
        # The tcgen05 LoFTR path works in fp16 hi/lo pairs; a value beyond the fp16 range raises a device flag, which is read
        # at this function's own sync point (no extra stall in the common case).  If it was raised, the latent path and
        # everything after it are recomputed once with the fp32 SIMT kernels.
        if tf is not None:
            tf.defer_overflow_check = attempt == 0
            tf.force_impl = ops.LOFTR_SIMT if attempt == 1 else None
        if hasattr(dec, 'defer_tail_check'):             # the tcgen05 decoder tail has the same guard (fp16 operands)
            dec.defer_tail_check = attempt == 0
            dec.force_tail_impl = 'mma_sync' if attempt == 1 else None
        try:
            out, logits, is_logits = _latent_to_maps(model, enc, dec, latent_image, skips_image, radar_points, rois, frame_offsets,
                                                     H, W, ph, pw, response_thr, chunk_points, return_select)
        finally:
            if tf is not None:
                tf.defer_overflow_check = False
                tf.force_impl = None
            if hasattr(dec, 'defer_tail_check'):
                dec.defer_tail_check = False
                dec.force_tail_impl = None
        if attempt == 0 and ops.LoftrOverflow.pending and ops.loftr_overflow_pending():
            continue
        break
    if retry_empty:
        thr = float(response_thr)
        empty = (out[0].flatten(1).amax(dim=1) == 0).nonzero().flatten().tolist()          # D2H: the caller's np.sum(...) == 0
        while empty and thr > -1.0:
            thr -= 0.05
            for b in empty:
                s, e = frame_offsets[b], frame_offsets[b + 1]
                if e == s:
                    continue
                again = ops.confidence_reduce(logits[s:e], radar_points[s:e], [0, e - s], H, W, thr, logits=is_logits,
                                              want_select=return_select)
                for o, a in zip(out, again):
                    if o is not None:
                        o[b] = a[0]
            empty = [b for b in empty if frame_offsets[b + 1] > frame_offsets[b] and float(out[0][b].amax()) == 0.0]
    return tuple(out[:3 if return_select else 2])
