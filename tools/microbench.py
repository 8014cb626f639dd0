"""Per-kernel timings at BASELINE sizes (CUDA events, L2 flushed between iterations).
Usage: python tools/microbench.py [reduce lsq gather loftr interp]   -> JSON lines on stdout."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riders_b200 import ops, synth   # noqa: E402

DEV = torch.device("cuda:0")
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(PEAKS.get("hbm_gbs", 6650.0))
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    _flush.fill_(1)


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def report(name, ms, best, nbytes=None, flops=None, **kw):
    rec = {"kernel": name, "ms_median": round(ms, 4), "ms_min": round(best, 4)}
    if nbytes:
        rec.update(GBps=round(nbytes / ms / 1e6, 1), frac_hbm=round(nbytes / ms / 1e6 / HBM, 3), bytes=int(nbytes))
    if flops:
        rec.update(TFLOPs=round(flops / ms / 1e9, 2), flops=int(flops))
    rec.update(kw)
    print(json.dumps(rec), flush=True)


def bench_reduce(B=16, K=512, H=512, W=640, patch=(240, 100)):
    g = torch.Generator(device=DEV).manual_seed(0)
    conf = torch.sigmoid(torch.randn((B * K,) + patch, device=DEV, generator=g) * 2 - 1.5)
    pts = torch.cat([torch.from_numpy(synth.radar_points(f, K, H, W, patch)[0]) for f in range(B)]).to(DEV)
    offs = torch.arange(0, B * K + 1, K, dtype=torch.int32, device=DEV)
    nbytes = B * (K * patch[0] * patch[1] * 4 + 12 * K + 3 * H * W * 4)
    from riders_b200 import _lib
    for rows, tw in ((4, 32), (8, 32), (4, 16)):
        _lib.lib().riders_debug_set(b"reduce_rows", rows)
        _lib.lib().riders_debug_set(b"reduce_tile_w", tw)
        for logits in (False, True):
            for sel in (True, False):
                nb = B * (K * patch[0] * patch[1] * 4 + 12 * K + (3 if sel else 2) * H * W * 4)
                ms, best = timeit(lambda: ops.confidence_reduce(conf, pts, offs, H, W, 0.1, logits=logits, want_select=sel))
                report("confidence_reduce" + ("_logits" if logits else "") + "_rows%d_tw%d_%dmaps" % (rows, tw, 3 if sel else 2), ms, best, nb, B=B, K=K)
    _lib.lib().riders_debug_set(b"reduce_rows", 4)
    _lib.lib().riders_debug_set(b"reduce_tile_w", 16)


def bench_lsq(B=64, H=512, W=640, K=512):
    data = [synth.alignment_inputs(s, K, H, W) for s in range(4)]
    p = torch.from_numpy(np.stack([d[0] for d in data])).to(DEV).repeat(B // 4, 1, 1)
    t = torch.from_numpy(np.stack([d[1] for d in data])).to(DEV).repeat(B // 4, 1, 1)
    m = torch.from_numpy(np.stack([d[2] for d in data])).to(DEV).repeat(B // 4, 1, 1)
    nbytes = B * H * W * 13
    ms, best = timeit(lambda: ops.lsq_align(p, t, m, 10.0, 1 / 255.0))
    report("lsq_align", ms, best, nbytes, B=B)
    dense = torch.ones_like(m)
    ms, best = timeit(lambda: ops.lsq_align(p, t, dense, 10.0, 1 / 255.0))
    report("lsq_align_dense_mask", ms, best, nbytes, B=B)


def bench_gather(B=2, K=512, patch=(240, 100)):
    shapes = [(32, 376, 370), (64, 188, 185), (128, 94, 93), (128, 47, 47), (128, 24, 24)]
    scales = [1 / 2., 1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    sizes = [(120, 50), (60, 25), (30, 12), (15, 6), (7, 3)]
    g = torch.Generator(device=DEV).manual_seed(0)
    feats = [torch.randn((B,) + s, device=DEV, generator=g) for s in shapes]
    boxes = [torch.from_numpy(synth.radar_points(f, K, 512, 640, patch)[1]).to(DEV) for f in range(B)]
    rois = ops.boxes_to_rois(boxes, DEV)
    per_point = sum(c * oh * ow for (c, _, _), (oh, ow) in zip(shapes, sizes)) * 4
    nbytes = B * K * (per_point + 20) + B * sum(c * h * w for c, h, w in shapes) * 4
    for name, fs, mr in (("nchw_generic", feats, None),
                         ("nhwc_generic", [f.contiguous(memory_format=torch.channels_last) for f in feats], None),
                         ("nhwc_tma", [f.contiguous(memory_format=torch.channels_last) for f in feats], (100, 240))):
        ms, best = timeit(lambda: ops.roi_gather(fs, rois, scales, sizes, max_roi=mr), iters=5)
        report("roi_gather_" + name, ms, best, nbytes, B=B, K=K)
    # staging budget sweep of the TMA path (riders_debug_set knob): smaller strips -> more CTAs per SM
    from riders_b200 import _lib
    fs = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    for kb in (24, 36, 48, 56, 72, 100):
        _lib.lib().riders_debug_set(b"gather_smem_kb", kb)
        ms, best = timeit(lambda: ops.roi_gather(fs, rois, scales, sizes, max_roi=(100, 240)), iters=5)
        report("roi_gather_nhwc_tma_smem%dk" % kb, ms, best, nbytes, B=B, K=K)
    _lib.lib().riders_debug_set(b"gather_smem_kb", 48)


def bench_glue(R=2048):
    """Decoder glue kernels at RC-Net's largest shapes (per chunk of 2048 points)."""
    x = torch.randn((R, 32, 120, 50), device=DEV).contiguous(memory_format=torch.channels_last)
    ms, best = timeit(lambda: ops.upsample_nearest(x, (240, 100)), iters=5)
    report("upsample_nearest_32ch_240x100", ms, best, x.numel() * 4 * 5, R=R)
    y4 = torch.randn((R, 64, 120, 50), device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(16, device=DEV)
    ms, best = timeit(lambda: ops.d2s_bias_act_concat(y4, bias, None, 0.2), iters=5)
    report("d2s_bias_act_16ch_240x100", ms, best, y4.numel() * 4 * 2, R=R)
    z = torch.randn((R, 16, 240, 100), device=DEV).contiguous(memory_format=torch.channels_last)
    w = torch.randn((1, 16, 3, 3), device=DEV) * 0.1
    ms, best = timeit(lambda: ops.conv3x3_head(z, bias, 0.2, w), iters=5)
    report("conv3x3_head_16ch_240x100", ms, best, z.numel() * 4 + R * 24000 * 4, R=R)
    ms, best = timeit(lambda: ops.bias_act_(z, bias, 0.2), iters=5)
    report("bias_act_16ch_240x100", ms, best, z.numel() * 8, R=R)


def bench_tail(R=2048):
    """Fused decoder tail (d2s + 16->16 conv + head) against the three passes it replaces (d2s kernel, cuDNN TF32, head kernel)."""
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    g = torch.Generator(device=DEV).manual_seed(0)
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.1, torch.randn(16, device=DEV, generator=g) * 0.1
    wb = (torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.1).contiguous(memory_format=torch.channels_last)
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.1
    flops = R * 24000 * 16 * 16 * 9 * 2
    for taps in (3, 2):
        t = 1 if taps == 2 else 0
        y4 = torch.randn((R, 64, 120 + t, 50 + t), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
        nbytes = (y4.numel() + R * 24000) * 4
        ms, best = timeit(lambda: ops.decoder_tail(y4, ba, wb, bb, wh, 0.2, taps=taps, sigmoid=True), iters=5)
        report("decoder_tail_taps%d" % taps, ms, best, nbytes, flops=flops, R=R)
        ms, best = timeit(lambda: ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=taps, sigmoid=True, defer=True), iters=5)
        ops.LoftrOverflow.pending.clear()
        report("decoder_tail_tcgen05_taps%d" % taps, ms, best, nbytes, flops=flops, R=R)

        if taps == 2:
            x32 = torch.randn((R, 32, 120, 50), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
            w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
            ms, best = timeit(lambda: F.conv2d(x32, w2, None, 1, 1), iters=5)
            report("phase_conv_cudnn_tf32", ms, best, x32.numel() * 4 + y4.numel() * 4, R=R)
            ms, best = timeit(lambda: ops.subpixel_conv_tc_half(x32, w2, ba, 0.2), iters=5)
            report("phase_conv_tcgen05_fp16_map", ms, best, x32.numel() * 4 + R * 24000 * 32, R=R)
            yh, _ = ops.subpixel_conv_tc_half(x32, w2, ba, 0.2)
            ms, best = timeit(lambda: ops.decoder_tail_tc_half(yh, wb, bb, wh, 0.2, sigmoid=True), iters=5)
            report("decoder_tail_tcgen05_fp16_map", ms, best, R * 24000 * 36, flops=flops, R=R)
            ms, best = timeit(lambda: ops.decoder_upconv_tail_tc(x32, w2, ba, wb, bb, wh, 0.2, sigmoid=True, defer=True), iters=5)
            ops.LoftrOverflow.pending.clear()
            report("upconv_tail_tcgen05_fp16_map (both kernels)", ms, best, x32.numel() * 4 + R * 24000 * 4, flops=flops, R=R)
            ms, best = timeit(lambda: ops.decoder_tail_tc(F.conv2d(x32, w2, None, 1, 1), ba, wb, bb, wh, 0.2, taps=2, sigmoid=True, defer=True), iters=5)
            ops.LoftrOverflow.pending.clear()
            report("cudnn_phase_conv + decoder_tail_tcgen05", ms, best, x32.numel() * 4 + R * 24000 * 4, flops=flops, R=R)

        def separate():
            z = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=taps)
            return ops.conv3x3_head(F.conv2d(z, wb, None, 1, 1), bb, 0.2, wh, sigmoid=True)
        ms, best = timeit(separate, iters=5)
        report("separate_d2s_cudnn_head_taps%d" % taps, ms, best, nbytes, flops=flops, R=R)


def bench_subconv(R=2048):
    """Own TF32 phase convolution (bias + LeakyReLU of the producer fused in) against bias_act_ + cuDNN's 2x2 convolution."""
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn((R, 32, 120, 50), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    b = torch.randn(32, device=DEV, generator=g) * 0.1
    w2 = (torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1).contiguous(memory_format=torch.channels_last)
    flops = R * 121 * 51 * 128 * 64 * 2
    nbytes = (x.numel() + R * 64 * 121 * 51) * 4
    ms, best = timeit(lambda: ops.subpixel_conv_tc(x, w2), iters=5)
    report("subpixel_conv_tc (conv only)", ms, best, nbytes, flops=flops, R=R)
    ms, best = timeit(lambda: F.conv2d(x, w2, None, 1, 1), iters=5)
    report("cudnn_conv2x2 (conv only)", ms, best, nbytes, flops=flops, R=R)
    ms, best = timeit(lambda: ops.subpixel_conv_act(x, b, 0.2, w2), iters=5)
    report("subpixel_conv_act_mma (bias+act fused)", ms, best, nbytes, flops=flops, R=R)
    xc = x.clone(memory_format=torch.channels_last)
    ms, best = timeit(lambda: F.conv2d(ops.bias_act_(xc, b, 0.2), w2, None, 1, 1), iters=5)
    report("bias_act_plus_cudnn_conv2x2", ms, best, nbytes + 2 * x.numel() * 4, flops=flops, R=R)


def bench_mmasync():
    """Issue rate / latency of mma.sync.m16n8k8.tf32 (registers only): chains = independent accumulators per warp; mode 1 gives
    every MMA its own operand registers (a real kernel), mode 0 re-uses one A / B fragment."""
    from riders_b200 import _lib
    cyc = torch.zeros(148, dtype=torch.int64, device=DEV)
    sink = torch.zeros(4, device=DEV)
    iters = 4096
    for mode in (0, 1):
        for chains in (1, 4, 8):
            for warps in (4, 8, 16):
                _lib.check(_lib.lib().riders_debug_mma_sync_rate(warps | (chains << 8) | (mode << 12), iters, 148, cyc.data_ptr(), sink.data_ptr(), None), "mma_sync_rate")
                torch.cuda.synchronize()
                c = float(cyc.double().mean())
                per_sm = c / (iters * 8 * warps)
                print(json.dumps({"kernel": "mma_sync_m16n8k8_tf32", "distinct_operands": mode, "chains_per_warp": chains, "warps": warps,
                                  "cycles_per_mma_per_sm": round(per_sm, 3), "cycles_per_mma_per_warp": round(c / (iters * 8), 2),
                                  "mac_per_cycle_per_sm": round(1024 / per_sm, 1)}), flush=True)


def bench_loftr(N=4096):
    import oracle  # noqa: F401  (only to keep the import graph explicit: weights are random here)
    from riders_b200.rcnet.linear_attention import LocalFeatureTransformer
    torch.manual_seed(0)
    tf = LocalFeatureTransformer(['self', 'cross'], n_layers=4, d_model=128).to(DEV).eval()
    f0 = torch.randn(N, 21, 128, device=DEV)
    f1 = torch.relu(torch.randn(N, 21, 128, device=DEV))
    flops = N * 112939008
    params = [l.kernel_params() for l in tf.layers]
    for impl, name in ((ops.LOFTR_SIMT, "simt"), (ops.LOFTR_TCGEN05, "tcgen05")):
        try:
            ms, best = timeit(lambda: ops.loftr_transformer(f0, f1, params, tf.layer_names, 8, impl), iters=5)
            report("loftr_transformer_" + name, ms, best, flops=flops, N=N)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"kernel": "loftr_transformer_" + name, "error": str(e)[:200]}), flush=True)


def bench_interp(B=64, K=512, H=512, W=640):
    xy = np.zeros((B, K, 2), np.int32)
    for b in range(B):
        flat = np.sort(np.random.default_rng(b).choice(H * W, K, replace=False))
        xy[b] = np.stack([flat % W, flat // W], 1)
    val = torch.rand(B, K, device=DEV) + 0.5
    xyd = torch.from_numpy(xy).to(DEV)
    cnt = torch.full((B,), K, dtype=torch.int32, device=DEV)
    ms, best = timeit(lambda: ops.interp_linear(xyd, val, cnt, H, W, 1.0), iters=5)
    report("interp_linear(delaunay+fill)", ms, best, B * (H * W * 4 + 12 * K), B=B, K=K)
    data = [synth.alignment_inputs(s, K, H, W) for s in range(4)]
    p = torch.from_numpy(np.stack([d[0] for d in data])).to(DEV).repeat(B // 4, 1, 1)
    t = torch.from_numpy(np.stack([d[1] for d in data])).to(DEV).repeat(B // 4, 1, 1)
    m = torch.from_numpy(np.stack([d[2] for d in data])).to(DEV).repeat(B // 4, 1, 1)
    ms, best = timeit(lambda: ops.extract_knots(p, t, m, K), iters=5)
    report("extract_knots", ms, best, B * H * W * 9, B=B, K=K)


def bench_frontend(H=512, W=640, K=512):
    """The SML front-end kernels ALONE on an otherwise idle GPU (in the bench step they share the device with RC-Net), at the
    batch sizes of configs[0] (1 frame), configs[1] (16) and 64: the three launch forms of the LSQ alignment, the L1 Optimizer,
    knot extraction, and the two launches of the interpolation separately."""
    import ctypes as C
    from riders_b200 import _lib
    L = _lib.lib()
    for B in (1, 16, 64):
        data = [synth.alignment_inputs(s, K, H, W) for s in range(4)]
        rep = (B + 3) // 4
        p = torch.from_numpy(np.stack([d[0] for d in data])).to(DEV).repeat(rep, 1, 1)[:B].contiguous()
        t = torch.from_numpy(np.stack([d[1] for d in data])).to(DEV).repeat(rep, 1, 1)[:B].contiguous()
        m = torch.from_numpy(np.stack([d[2] for d in data])).to(DEV).repeat(rep, 1, 1)[:B].contiguous()
        nb = B * H * W * 13
        for mode, name in ((0, "two_kernels"), (1, "cooperative_one_launch"), (2, "pdl"), (3, "single_launch_flag_barrier")):
            ops.debug_set("lsq_fused", mode)
            ms, best = timeit(lambda: ops.lsq_align(p, t, m, 10.0, 1 / 255.0), iters=20)
            report("lsq_align_" + name, ms, best, nb, B=B)
        ops.debug_set("lsq_fused", 2)
        ms, best = timeit(lambda: ops.l1_scale_align(p, t, m, clamp_hi=10.0, clamp_lo=1 / 255.0), iters=20)
        report("l1_scale_align", ms, best, nb, B=B)
        _, _, aligned = ops.lsq_align(p, t, m, 10.0, 1 / 255.0)
        ms, best = timeit(lambda: ops.extract_knots(aligned, t, m, K), iters=20)
        report("extract_knots", ms, best, B * (H * W + 20 * K), B=B, K=K)
        xy, sc, _, cnt = ops.extract_knots(aligned, t, m, K)
        ws = torch.empty(L.riders_interp_workspace_bytes(B, K, H, W), dtype=torch.uint8, device=DEV)
        out = torch.empty((B, H, W), device=DEV)
        status = torch.zeros(B, dtype=torch.int32, device=DEV)
        vp = lambda x: C.c_void_p(x.data_ptr())
        st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ms, best = timeit(lambda: L.riders_interp_triangulate(vp(xy), vp(sc), vp(cnt), B, K, vp(status), None, None, vp(ws), st()), iters=20)
        report("interp_stars (delaunay_star_kernel, no fill value)", ms, best, B * 12 * K, B=B, K=K)
        ms, best = timeit(lambda: L.riders_interp_fill(B, K, H, W, 1.0, vp(out), vp(ws), st()), iters=20)
        report("interp_fill (fill value + raster_kernel)", ms, best, B * H * W * 4, B=B, K=K)
        ms, best = timeit(lambda: ops.interp_linear(xy, sc, cnt, H, W, 1.0), iters=20)
        report("interp_linear (stars + fill value, raster)", ms, best, B * (H * W * 4 + 12 * K), B=B, K=K)
        assert int(status.abs().sum()) == 0
        depth = torch.where(m, 1.0 / t.clamp_min(1e-6), torch.zeros_like(t))
        ms, best = timeit(lambda: ops.scale_scaffold(aligned, depth, t, m), iters=20)
        report("scale_scaffold", ms, best, B * H * W * 25, B=B)


def bench_membw(nbytes=4 << 30):
    """Direction-specific HBM rooflines on this box (library kernels, for reference next to MEASURED_PEAKS' copy figure):
    write-only (fill_), read-only (sum), copy."""
    a = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
    b = torch.empty_like(a)
    ms, best = timeit(lambda: a.fill_(1.0), iters=8)
    report("hbm_write_only_fill", ms, best, nbytes)
    ms, best = timeit(lambda: a.zero_(), iters=8)
    report("hbm_write_only_memset", ms, best, nbytes)
    ms, best = timeit(lambda: a.sum(), iters=8)
    report("hbm_read_only_sum", ms, best, nbytes)
    ms, best = timeit(lambda: b.copy_(a), iters=8)
    report("hbm_copy", ms, best, 2 * nbytes)


if __name__ == "__main__":
    which = sys.argv[1:] or ["reduce", "lsq", "gather", "loftr", "interp", "glue"]
    torch.cuda.set_device(DEV)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": HBM}), flush=True)
    with torch.no_grad():
        for w in which:
            globals()["bench_" + w]()
