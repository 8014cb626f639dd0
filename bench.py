#!/usr/bin/env python
"""RIDERS radar-to-dense-depth hot path: frames/s on B200 (+ roofline of the dominant kernel,
CPU baseline, end-to-end figure).  Contract: see the task statement; one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1] + the SML front-end the metric names): per step and per GPU a
batch of 16 synthetic 640x512 thermal frames with 512 radar points each goes through
  RC-Net quasi-dense depth (backbone -> roi gather -> point MLP -> LoFTR -> decoder -> reduce, thr 0.1)
  -> global scale/shift LSQ alignment of the mono prior -> Delaunay/barycentric scale-map interpolation.
Frames are independent, so N GPUs = N processes each with its own 16 frames (weak scaling), no
data-path collective; the one all-reduce is the final metric reduction.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RC-Net+SML-align frames/s @640x512"
UNIT = "frames/s"
H, W, PATCH, K_POINTS, FRAMES_PER_GPU, THR = 512, 640, (240, 100), 512, 16, 0.1
MODEL_ARGS = dict(input_channels_image=3, input_channels_depth=3, input_patch_size_image=list(PATCH),
                  encoder_type=["rcnet", "batch_norm"], n_filters_encoder_image=[32, 64, 128, 128, 128],
                  n_neurons_encoder_depth=[32, 64, 128, 128, 128], decoder_type=["multiscale", "batch_norm"],
                  n_filters_decoder=[256, 128, 64, 32, 16])          # run_rcnet_zju.py:30-45

# algorithmic work per unit
GATHER_BYTES_PER_POINT = 1393152 + 20
GATHER_MAP_BYTES_PER_FRAME = 32612864
LOFTR_FLOP_PER_POINT = 112939008
REDUCE_BYTES_PER_POINT = 96000 + 12
MAP_BYTES = H * W * 4


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": float(p["hbm_gbs"]), "tensor": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json; bf16 sustained)"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def build_model(device, seed=0):
    from riders_b200.rcnet.rcnet_model import RCNetModel
    torch.manual_seed(seed)
    model = RCNetModel(device=torch.device("cpu"), **MODEL_ARGS)      # init on the CPU: same weights on every arm
    model.eval()
    if device.type == "cuda":
        model.device = device
        model.to(device)
        model.optimize_for_inference()
        torch.backends.cudnn.benchmark = os.environ.get("RIDERS_CUDNN_BENCHMARK", "1") == "1"   # autotune the torch convs
    return model


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows if r[2].replace(".", "").isdigit())}


def cpu_frame_rate(seconds_budget, model_cpu=None, max_frames=4, global_alignment="st"):
    """The oracle port of the reference path on the host cores: frames/s over a bounded sample.  Also returns the
    reference outputs of those frames (the first frames of rank 0's batch) for the parity block."""
    from oracle import chain_ref
    from riders_b200 import synth
    model_cpu = model_cpu or build_model(torch.device("cpu"))
    refs, t0 = [], time.perf_counter()
    while len(refs) < max_frames:
        refs.append(chain_ref.chain_frame_ref(model_cpu, synth.chain_frame(len(refs), K_POINTS, H, W, PATCH), THR,
                                              want_conf=True, global_alignment=global_alignment))
        if time.perf_counter() - t0 > seconds_budget:
            break
    dt = time.perf_counter() - t0
    return len(refs) / dt, len(refs), dt, refs


def parity_block(out, refs):
    """GPU outputs of the benched configuration (exactly the flags that were timed) against the oracle port on the same
    frames, at north_star's rules: confidences 1e-4 relative, select exact where the top-2 margin > 1e-6, scale / shift
    1e-6 relative, scale map 1e-5 relative inside the hull and exactly the fill value outside."""
    from oracle import confidence_reduce_ref as cr, interpolator_ref as ir
    from riders_b200 import synth
    res = {"frames": len(refs), "conf_max_rel_err": 0.0, "conf_fired_mismatch_frac": 0.0, "select_agree_frac_margin_gt_1e-6": 1.0,
           "select_margin_gt_1e-6_frac": 0.0, "depth_max_rel_err": 0.0, "scale_rel_err": 0.0, "shift_rel_err": 0.0,
           "int_depth_max_rel_err": 0.0, "scale_map_max_rel_err_inside": 0.0, "scale_map_outside_mismatch_px": 0,
           "scale_map_masked_degenerate_frac": 0.0}
    sel_ok, sel_n, px = 0, 0, 0
    for b, ref in enumerate(refs):
        r = out["response"][b].cpu().numpy()
        d = out["depth"][b].cpu().numpy()
        fired, fired_ref = r > 0, ref["response"] > 0
        both = fired & fired_ref
        res["conf_fired_mismatch_frac"] = max(res["conf_fired_mismatch_frac"], float((fired != fired_ref).mean()))
        if both.any():
            res["conf_max_rel_err"] = max(res["conf_max_rel_err"], float((np.abs(r[both] - ref["response"][both]) / ref["response"][both]).max()))
            res["depth_max_rel_err"] = max(res["depth_max_rel_err"], float((np.abs(d[both] - ref["depth"][both]) / ref["depth"][both]).max()))
        if "select" in out:
            fr = synth.chain_frame(b, K_POINTS, H, W, PATCH)
            margin = cr.top2_margin(ref["conf"], fr["points"], H, W, THR)
            m = (margin > 1e-6) & fired_ref
            sel_ok += int((out["select"][b].cpu().numpy()[m] == ref["select"][m]).sum())
            sel_n += int(m.sum())
            px += int(fired_ref.sum())
        ss = out["scale_shift"][b].cpu().numpy()
        res["scale_rel_err"] = max(res["scale_rel_err"], abs(ss[0] - ref["scale"]) / max(abs(ref["scale"]), 1e-30))
        if ref["shift"] != 0.0:
            res["shift_rel_err"] = max(res["shift_rel_err"], abs(ss[1] - ref["shift"]) / abs(ref["shift"]))
        idp = out["int_depth"][b].cpu().numpy()
        res["int_depth_max_rel_err"] = max(res["int_depth_max_rel_err"], float((np.abs(idp - ref["int_depth"]) / ref["int_depth"]).max()))
        coords, _, _ = ir.extract_knots_ref(ref["int_depth"], ref["sparse_inv"], ref["valid"])
        ok = ~ir.degenerate_pixel_mask(coords.T, H, W)
        sm, smr = out["scale_map"][b].cpu().numpy(), ref["scale_map"]
        res["scale_map_masked_degenerate_frac"] = max(res["scale_map_masked_degenerate_frac"], float(1.0 - ok.mean()))
        res["scale_map_outside_mismatch_px"] += int(((sm == 1.0) != (smr == 1.0))[ok].sum())
        res["scale_map_max_rel_err_inside"] = max(res["scale_map_max_rel_err_inside"], float((np.abs(sm - smr) / np.abs(smr))[ok].max()))
    if sel_n:
        res["select_agree_frac_margin_gt_1e-6"] = sel_ok / sel_n
        res["select_margin_gt_1e-6_frac"] = sel_n / max(px, 1)
    res["meets_north_star"] = bool(res["conf_max_rel_err"] <= 1e-4 and res["select_agree_frac_margin_gt_1e-6"] == 1.0 and
                                   res["scale_rel_err"] <= 1e-6 and res["shift_rel_err"] <= 1e-6 and
                                   res["scale_map_max_rel_err_inside"] <= 1e-5 and res["scale_map_outside_mismatch_px"] == 0)
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import chain_ref
    from riders_b200 import synth
    model = build_model(torch.device("cpu"))
    frames = [synth.chain_frame(i, K_POINTS, H, W, PATCH) for i in range(2)]
    for i in range(max(args.warmup, 0)):                      # bounded: warm-up on a K=64 slice of a frame
        chain_ref.chain_frame_ref(model, synth.chain_frame(100 + i, 64, H, W, PATCH), THR, global_alignment=args.align)
    steps = max(1, min(args.steps, 6))                           # one frame per step, <= ~1-2 min of CPU work
    t0 = time.perf_counter()
    for i in range(steps):
        chain_ref.chain_frame_ref(model, frames[i % len(frames)], THR, global_alignment=args.align)
    dt = time.perf_counter() - t0
    v = steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(1, frames_per_gpu=1, align=args.align,
                                      note="reference arm: one 640x512 frame (K=512) per step on the host cores"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d steps x 1 frame (K=512) of the same synthetic workload; oracle port of "
                                       "forward_output + LSQ align + griddata-linear (CPU torch/torchvision/NumPy/SciPy)" % steps},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def dtype_label(tf32):
    """The arithmetic the timed step actually ran in (label it truthfully)."""
    if tf32:
        tail = os.environ.get("RIDERS_TAIL_IMPL", "tcgen05")
        fp16map = tail == "tcgen05" and os.environ.get("RIDERS_TAIL_FP16_MAP", "1") != "0"
        return ("tf32 (cuDNN backbone/decoder convolutions, torch's and the reference's GPU default) + "
                + (("last two up-convolutions and fused decoder tail on tcgen05.mma with fp16 operands / fp32 accumulation (the same 11-bit "
                    "significand as TF32, range-guarded; fp16 map between the last up-convolution and the tail)"
                    if os.environ.get("RIDERS_TAIL_UPCONV", "1") != "0" else
                    "last up-convolution (TF32 tcgen05.mma) and fused decoder tail (tcgen05.mma, fp16 operands / fp32 accumulation: the "
                    "same 11-bit significand as TF32, range-guarded) with an fp16 map between them") if fp16map else
                   "fused decoder tail on tcgen05 with fp16 operands / fp32 accumulation (same 11-bit significand as TF32, range-guarded)"
                   if tail == "tcgen05" else "fused decoder tail on TF32 mma.sync") + " / "
                "f32 (gather, LoFTR [fp16 hi/lo split, fp32-accurate], reduce, prologue) / f64 (alignment sums, barycentric fill)")
    return "f32 (cuDNN fp32 convolutions, no TF32 anywhere) / f64 (alignment sums, barycentric fill)"


def pin_rank_cores(local, world):
    """N ranks share one host: give each rank its own slice of the cores so the Python launch threads and the pinned-memory
    copies of different ranks do not migrate over each other."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(len(cores) // world, 1)
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(per, 8)))
    except OSError:
        pass


def run_eager(args):
    """Same-box LIBRARY comparator: the reference's own formulation on the GPU with stock
    library kernels only - torchvision roi_pool x5 (CUDA), the three einsums + nn.functional linears of LinearAttention /
    LoFTREncoderLayer through cuBLAS, cuDNN convolutions in the reference's NCHW layout, and the literal (K, H', W') zero
    canvases + torch.max + weighted mean of rcnet_main.py:460-485 - followed by the NumPy / SciPy alignment and
    interpolation on the host, one frame per step like run_rcnet_zju.py.  None of this repo's kernels run here
    (``oracle/rcnet_ref`` with device='cuda'; the model is the mirror's plain torch modules in NCHW, which take no fast path)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "eager", "unavailable": "no CUDA device"}), flush=True)
        return 0
    from oracle import chain_ref, rcnet_ref
    from riders_b200 import synth
    dev = torch.device("cuda", 0)
    if args.strict_fp32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    from riders_b200.rcnet.rcnet_model import RCNetModel
    model = RCNetModel(device=torch.device("cpu"), **MODEL_ARGS)
    model.eval()
    model.device = dev
    model.to(dev)                                            # NCHW, BatchNorm modules in place: plain torch everywhere
    frames = [synth.chain_frame(i, K_POINTS, H, W, PATCH) for i in range(4)]

    def one(fr, rcnet_only=False):
        with torch.no_grad():
            d, r, _ = rcnet_ref.forward_output_ref(model, torch.from_numpy(fr["image"])[None].to(dev), torch.from_numpy(fr["points"]).to(dev),
                                                   [torch.from_numpy(fr["boxes"]).to(dev)], THR, literal=True)
            d = d.cpu().numpy()                              # run_rcnet_zju.py:250
        if not rcnet_only:
            chain_ref.sml_front_end_ref(fr["mono"], fr["sparse"], global_alignment=args.align)
        return d

    for i in range(max(args.warmup, 2)):
        one(frames[i % 4])
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 12))
    t0 = time.perf_counter()
    for i in range(steps):
        one(frames[i % 4])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    for i in range(steps):
        one(frames[i % 4], rcnet_only=True)
    torch.cuda.synchronize()
    dt_rc = time.perf_counter() - t1
    v = steps / dt
    print(json.dumps({"impl": "eager", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": dtype_label(bool(torch.backends.cudnn.allow_tf32)).split(" / ")[0] + " (library kernels)", "data": "synthetic",
                      "config": workload_config(1, frames_per_gpu=1, align=args.align, strict=args.strict_fp32,
                                                note="library comparator: reference formulation with stock torch / torchvision CUDA kernels, "
                                                     "one frame per step; alignment + griddata on the host as in the reference"),
                      "rcnet_only_frames_per_s": steps / dt_rc, "rcnet_only_ms_per_frame": 1e3 * dt_rc / steps}), flush=True)
    return 0


def run_config1(args):
    """BASELINE configs[0]: the val_zju.py SML front-end for ONE 640x512 frame with 200 radar points - global alignment,
    knot extraction, Delaunay, barycentric fill, scale scaffold, SML input assembly - as latency per frame on one B200.
    Device-resident inputs, (a) eager launches and (b) the same launches replayed from a CUDA graph (the inner loop is
    launch-bound); e2e = NumPy in / NumPy out through the reference-facing classes.  The SML network itself cannot be built
    offline (torch.hub encoder): attach any forward(x, d) module with RadarDepthPipeline(sml_model=...)."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --workload config1 needs a CUDA device")
    from riders_b200 import ops, synth
    from riders_b200.pipeline import RadarDepthPipeline
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    K1 = 200
    fr = synth.chain_frame(0, K1, H, W, PATCH)
    pipe = RadarDepthPipeline(None, global_alignment=args.align)
    t = {k: torch.from_numpy(fr[k])[None].to(dev) for k in ("mono", "sparse", "image")}
    rc_depth = t["sparse"].clone()                             # stand-in for the PNG run_rcnet_zju.py wrote: radar depths at the radar pixels

    def step():
        return pipe.sml_front_end(t["mono"], t["sparse"], rc_depth, t["image"], K1)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    steps = max(args.steps, 50)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2: flushed between timed iterations
    evs = []
    launches0 = ops.launch_count()
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = step(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    launches = (ops.launch_count() - launches0) / steps
    lat = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    with ops.StageTimer() as timer:
        for _ in range(steps):
            flush.zero_()
            step()
    stages = {k: round(1e3 * ms / calls, 2) for k, (calls, ms, _) in timer.summary().items()}
    # CUDA graph replay of the same launches
    graph_us = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            gout = step()
        g.replay()
        torch.cuda.synchronize()
        ge = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            ge.append((a, b))
        torch.cuda.synchronize()
        graph_us = sorted(a.elapsed_time(b) * 1e3 for a, b in ge)[len(ge) // 2]
        assert torch.equal(gout["scale_map"], out["scale_map"])
    except Exception as e:                                                   # noqa: BLE001
        graph_us = "unavailable: %s" % str(e)[:120]
    # end to end through the reference-facing NumPy classes (host buffers in / out)
    from riders_b200.modules.estimator import LeastSquaresEstimator, Optimizer
    from riders_b200.modules.interpolator import Interpolator2D
    from oracle import chain_ref
    sparse_inv, valid = chain_ref.sparse_inverse_ref(fr["sparse"])

    def host_step():
        if args.align == "st":
            ga = LeastSquaresEstimator(fr["mono"], sparse_inv, valid)
            ga.compute_scale_and_shift(); ga.apply_scale_and_shift(); ga.clamp_min_max(clamp_min=0.1, clamp_max=255.0)
        else:
            ga = Optimizer(fr["mono"], sparse_inv, valid, "inv")
            ga.optimize_scale(); ga.apply_scale(); ga.clamp_min_max(clamp_min=0.1, clamp_max=255.0)
        int_depth = ga.output.astype(np.float32)
        it = Interpolator2D(int_depth, sparse_inv, valid)
        it.generate_interpolated_scale_map("linear")                         # modules/interpolator.py:43-50 stores, returns nothing
        return it.interpolated_scale_map

    for _ in range(3):
        host_step()
    t0 = time.perf_counter()
    n_host = 20
    for _ in range(n_host):
        sm = host_step()
    host_us = (time.perf_counter() - t0) / n_host * 1e6
    # CPU reference of the same frame (LSQ / Brent + griddata), and parity of this run
    t0 = time.perf_counter()
    ref = chain_ref.sml_front_end_ref(fr["mono"], fr["sparse"], global_alignment=args.align)
    cpu_us = (time.perf_counter() - t0) * 1e6
    from oracle import interpolator_ref as ir
    coords, _, _ = ir.extract_knots_ref(ref["int_depth"], ref["sparse_inv"], ref["valid"])
    ok = ~ir.degenerate_pixel_mask(coords.T, H, W)
    got = out["scale_map"][0].cpu().numpy()
    ss = out["scale_shift"][0].cpu().numpy()
    parity = {"scale_rel_err": abs(ss[0] - ref["scale"]) / abs(ref["scale"]),
              "shift_rel_err": (abs(ss[1] - ref["shift"]) / abs(ref["shift"])) if ref["shift"] else 0.0,
              "int_depth_bit_exact": bool(np.array_equal(out["int_depth"][0].cpu().numpy(), ref["int_depth"])),
              "scale_map_max_rel_err_inside": float((np.abs(got - ref["scale_map"]) / np.abs(ref["scale_map"]))[ok].max()),
              "scale_map_outside_mismatch_px": int(((got == 1.0) != (ref["scale_map"] == 1.0))[ok].sum()),
              "host_class_path_equals_device_path": bool(np.array_equal(sm, got))}
    med = lat[len(lat) // 2]
    algorithmic = 4259840 * 2                                                # align 4.26 MB + interp 4.26 MB per frame
    peaks = load_peaks()
    print(json.dumps({"metric": "SML front-end latency @640x512, 200 radar points (configs[0])", "value": med, "unit": "us",
                      "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": med * 1e-3, "higher_is_better": False,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32 maps / f64 sums + barycentric", "data": "synthetic",
                      "config": {"workload": "configs[0] val_zju.py SML front-end: one 640x512 frame, 200 radar points: sparse inverse + "
                                             "mask, %s align + clamp, knot extraction, Delaunay, barycentric fill, scale scaffold, SML input "
                                             "assembly (352x288)" % ("LSQ" if args.align == "st" else "L1 Optimizer"),
                                 "global_alignment": args.align, "l2": "256 MB flush between timed iterations"},
                      "latency_us": {"p10": lat[len(lat) // 10], "median": med, "p90": lat[(9 * len(lat)) // 10], "cuda_graph_replay_median": graph_us},
                      "stage_us": stages, "gpu_launches": launches,
                      "e2e": {"value": host_us, "unit": "us", "note": "LeastSquaresEstimator/Optimizer + Interpolator2D classes, NumPy in / NumPy out",
                              "h2d_bytes_per_step": int(3 * H * W * 4 + H * W), "d2h_bytes_per_step": int(2 * H * W * 4)},
                      "roofline": {"kernel": "front-end (all launches)", "bound": "hbm", "achieved": algorithmic / (med * 1e-6) / 1e9, "peak": peaks["hbm"],
                                   "unit": "GB/s", "frac": algorithmic / (med * 1e-6) / 1e9 / peaks["hbm"], "traffic": None,
                                   "note": "one frame = 8.5 MB of algorithmic traffic = 1.3 us at the HBM peak: a single frame is latency-bound "
                                           "by construction (SURVEY H7); the figure of merit is the latency"},
                      "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "the same frame through oracle.chain_ref.sml_front_end_ref (NumPy LSQ / SciPy Brent + SciPy griddata)"},
                      "parity": parity}), flush=True)
    return 0


def run_train(args):
    """BASELINE configs[3]: RC-Net training step (train_rcnet_zju path: 4 frames x 30 radar points per GPU, patch 240x100, masked
    BCE, Adam) with the hand-written gather / LoFTR forward + backward kernels, bf16 autocast, and the gradient all-reduce over
    NCCL at N GPUs.  value = training frames/s over all ranks (weak scaling: every rank has its own 4 frames); the all-reduce is
    timed separately with CUDA events - it is the one collective that can limit this curve."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --workload train needs a CUDA device")
    from riders_b200 import dist as rdist, ops, synth
    from riders_b200.rcnet.rcnet_model import RCNetModel
    from riders_b200.train import TrainStep
    rank, world, local = rdist.init()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    pin_rank_cores(local, world)
    FR, KP = 4, 30                                            # train_rcnet_zju.py:28-30
    bf16 = not args.strict_fp32
    if args.strict_fp32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)                                      # identical initial weights on every rank (the all-reduce keeps them in step)
    model = RCNetModel(device=torch.device("cpu"), **MODEL_ARGS)
    model.device = device
    model.to(device)
    model.channels_last()
    torch.backends.cudnn.benchmark = True
    ts = TrainStep(model, bf16=bf16)
    # synthetic batch: padded frames, pad-shifted points + boxes, lidar crops around every radar point (depth of a smooth scene)
    rng = np.random.default_rng(1234 + rank)
    frames = [synth.chain_frame(1000 * rank + i, KP, H, W, PATCH) for i in range(FR)]
    img = np.stack([np.pad(f["image"], ((0, 0), (PATCH[0] // 2,) * 2, (PATCH[1] // 2,) * 2), mode="edge") for f in frames])
    pts = np.concatenate([f["points"] for f in frames]).astype(np.float32)
    gt = np.zeros((FR * KP, 1) + PATCH, np.float32)
    for r in range(FR * KP):                                   # sparse lidar crop: 10 % of the pixels carry a depth near / far from z
        keep = rng.uniform(size=PATCH) < 0.10
        gt[r, 0] = np.where(keep, pts[r, 2] + rng.normal(0, 1.0, PATCH), 0).astype(np.float32).clip(0)
    host = {"img": torch.from_numpy(img).pin_memory(), "pts": torch.from_numpy(pts).pin_memory(), "gt": torch.from_numpy(gt).pin_memory(),
            "boxes": [torch.from_numpy(f["boxes"]).pin_memory() for f in frames]}
    dev = {"img": host["img"].to(device).contiguous(memory_format=torch.channels_last), "pts": host["pts"].to(device),
           "gt": host["gt"].to(device), "boxes": [b.to(device) for b in host["boxes"]]}
    warm = max(args.warmup, 3)
    for _ in range(warm):
        loss = ts.step(dev["img"], dev["pts"], dev["boxes"], dev["gt"])
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ops.launch_count()
    rdist.barrier(); torch.cuda.synchronize()
    wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = ts.step(dev["img"], dev["pts"], dev["boxes"], dev["gt"], time_allreduce=True)
    e1.record()
    torch.cuda.synchronize(); rdist.barrier()
    wall1 = time.time()
    ms_total = rdist.max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    ar_ms = sum(a.elapsed_time(b) for a, b in ts.allreduce_events) / max(len(ts.allreduce_events), 1) if ts.allreduce_events else 0.0
    ar_ms = rdist.max_over_ranks(ar_ms)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    # end to end: pinned host batch -> H2D -> step -> loss read on the host, every step
    def e2e_step():
        d = {"img": host["img"].to(device, non_blocking=True).contiguous(memory_format=torch.channels_last),
             "pts": host["pts"].to(device, non_blocking=True), "gt": host["gt"].to(device, non_blocking=True),
             "boxes": [b.to(device, non_blocking=True) for b in host["boxes"]]}
        return float(ts.step(d["img"], d["pts"], d["boxes"], d["gt"]).item())
    e2e_step()
    rdist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = e2e_step()
    torch.cuda.synchronize(); rdist.barrier()
    e2e_ms = rdist.max_over_ranks((time.perf_counter() - t0) * 1e3)
    h2d = sum(host[k].numel() * 4 for k in ("img", "pts", "gt")) + sum(b.numel() * 4 for b in host["boxes"])
    n_param = sum(p.numel() for p in ts.params)
    if rank == 0:
        print(json.dumps({"metric": "RC-Net training frames/s @640x512 (configs[3]: 4 frames x 30 radar points per GPU)", "value": world * FR * args.steps / (ms_total * 1e-3),
                          "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "bf16 autocast (cuDNN convolutions, LoFTR tokens) / f32 master weights, loss, optimizer" if bf16 else "f32",
                          "data": "synthetic",
                          "config": {"workload": "configs[3] RC-Net training step (train_rcnet_zju path): forward with logits, masked BCE (pos_weight 2.5), "
                                                 "backward, gradient all-reduce, Adam(2e-4)", "frames_per_gpu_per_step": FR, "radar_points_per_frame": KP,
                                     "patch": list(PATCH), "parameters": n_param, "parallelism": "data parallel x%d, flat-bucket NCCL all-reduce of %.1f MB fp32 gradients" % (world, n_param * 4 / 1e6),
                                     "l2": "activations of a step (GBs) >> 126 MB L2"},
                          "clocks": clocks, "gpu_launches": int(launches), "loss": last,
                          "allreduce": {"ms_per_step": ar_ms, "bytes": n_param * 4, "share_of_step": ar_ms / (ms_total / args.steps) if world > 1 else 0.0},
                          "e2e": {"value": world * FR * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4}}), flush=True)
    rdist.shutdown()
    return 0


def workload_config(n_gpus, note=None, frames_per_gpu=FRAMES_PER_GPU, total=None, align="st", strict=False):
    total = total if total is not None else frames_per_gpu * n_gpus
    c = {"workload": "configs[1] RC-Net inference (run_rcnet_zju path) + SML front-end (%s + scale-map "
                     "interpolation): synthetic 640x512 thermal frames, 512 radar points/frame, thr 0.1"
                     % ("LSQ scale/shift align" if align == "st" else "L1 Optimizer scale align (the scripts' default)"),
         "frames_per_gpu_per_step": frames_per_gpu, "radar_points_per_frame": K_POINTS, "patch": list(PATCH),
         "global_frames_per_step": total, "parallelism": "frame-sharded x%d (no data-path collective)" % n_gpus,
         "global_alignment": align, "strict_fp32": bool(strict),
         "inputs": "as loaded: [0,255] frames + radar points in frame coordinates; prologue (shift/boxes/normalise/pad), "
                   "empty-frame retry check, scale scaffold and SML input assembly (352x288) inside the timed region",
         "l2": "per-step working set (11.4 GB gathered features, 0.8 GB logits) >> 126 MB L2; no flush needed"}
    if note:
        c["note"] = note
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--workload", default="chain", choices=["chain", "config1", "train"],
                    help="chain: configs[1] (default); config1: one frame, 200 points, SML front-end latency; "
                         "train: configs[3] RC-Net training step (bf16 autocast, gradient all-reduce)")
    ap.add_argument("--strict-fp32", action="store_true",
                    help="cuDNN / cuBLAS TF32 off and no TF32 decoder tail: every stage in fp32 (the 1e-4 parity configuration)")
    ap.add_argument("--align", default="st", choices=["st", "s"],
                    help="global alignment: st = least squares (north_star kernel 4), s = the scripts' default L1 Optimizer")
    ap.add_argument("--global-frames", type=int, default=0,
                    help="strong scaling (configs[2]): this many frames per step split over the ranks")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--profile", action="store_true", help=argparse.SUPPRESS)   # under ncu: 1 warm-up, no e2e / cpu legs
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "eager":
        return run_eager(args)
    if args.workload == "config1":
        return run_config1(args)
    if args.workload == "train":
        return run_train(args)
    args.warmup = 1 if args.profile else max(args.warmup, 3)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
    from riders_b200 import dist as rdist, ops, synth
    from riders_b200.pipeline import RadarDepthPipeline
    rank, world, local = rdist.init()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    pin_rank_cores(local, world)
    strong = args.global_frames > 0
    B = rdist.shard_frames(args.global_frames, rank, world)[1] - rdist.shard_frames(args.global_frames, rank, world)[0] \
        if strong else args.frames
    first_frame = rdist.shard_frames(args.global_frames, rank, world)[0] if strong else rank * B
    peaks = load_peaks()
    if args.strict_fp32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    tf32 = bool(torch.backends.cudnn.allow_tf32)

    model = build_model(device)
    # inputs "as loaded": [0,255] frames and radar points in frame coordinates; the prologue (shift, boxes, normalise,
    # pad), the empty-frame retry check and the SML input assembly all run inside the timed region
    # The alignment / interpolation half of the front-end can run on a side stream under the RC-Net forward
    # (RadarDepthPipeline(overlap_front_end=True), worth ~0.5 % of the step).  The bench keeps it on the main stream so that every
    # stage in `stages` is timed with the GPU to itself (round 1's in-step figures of those 20-100 us kernels were dominated by
    # waiting for SMs the decoder occupied); RIDERS_OVERLAP_FRONT_END=1 restores the overlap.
    overlap = os.environ.get("RIDERS_OVERLAP_FRONT_END", "0") == "1"
    pipe = RadarDepthPipeline(model, response_thr=THR, chunk_points=int(os.environ.get("RIDERS_CHUNK_POINTS", "2048")), raw_inputs=True,
                              retry_empty=True, sml_input=True, global_alignment=args.align, overlap_front_end=overlap)
    batch = synth.chain_batch(first_frame, B, K_POINTS, H, W, PATCH)
    batch["images"] = np.rint(batch["images"] * 255.0).astype(np.float32)
    batch["points"] = (batch["points"] - np.array([PATCH[1] // 2, PATCH[0] // 2, 0], np.float32)).astype(np.float32)
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("images", "points", "mono", "sparse")}
    dev = {k: v.to(device) for k, v in host.items()}
    offsets = batch["offsets"]

    def step_device():
        o = pipe.run_device(dev["images"], dev["points"], offsets, dev["mono"], dev["sparse"])
        # the path's only collective, INSIDE the timed step: per-frame means of the result maps -> one all-reduce of 5 doubles
        # over NCCL (asynchronous, no host sync); rank 0 divides after the timed region
        o["metric_sums"] = rdist.all_reduce_sums_device(torch.stack([
            o["depth"].double().mean(dim=(1, 2)).sum(), (o["response"] > 0).double().mean(dim=(1, 2)).sum(),
            o["int_depth"].double().mean(dim=(1, 2)).sum(), o["scale_map"].double().mean(dim=(1, 2)).sum(),
            torch.tensor(float(o["depth"].shape[0]), dtype=torch.float64, device=device)]))
        return o

    sampler = ClockSampler(local) if rank == 0 else None        # started early: nvidia-smi needs ~0.5 s to deliver samples
    # cuDNN picks its convolution algorithms by timing them in the first call: run that one (untimed, not counted in W)
    # without the side stream so that the alignment / interpolation kernels do not perturb the measurements it is based on
    if not args.profile:
        pipe.overlap_front_end = False
        out = step_device()
        torch.cuda.synchronize()
        pipe.overlap_front_end = overlap
    for _ in range(args.warmup):
        out = step_device()
    torch.cuda.synchronize()

    # ---------------------------------------------------------------- timed region (device-resident inputs)
    launches0 = ops.launch_count()
    rdist.barrier()
    torch.cuda.synchronize()
    wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.StageTimer() as timer:
        ev0.record()
        for _ in range(args.steps):
            out = step_device()
        ev1.record()
        torch.cuda.synchronize()
    rdist.barrier()
    wall1 = time.time()
    ms_total = rdist.max_over_ranks(ev0.elapsed_time(ev1))
    launches = ops.launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if sampler else None
    stages = timer.summary()

    # the final metric reduction (the path's only collective): per-frame means of the three maps
    vec = out["metric_sums"].cpu()
    n_frames = int(round(vec[-1].item()))
    means = vec[:-1] / vec[-1]
    assert int(out["interp_status"].abs().sum().item()) == 0, "interpolation reported a failed frame"

    # ---------------------------------------------------------------- end to end (pinned host in / out)
    if args.profile:
        print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps,
                          "stages": {k: v[1] / args.steps for k, v in stages.items()}}), flush=True)
        rdist.shutdown()
        return 0
    # The public call a user makes: RadarDepthPipeline.submit_host (pinned host buffers in, pinned host buffers out).  Two
    # steps are in flight: the copies of step k + 1 / k - 1 ride on copy streams under the compute of step k, and the
    # caller reads step k - 1's host result (handle.wait()) every step.  Every step's H2D and D2H are inside the timed region.
    def e2e_steps(n):
        pending = None
        for _ in range(n):
            handle = pipe.submit_host(host["images"], host["points"], offsets, host["mono"], host["sparse"], device)
            if pending is not None:
                pending.wait()                                 # the caller reads the host result of the previous step
            pending = handle
        res = pending.wait()
        torch.cuda.current_stream().wait_event(pending.done)
        return res

    e2e_steps(max(args.warmup, 3))                             # warm-up in the same two-in-flight pattern (allocator steady state)
    torch.cuda.synchronize()
    rdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hout, h2d, d2h = e2e_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    rdist.barrier()
    e2e_ms = rdist.max_over_ranks(e0.elapsed_time(e1))

    # ---------------------------------------------------------------- roofline of the dominant kernel
    # Every ops.* wrapper reports the algorithmic bytes (flops for LoFTR) of its call next to the CUDA events that
    # bracket it.  The gather additionally reads each frame's feature maps once.
    TENSOR = {"loftr_transformer", "decoder_tail", "subpixel_conv"}
    # the roofline object is for the dominant kernel among the SURVEY section-8 rows (the path this repo exists for);
    # the decoder glue / tail around the cuDNN convolutions (out of scope as kernels) stay in `stages`
    SECTION8 = {"roi_gather", "loftr_transformer", "confidence_reduce", "lsq_align", "l1_scale_align", "extract_knots",
                "interp_triangulate", "interp_fill", "interp_linear"}
    stage_report = {}
    for name, (calls, ms, amount) in stages.items():
        bound = "tensor" if name in TENSOR else "hbm"
        if name == "roi_gather":
            amount += args.steps * B * GATHER_MAP_BYTES_PER_FRAME
        ach = amount / (ms * 1e-3) / (1e12 if bound == "tensor" else 1e9) if ms > 0 else 0.0
        stage_report[name] = {"bound": bound, "calls_per_step": calls / args.steps, "ms_per_step": ms / args.steps,
                              "work_per_step": amount / args.steps, "achieved": round(ach, 2),
                              "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": round(ach / peaks[bound], 4)}
    ms_step = ms_total / args.steps
    dom = max((k for k in stage_report if k in SECTION8), key=lambda k: stage_report[k]["ms_per_step"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    d = stage_report[dom]
    roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": peaks[d["bound"]], "unit": d["unit"],
                "frac": d["frac"], "traffic": traffic, "peak_source": peaks["source"],
                "launches_per_step": d["calls_per_step"], "algorithmic_work_per_launch": d["work_per_step"] / max(d["calls_per_step"], 1),
                "avg_launch_ms": d["ms_per_step"] / max(d["calls_per_step"], 1), "share_of_step": round(d["ms_per_step"] / ms_step, 4)}
    roofline["scope"] = "dominant kernel among the SURVEY section-8 rows; decoder glue / tail (out of scope as kernels) are in `stages`"
    if dom == "loftr_transformer":
        roofline["note"] = ("algorithmic fp32 FLOPs of the 16 layer calls against the measured bf16 dense peak; the fp32-accurate "
                            "fp16 hi/lo operand split issues 3 MMAs per product, so tensor-pipe occupancy is ~3x this fraction")

    line = None
    if rank == 0:
        total_frames = args.global_frames if strong else world * B
        line = {"metric": METRIC, "value": total_frames * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": dtype_label(tf32), "data": "synthetic",
                "config": workload_config(world, frames_per_gpu=B, total=total_frames, align=args.align, strict=args.strict_fp32),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": total_frames * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": roofline, "stages": stage_report,
                "our_kernels_share_of_step": round(sum(s["ms_per_step"] for s in stage_report.values()) / ms_step, 4),
                "metrics": {"frames": n_frames, "mean_depth": means[0].item(), "fired_fraction": means[1].item(),
                            "mean_int_depth": means[2].item(), "mean_scale": means[3].item()},
                "loftr_impl": os.environ.get("RIDERS_LOFTR_IMPL", "default")}
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, range(os.cpu_count() or 1)) if hasattr(os, "sched_setaffinity") else None
            torch.set_num_threads(os.cpu_count() or 1)
            v, n, dt, refs = cpu_frame_rate(12.0, global_alignment=args.align)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "%d frame(s) (K=512) of the same batch through the oracle port, %.1f s" % (n, dt)}
            # parity of the configuration that was just timed (same model object, same flags), on the frames the CPU leg ran
            pipe.return_select = True
            pout = step_device()
            torch.cuda.synchronize()
            pipe.return_select = False
            line["parity"] = parity_block(pout, refs)
            line["parity"]["mode"] = dtype_label(tf32)
        print(json.dumps(line), flush=True)
    rdist.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
