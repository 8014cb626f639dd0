"""Oracle for the whole chain (BASELINE configs 2/3) on the CPU, one frame at a time like the reference.

  RC-Net quasi-dense depth   run_rcnet_zju.py:204-263 -> rcnet_main.forward_output (oracle.rcnet_ref)
  inverse sparse depth+mask  val_zju.py:130-133
  LSQ alignment + clamp      val_zju.py:146-155 -> modules/estimator.py:5-29,108-125 (oracle.estimator_ref)
  scale scaffold             modules/interpolator.py:20-50 (oracle.interpolator_ref, SciPy griddata)

Used by tests/ as the checker of riders_b200.pipeline, and - timed on the host cores - as the
``cpu_baseline`` / ``--impl reference`` leg of bench.py (kind "port": the reference is pure Python
and cannot travel to the GPU box).  Test infrastructure only.
"""
import numpy as np
import torch

from . import estimator_ref as er
from . import interpolator_ref as ir
from .rcnet_ref import forward_output_ref


def sparse_inverse_ref(sparse_depth, min_depth=0.0, max_depth=100.0):
    """val_zju.py:130-133 in NumPy float32."""
    sd = np.asarray(sparse_depth, dtype=np.float32).copy()
    valid = (sd < np.float32(max_depth)) & (sd > np.float32(min_depth))
    sd[~valid] = np.inf
    with np.errstate(divide="ignore"):
        return (np.float32(1.0) / sd).astype(np.float32), valid


def sml_front_end_ref(mono_pred, sparse_depth, min_depth=0.0, max_depth=100.0, clamp_min=0.1, clamp_max=255.0,
                      acc_dtype=np.float64, interpolate=True, global_alignment="st", mono_type="inv"):
    """Alignment + scale scaffold of one frame.  acc_dtype=float64 is 'the reference fed float64
    inputs' (SURVEY H2); float32 the as-shipped behaviour.  global_alignment 'st' = val_zju.py:146-155
    (least squares), 's' = val_zju.py:156-167 (Optimizer: SciPy bounded Brent on the masked L1 objective,
    fed float64 inputs; output = estimate * float32(scale), the numpy 1.23 casting of the pinned environment)."""
    sparse_inv, valid = sparse_inverse_ref(sparse_depth, min_depth, max_depth)
    if global_alignment == "s":
        s = er.optimize_scale_ref(np.asarray(mono_pred, np.float64), np.asarray(sparse_inv, np.float64), valid, mono_type)
        t = 0.0
    elif global_alignment == "st":
        s, t = er.scale_shift_ls_ref(mono_pred, sparse_inv, valid, acc_dtype)
    else:
        raise NotImplementedError(global_alignment)
    int_depth = er.apply_clamp_ref(mono_pred, s, t, clamp_min, clamp_max)
    out = {"scale": float(s), "shift": float(t), "int_depth": int_depth, "valid": valid, "sparse_inv": sparse_inv}
    if interpolate:
        out["scale_map"] = ir.interpolated_scale_map_ref(int_depth, sparse_inv, valid)
    return out


@torch.no_grad()
def chain_frame_ref(model_cpu, frame, response_thr=0.1, patch=None, **kw):
    """frame: dict from riders_b200.synth.chain_frame.  model_cpu: RCNetModel built on the CPU (its
    CNN / MLP / decoder are plain torch modules; gather + LoFTR + reduce are recomputed by
    oracle.rcnet_ref with library CPU ops)."""
    image = torch.from_numpy(frame["image"])[None]
    pts = torch.from_numpy(frame["points"])
    boxes = [torch.from_numpy(frame["boxes"])]
    want_conf = kw.pop("want_conf", False)
    res = forward_output_ref(model_cpu, image, pts, boxes, response_thr, want_conf=want_conf)
    depth, response, select = res[:3]
    out = sml_front_end_ref(frame["mono"], frame["sparse"], **kw)
    out.update(depth=depth[0].numpy(), response=response[0].numpy(), select=select[0].numpy())
    if want_conf:
        out["conf"] = res[3]                       # (K, ph, pw) confidences, for the top-2 margin of the select rule
    return out
