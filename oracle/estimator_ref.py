"""Oracle for SURVEY 8(a) rows A6/A7 (+ 'next' row N1): global scale/shift alignment.

Follows modules/estimator.py of the reference:
  * ``compute_scale_and_shift_ls``   lines 5-29   (2x2 normal equations, det > 0 guard)
  * ``LeastSquaresEstimator``        lines 90-125 (apply line 113, in-place clamp lines 115-125)
  * ``objective_function``/``Optimizer`` lines 129-160 (masked L1, bounded Brent via
    scipy.optimize.minimize_scalar, bounds (0.01,0.3) 'inv' / (0.5,1.6) otherwise)

``dtype`` selects the accumulation type: float32 restates the shipped
behaviour (inputs are float32 arrays, np.sum pairwise in float32); float64 is
"the same reference function fed float64 inputs", which is the parity target
for the fp64-accumulating CUDA kernel (SURVEY hard part H2).
Test infrastructure only.
"""
import numpy as np


def normal_equation_sums(prediction, target, mask, dtype=np.float64):
    p = np.asarray(prediction).astype(dtype)
    t = np.asarray(target).astype(dtype)
    m = np.asarray(mask).astype(dtype)
    a00 = np.sum(m * p * p)
    a01 = np.sum(m * p)
    a11 = np.sum(m)
    b0 = np.sum(m * p * t)
    b1 = np.sum(m * t)
    return a00, a01, a11, b0, b1


def scale_shift_ls_ref(prediction, target, mask, dtype=np.float64):
    a00, a01, a11, b0, b1 = normal_equation_sums(prediction, target, mask, dtype)
    det = a00 * a11 - a01 * a01
    if det > 0:
        return (a11 * b0 - a01 * b1) / det, (-a01 * b0 + a00 * b1) / det
    return dtype(0), dtype(0)


def apply_clamp_ref(estimate, scale, shift, clamp_min=None, clamp_max=None):
    """estimator.py:112-125 in float32: out = est*scale + shift (two roundings), then clamp."""
    est = np.asarray(estimate, dtype=np.float32)
    out = est * np.float32(scale) + np.float32(shift)
    if clamp_min is not None and clamp_min > 0:
        hi = np.float32(1.0 / clamp_min)
        out[out > hi] = hi
    if clamp_max is not None:
        lo = np.float32(1.0 / clamp_max)
        out[out < lo] = lo
    return out


def l1_objective_ref(x0, prediction, target, mask):
    return np.sum(mask * np.abs(x0 * prediction - target))


def optimize_scale_ref(prediction, target, mask, depth_type="inv"):
    from scipy.optimize import minimize_scalar
    bounds = (0.01, 0.3) if depth_type == "inv" else (0.5, 1.6)
    res = minimize_scalar(l1_objective_ref, args=(prediction, target, mask), bounds=bounds)
    return res.x
