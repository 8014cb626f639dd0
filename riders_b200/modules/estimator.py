"""Global scale/shift alignment, host side (mirror of the reference's ``modules/estimator.py``).

Same names and call signatures: ``compute_scale_and_shift_ls`` (estimator.py:5-29),
``LeastSquaresEstimator`` (:90-125) and ``Optimizer`` (:136-185).  NumPy in, NumPy out like the
reference (one pinned H2D, the kernels, one D2H); torch CUDA tensors are accepted too and then
stay on the device.  The five normal-equation sums are accumulated in fp64 by a deterministic
block tree (BASELINE north_star kernel 4), so ``scale``/``shift`` match the reference fed
float64 inputs to ~1e-15 and the as-shipped float32 reference to its own rounding error.
The RANSAC variant (estimator.py:32-87, never called by the shipped scripts) is provided with the reference's own use of
NumPy's global RNG: the valid (prediction, target) pairs are compacted on the device in ``np.where`` order, the 60 x 5-sample
fits and inlier counts run over those few hundred pairs.
"""
import numpy as np
import torch

from .. import ops


def _to_device(a, dtype=None):
    if torch.is_tensor(a):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.pin_memory().cuda(non_blocking=True) if torch.cuda.is_available() else t.cuda()
    return t


def _as3d(t):
    return t.unsqueeze(0) if t.dim() == 2 else t


def compute_scale_and_shift_ls(prediction, target, mask):
    """Returns (x_0, x_1) as 0-d float32 arrays (or (B,) arrays for batched (B,H,W) input)."""
    batched = np.ndim(prediction) == 3
    p, t, m = _as3d(_to_device(prediction, torch.float32)), _as3d(_to_device(target, torch.float32)), _as3d(_to_device(mask))
    _, ss, _ = ops.lsq_align(p, t, m, want_out=False)
    ss = ss.cpu().numpy()
    if batched:
        return ss[:, 0].astype(np.float32), ss[:, 1].astype(np.float32)
    return np.float32(ss[0, 0]).reshape(()), np.float32(ss[0, 1]).reshape(())


def compute_scale_and_shift_ransac(prediction, target, mask, num_iterations, sample_size, inlier_threshold,
                                   inlier_ratio_threshold):
    """estimator.py:32-87.  Same draws from ``np.random`` (one ``choice(valid_count, sample_size, replace=False)`` per
    iteration), same update / early-exit rules; the sums run over the compacted valid pairs instead of the masked full maps
    (the masked-out terms of the reference are exact zeros)."""
    best_scale, best_shift, best_inlier_count = 0.0, 0.0, 0
    p, t, m = _to_device(prediction, torch.float32), _to_device(target, torch.float32), _to_device(mask).bool()
    pv, tv = p[m].cpu().numpy(), t[m].cpu().numpy()               # row-major order of the valid pixels = np.where(mask)
    valid_count = pv.shape[0]
    for _ in range(num_iterations):
        if valid_count < sample_size:
            break
        indices = np.random.choice(valid_count, size=sample_size, replace=False)
        ps, ts_ = pv[indices], tv[indices]
        a_00, a_01, a_11 = np.sum(ps * ps), np.sum(ps), np.float32(sample_size)
        b_0, b_1 = np.sum(ps * ts_), np.sum(ts_)
        det = a_00 * a_11 - a_01 * a_01
        x_0, x_1 = np.zeros_like(b_0), np.zeros_like(b_1)
        if det > 0:
            x_0 = (a_11 * b_0 - a_01 * b_1) / det
            x_1 = (-a_01 * b_0 + a_00 * b_1) / det
        residuals = np.abs(pv * x_0 + x_1 - tv)
        inlier_count = np.sum(residuals < inlier_threshold)
        if inlier_count > best_inlier_count:
            best_scale, best_shift, best_inlier_count = x_0, x_1, inlier_count
            if inlier_count / valid_count > inlier_ratio_threshold:
                break
    print('best_inlier_count: ', best_inlier_count)
    print('inlier_ratio: ', best_inlier_count / valid_count if valid_count else float('nan'))
    return best_scale, best_shift


class LeastSquaresEstimator(object):
    def __init__(self, estimate, target, valid):
        self.estimate = estimate
        self.target = target
        self.valid = valid
        # to be computed
        self.scale = 1.0
        self.shift = 0.0
        self.output = None
        self.scale64 = None       # fp64 solution of the normal equations (extension)
        self.shift64 = None
        self._dev = None

    def _device_inputs(self):
        if self._dev is None:
            self._dev = (_as3d(_to_device(self.estimate, torch.float32)),
                         _as3d(_to_device(self.target, torch.float32)),
                         _as3d(_to_device(self.valid)))
        return self._dev

    def compute_scale_and_shift_ran(self, num_iterations=60, sample_size=5, inlier_threshold=0.02, inlier_ratio_threshold=0.8):
        """estimator.py:101-106."""
        self.scale, self.shift = compute_scale_and_shift_ransac(self.estimate, self.target, self.valid, num_iterations,
                                                                sample_size, inlier_threshold, inlier_ratio_threshold)
        self._applied = None

    def compute_scale_and_shift(self):
        # one pass pair on the device: sums -> solve -> unclamped output kept for apply_scale_and_shift()
        p, t, m = self._device_inputs()
        _, ss, out = ops.lsq_align(p, t, m, want_out=True)
        ss = ss.cpu().numpy()
        self.scale64, self.shift64 = ss[0, 0], ss[0, 1]
        self.scale = np.float32(ss[0, 0]).reshape(())
        self.shift = np.float32(ss[0, 1]).reshape(())
        self._applied = (self.scale, self.shift, out)

    def apply_scale_and_shift(self):
        cached = getattr(self, '_applied', None)
        if cached is not None and cached[0] is self.scale and cached[1] is self.shift:
            out = cached[2]
        else:
            # caller overwrote scale/shift by hand: plain elementwise apply (two float32 roundings)
            p = self._device_inputs()[0]
            out = p * float(np.float32(self.scale)) + float(np.float32(self.shift))
        self._store(out.clone())

    def clamp_min_max(self, clamp_min=None, clamp_max=None):
        """In place on ``self.output`` like the reference (estimator.py:115-125)."""
        if self.output is None:
            raise TypeError("apply_scale_and_shift() must run before clamp_min_max()")
        hi = np.float32(1.0 / clamp_min) if (clamp_min is not None and clamp_min > 0) else None
        lo = np.float32(1.0 / clamp_max) if clamp_max is not None else None
        if torch.is_tensor(self.output):
            if hi is not None:
                self.output.clamp_(max=float(hi))
                assert float(self.output.max()) <= hi
            if lo is not None:
                self.output.clamp_(min=float(lo))
            return
        if hi is not None:
            self.output[self.output > hi] = hi
            assert np.max(self.output) <= hi
        if lo is not None:
            self.output[self.output < lo] = lo

    def align(self, clamp_min=None, clamp_max=None):
        """compute + apply + clamp in one device pass (what val_zju.py:147-155 does in 4 calls)."""
        hi = float(np.float32(1.0 / clamp_min)) if (clamp_min is not None and clamp_min > 0) else None
        lo = float(np.float32(1.0 / clamp_max)) if clamp_max is not None else None
        self._apply(hi, lo, solve=True)
        return self.output

    def _apply(self, hi, lo, solve=True):
        p, t, m = self._device_inputs()
        _, ss, out = ops.lsq_align(p, t, m, clamp_hi=hi, clamp_lo=lo, want_out=True)
        ssh = ss.cpu().numpy()
        self.scale64, self.shift64 = ssh[0, 0], ssh[0, 1]
        self.scale = np.float32(ssh[0, 0]).reshape(())
        self.shift = np.float32(ssh[0, 1]).reshape(())
        self._store(out)

    def _store(self, out):
        if torch.is_tensor(self.estimate):
            self.output = out.reshape(self.estimate.shape)
            if not self.estimate.is_cuda:
                self.output = self.output.cpu()
        else:
            self.output = out.reshape(np.shape(self.estimate)).cpu().numpy()


def objective_function(x_0, prediction, target, mask):
    """estimator.py:129-134: masked L1 error of a scale, evaluated on the device (fp64 accumulate)."""
    p, t, m = _as3d(_to_device(prediction, torch.float32)), _as3d(_to_device(target, torch.float32)), _as3d(_to_device(mask))
    err = (m.to(torch.float64) * (float(x_0) * p.to(torch.float64) - t.to(torch.float64)).abs()).sum()
    return float(err.item())


class Optimizer(object):
    """estimator.py:136-185: scale-only alignment by bounded Brent on the masked L1 objective (the
    scripts' default, val_zju.py:156-167).  One launch group: compaction of the valid pixels, the
    fminbound iteration (fp64 objective over the compacted pairs), apply + clamp."""

    def __init__(self, estimate, target, valid, depth_type):
        self.estimate = estimate
        self.target = target
        self.valid = valid
        self.depth_type = depth_type
        # to be computed
        self.scale = 1.0
        self.output = None
        self.nfev = None
        self._dev = None

    def _device_inputs(self):
        if self._dev is None:
            self._dev = (_as3d(_to_device(self.estimate, torch.float32)),
                         _as3d(_to_device(self.target, torch.float32)),
                         _as3d(_to_device(self.valid)))
        return self._dev

    def _bounds(self):
        return (0.01, 0.3) if self.depth_type == 'inv' else (0.5, 1.6)      # estimator.py:147-150

    def optimize_scale(self):
        p, t, m = self._device_inputs()
        scale, _, nfev, _ = ops.l1_scale_align(p, t, m, bounds=self._bounds(), want_out=False)
        self.scale = np.float64(scale[0].item())                             # result.x is a NumPy float64
        self.nfev = int(nfev[0].item())

    def apply_scale(self):
        p = self._device_inputs()[0]
        out = p * float(np.float32(self.scale))                              # float32 array * scalar (numpy 1.23 casting)
        self._store(out)

    def clamp_min_max(self, clamp_min=None, clamp_max=None):
        LeastSquaresEstimator.clamp_min_max(self, clamp_min, clamp_max)

    def clamp_min_max_pos(self, clamp_min=None, clamp_max=None):
        """estimator.py:178-185 (positive-depth variant)."""
        if clamp_min is not None and clamp_min >= 0:
            self.output[self.output < clamp_min] = clamp_min
        if clamp_max is not None:
            self.output[self.output > clamp_max] = clamp_max

    def align(self, clamp_min=None, clamp_max=None):
        """optimize + apply + clamp in one device call (val_zju.py:157-166 in 4 calls)."""
        hi = float(np.float32(1.0 / clamp_min)) if (clamp_min is not None and clamp_min > 0) else None
        lo = float(np.float32(1.0 / clamp_max)) if clamp_max is not None else None
        p, t, m = self._device_inputs()
        scale, _, nfev, out = ops.l1_scale_align(p, t, m, bounds=self._bounds(), clamp_hi=hi, clamp_lo=lo)
        self.scale = np.float64(scale[0].item())
        self.nfev = int(nfev[0].item())
        self._store(out)
        return self.output

    _store = LeastSquaresEstimator._store
