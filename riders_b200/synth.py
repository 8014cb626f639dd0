"""Seeded synthetic inputs of SURVEY 8(d) (frame f uses seed f).  NumPy only, so the same
arrays feed the CUDA path, the oracle and the CPU baseline."""
import numpy as np


def thermal_frame(seed, H=512, W=640):
    """(3,H,W) float32 in [0,1]: three identical channels of smooth noise, uint8-quantised like the
    loader's int()->/255 normalisation (RCNet/rcnet_transforms.py:100-103)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.zeros((H, W), np.float32)
    for _ in range(4):
        fx, fy, ph = rng.uniform(0.002, 0.02), rng.uniform(0.002, 0.02), rng.uniform(0, 6.28)
        img += np.sin(xx * fx * 6.28 + yy * fy * 6.28 + ph).astype(np.float32)
    img = 128 + 30 * img + rng.normal(0, 4, (H, W)).astype(np.float32)
    img = np.clip(np.floor(img), 0, 255).astype(np.float32) / np.float32(255)
    return np.repeat(img[None], 3, axis=0)


def radar_points(seed, K, H=512, W=640, patch=(240, 100)):
    """K unique integer pixels in row-major scan order with z ~ U(1.5,100) quantised to 1/256 m
    (PNG codec, data/data_utils.py:128-143).  Returns pad-shifted (K,3) float32 (x+pw/2, y+ph/2, z)
    and the (K,4) boxes of run_rcnet_zju.py:227-232."""
    rng = np.random.default_rng(1000003 * (seed + 1))
    flat = np.sort(rng.choice(H * W, size=K, replace=False))
    ys, xs = (flat // W).astype(np.float32), (flat % W).astype(np.float32)
    z = (np.round(rng.uniform(1.5, 100.0, K) * 256) / 256).astype(np.float32)
    px, py = patch[1] // 2, patch[0] // 2
    pts = np.stack([xs + px, ys + py, z], axis=1).astype(np.float32)
    boxes = np.stack([pts[:, 0] - px, pts[:, 1] - py, pts[:, 0] + px, pts[:, 1] + py], axis=1).astype(np.float32)
    return pts, boxes


def confidences(seed, K, patch=(240, 100), margin=None):
    """(K,ph,pw) sigmoid(N(-1.5,2)) so ~35% fall below 0.1; ``margin`` enforces distinct values by
    adding a tiny per-point offset (for the bit-exact select test)."""
    rng = np.random.default_rng(7919 * (seed + 1))
    c = (1.0 / (1.0 + np.exp(-rng.normal(-1.5, 2.0, (K,) + tuple(patch))))).astype(np.float32)
    if margin:
        c = (c + (np.arange(K, dtype=np.float32) * np.float32(margin))[:, None, None]).astype(np.float32)
    return c


def alignment_inputs(seed, K, H=512, W=640, s=0.07, t=0.002):
    """Mono prior (relative inverse depth, /256-quantised), sparse inverse-depth target (0 where
    invalid) and its mask: SURVEY 8(d)."""
    rng = np.random.default_rng(15485863 * (seed + 1))
    yy, xx = np.mgrid[0:H, 0:W]
    d_true = 20 + 15 * np.sin(xx / 37.0) + 10 * np.cos(yy / 23.0) + rng.uniform(0, 2, (H, W))
    pred = (np.round(((1.0 / d_true - t) / s) * 256 * 64) / (256 * 64)).astype(np.float32)
    idx = rng.choice(H * W, K, replace=False)
    mask = np.zeros(H * W, bool)
    mask[idx] = True
    mask = mask.reshape(H, W)
    target = np.zeros((H, W), np.float32)
    target[mask] = ((1.0 / d_true[mask]) * (1 + 0.02 * rng.standard_normal(K))).astype(np.float32)
    return pred, target, mask


def chain_frame(seed, K, H=512, W=640, patch=(240, 100), s=0.07, t=0.002):
    """One frame of the full chain (BASELINE configs 2/3): thermal image, K radar points whose
    depths are noisy samples of a smooth scene, the sparse radar depth map the SML loader would
    read (data/UTV_dataset.py:14-17: PNG/256, 0 = empty) and the mono prior (relative inverse
    depth of the same scene, /256-quantised).  Returns dict of float32 arrays."""
    rng = np.random.default_rng(32452843 * (seed + 1))
    pts, boxes = radar_points(seed, K, H, W, patch)
    px, py = patch[1] // 2, patch[0] // 2
    yy, xx = np.mgrid[0:H, 0:W]
    d_true = 25 + 15 * np.sin(xx / 41.0 + seed) + 8 * np.cos(yy / 29.0) + rng.uniform(0, 1.5, (H, W))
    mono = (np.round(((1.0 / d_true - t) / s) * 256 * 64) / (256 * 64)).astype(np.float32)
    xs, ys = (pts[:, 0] - px).astype(np.int64), (pts[:, 1] - py).astype(np.int64)
    z = d_true[ys, xs] * (1 + 0.02 * rng.standard_normal(K))
    z = (np.floor(np.clip(z, 1.5, 99.0) * 256) / 256).astype(np.float32)      # uint32(z*256) PNG codec
    pts = pts.copy()
    pts[:, 2] = z
    sparse = np.zeros((H, W), np.float32)
    sparse[ys, xs] = z
    return {"image": thermal_frame(seed, H, W), "points": pts, "boxes": boxes, "mono": mono, "sparse": sparse}


def chain_batch(first_seed, B, K, H=512, W=640, patch=(240, 100)):
    """B frames stacked: images (B,3,H,W), points (B*K,3), offsets list, mono (B,H,W), sparse (B,H,W)."""
    fr = [chain_frame(first_seed + b, K, H, W, patch) for b in range(B)]
    return {"images": np.stack([f["image"] for f in fr]),
            "points": np.concatenate([f["points"] for f in fr]).astype(np.float32),
            "offsets": [b * K for b in range(B + 1)],
            "mono": np.stack([f["mono"] for f in fr]),
            "sparse": np.stack([f["sparse"] for f in fr]),
            "boxes": [f["boxes"] for f in fr]}
