"""BASELINE.json configs[4]: 1280x1024 synthetic frames with 128-4096 radar points per frame - scaling of the
gather / reduce / interpolation stages with radar density.  JSON lines on stdout (CUDA events, L2 flushed)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riders_b200 import ops, synth          # noqa: E402
from tools.microbench import timeit, HBM      # noqa: E402

DEV = torch.device("cuda:0")
H, W, PATCH = 1024, 1280, (240, 100)


def main():
    torch.cuda.set_device(DEV)
    Hp, Wp = H + PATCH[0], W + PATCH[1]
    chans = [32, 64, 128, 128, 128]
    scales = [1 / 2., 1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    sizes = [(120, 50), (60, 25), (30, 12), (15, 6), (7, 3)]
    dims = []
    h, w = Hp, Wp
    for _ in range(5):
        h, w = (h + 1) // 2, (w + 1) // 2
        dims.append((h, w))
    g = torch.Generator(device=DEV).manual_seed(0)
    feats = [torch.randn((1, c, hh, ww), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
             for c, (hh, ww) in zip(chans, dims)]
    map_bytes = sum(c * hh * ww for c, (hh, ww) in zip(chans, dims)) * 4
    per_point = sum(c * oh * ow for c, (oh, ow) in zip(chans, sizes)) * 4
    print(json.dumps({"config": "stress sweep 1280x1024", "hbm_peak_gbs": HBM}), flush=True)
    with torch.no_grad():
        for K in (128, 256, 512, 1024, 2048, 4096):
            pts, boxes = synth.radar_points(K, K, H, W, PATCH)
            rois = ops.boxes_to_rois([torch.from_numpy(boxes).to(DEV)], DEV)
            ms, _ = timeit(lambda: ops.roi_gather(feats, rois, scales, sizes, max_roi=(PATCH[1], PATCH[0])), iters=5)
            nb = K * (per_point + 20) + map_bytes
            rec = {"K": K, "gather_ms": round(ms, 4), "gather_GBps": round(nb / ms / 1e6, 1), "gather_frac_hbm": round(nb / ms / 1e6 / HBM, 3)}
            conf = torch.sigmoid(torch.randn((K,) + PATCH, device=DEV, generator=g) * 2 - 1.5)
            p = torch.from_numpy(pts).to(DEV)
            offs = torch.tensor([0, K], dtype=torch.int32, device=DEV)
            ms, _ = timeit(lambda: ops.confidence_reduce(conf, p, offs, H, W, 0.1), iters=5)
            nb = K * (PATCH[0] * PATCH[1] * 4 + 12) + 3 * H * W * 4
            rec.update(reduce_ms=round(ms, 4), reduce_GBps=round(nb / ms / 1e6, 1), reduce_frac_hbm=round(nb / ms / 1e6 / HBM, 3))
            flat = np.sort(np.random.default_rng(K).choice(H * W, K, replace=False))
            xy = torch.from_numpy(np.stack([flat % W, flat // W], 1).astype(np.int32))[None].to(DEV)
            val = torch.rand(1, K, device=DEV) + 0.5
            cnt = torch.tensor([K], dtype=torch.int32, device=DEV)
            with ops.StageTimer() as t:
                for _ in range(5):
                    out, status = ops.interp_linear(xy, val, cnt, H, W, 1.0)
            s = t.summary()
            assert int(status[0]) == 0
            ims = s["interp_linear"][1] / 5          # stars + fill value + rasteriser (two launches)
            rec.update(interp_ms=round(ims, 4), interp_GBps=round((H * W * 4 + 12 * K) / ims / 1e6, 1),
                       interp_frac_hbm=round((H * W * 4 + 12 * K) / ims / 1e6 / HBM, 3))
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
