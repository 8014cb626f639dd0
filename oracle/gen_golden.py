"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (MMOCKING/RIDERS).

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden
The reference has no tests or golden vectors of its own, so these fixtures -
outputs of the reference's own functions on seeded synthetic inputs - are what pins the oracle.
Import recipe: SURVEY 8(c) (matplotlib is stubbed because utils/log_utils.py:4 imports it).
Test infrastructure only.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("RIDERS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise RuntimeError("reference checkout not found at %s" % REF)
    for p in (os.path.join(REF, "RCNet"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import linear_attention, networks, rcnet_main, rcnet_model          # noqa: E401
    from modules import estimator, interpolator
    return dict(linear_attention=linear_attention, networks=networks, rcnet_main=rcnet_main,
                rcnet_model=rcnet_model, estimator=estimator, interpolator=interpolator)


def state_np(sd, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def gen_roi_pool(ref):
    import torchvision
    rng = np.random.default_rng(10)
    # a miniature RC-Net pyramid: padded frame 104 x 136, patch 64 x 32 (pads 32 / 16)
    H, W, ph, pw = 104, 136, 64, 32
    pts = np.array([[16, 32], [W - 17, H - 33], [40, 50], [77, 61], [16, H - 33], [W - 17, 32]], np.float32)
    boxes = np.stack([pts[:, 0] - pw // 2, pts[:, 1] - ph // 2, pts[:, 0] + pw // 2, pts[:, 1] + ph // 2], 1)
    boxes2 = np.array([[3.4, 7.6, 30.2, 70.9], [-5.0, -3.0, 12.0, 9.0], [100.0, 80.0, 180.0, 140.0]], np.float32)
    out = {"boxes0": boxes, "boxes1": boxes2, "patch": np.array([ph, pw])}
    scales = [1 / 2.0, 1 / 4.0, 1 / 8.0, 1 / 16.0, 1 / 32.0]
    chans = [4, 8, 8, 8, 8]
    for l, (s, c) in enumerate(zip(scales, chans)):
        hl, wl = int(np.ceil(H * s)), int(np.ceil(W * s))
        feat = rng.standard_normal((2, c, hl, wl)).astype(np.float32)
        osz = (max(int(ph * s), 1), max(int(pw * s), 1))
        res = torchvision.ops.roi_pool(torch.from_numpy(feat), [torch.from_numpy(boxes), torch.from_numpy(boxes2)],
                                       spatial_scale=s, output_size=osz).numpy()
        out["feat%d" % l] = feat
        out["out%d" % l] = res
        out["scale%d" % l] = np.float32(s)
        out["osize%d" % l] = np.array(osz)
    np.savez_compressed(os.path.join(OUT, "roi_pool.npz"), **out)


def gen_loftr(ref):
    la = ref["linear_attention"]
    out = {}
    for tag, d_model, n_layers, N, L, S in (("small", 32, 2, 3, 5, 7), ("rcnet", 128, 1, 7, 21, 21)):
        torch.manual_seed(20 + d_model)
        tf = la.LocalFeatureTransformer(["self", "cross"], n_layers=n_layers, d_model=d_model).eval()
        # non-trivial LayerNorm affine so gamma/beta are exercised
        for layer in tf.layers:
            for nm in (layer.norm1, layer.norm2):
                torch.nn.init.normal_(nm.weight, 1.0, 0.2)
                torch.nn.init.normal_(nm.bias, 0.0, 0.2)
        Lx = L if tag == "rcnet" else L
        f0 = torch.randn(N, Lx, d_model) * 0.7
        f1 = torch.relu(torch.randn(N, S if tag == "rcnet" else L, d_model))   # 'self'+'cross' need L==S only for shape of f1 self
        with torch.no_grad():
            o0, o1 = tf(f0, f1)
            # single layer + bare attention too
            l0 = tf.layers[0](f0, f1)
            q = torch.randn(N, L, 8, d_model // 8)
            k = torch.randn(N, S, 8, d_model // 8)
            v = torch.randn(N, S, 8, d_model // 8)
            qm = (torch.rand(N, L) > 0.3).float()
            km = (torch.rand(N, S) > 0.3).float()
            a_plain = la.LinearAttention()(q, k, v)
            a_mask = la.LinearAttention()(q, k, v, q_mask=qm, kv_mask=km)
        out.update({tag + "/" + k_: v_ for k_, v_ in state_np(tf.state_dict()).items()})
        for name, val in (("f0", f0), ("f1", f1), ("o0", o0), ("o1", o1), ("layer0", l0), ("q", q), ("k", k),
                          ("v", v), ("qm", qm), ("km", km), ("attn", a_plain), ("attn_masked", a_mask)):
            out["%s/%s" % (tag, name)] = val.numpy()
        out[tag + "/n_layers"] = np.array(n_layers)
    np.savez_compressed(os.path.join(OUT, "loftr.npz"), **out)


class _FakeModel(object):
    """Stands in for RCNetModel so the reference's own forward_output tail runs on given crops."""

    def __init__(self, patch, conf):
        self.input_patch_size_image = patch
        self._conf = conf

    def forward(self, image, point, bounding_boxes, return_logits):
        assert return_logits is False
        return self._conf


def gen_reduce(ref):
    fo = ref["rcnet_main"].forward_output
    rng = np.random.default_rng(30)
    out = {}
    cases = {
        "basic": dict(H=48, W=64, ph=24, pw=10, K=9, thr=0.1),
        "single": dict(H=40, W=40, ph=16, pw=8, K=1, thr=0.1),
        "allbelow": dict(H=32, W=48, ph=16, pw=8, K=4, thr=0.99),
        "dups": dict(H=32, W=48, ph=16, pw=12, K=6, thr=0.3),
    }
    for tag, c in cases.items():
        H, W, ph, pw, K, thr = c["H"], c["W"], c["ph"], c["pw"], c["K"], c["thr"]
        xs = rng.integers(0, W, K)
        ys = rng.integers(0, H, K)
        if tag == "basic":
            xs[:4] = [0, W - 1, 0, W - 1]
            ys[:4] = [0, 0, H - 1, H - 1]
        if tag == "dups":
            xs[1], ys[1] = xs[0], ys[0]
        z = rng.uniform(1.5, 100.0, K)
        pts = np.stack([xs + pw // 2, ys + ph // 2, z], 1).astype(np.float32)
        conf = (1 / (1 + np.exp(-rng.normal(-1.5, 2.0, (K, 1, ph, pw))))).astype(np.float32)
        if tag == "basic":
            conf[0, 0, 0, 0] = np.float32(thr)        # exactly-at-threshold value survives (strict <)
        if tag == "allbelow":
            conf = np.minimum(conf, np.float32(0.98))
        image = torch.zeros(1, 3, H, W)
        with torch.no_grad():
            d, r = fo(_FakeModel([ph, pw], torch.from_numpy(conf)), image, torch.from_numpy(pts)[None],
                      None, response_thr=thr, device=torch.device("cpu"))
        for name, val in (("conf", conf), ("pts", pts), ("depth", d.numpy()), ("response", r.numpy()),
                          ("hw", np.array([H, W])), ("thr", np.float32(thr))):
            out["%s/%s" % (tag, name)] = val
    np.savez_compressed(os.path.join(OUT, "confidence_reduce.npz"), **out)


def synth_alignment(rng, H, W, K, s=0.07, t=0.002):
    """mono prior = (1/d - t)/s quantised /256, sparse target = noisy inverse radar depth."""
    yy, xx = np.mgrid[0:H, 0:W]
    d_true = 20 + 15 * np.sin(xx / 37.0) + 10 * np.cos(yy / 23.0) + rng.uniform(0, 2, (H, W))
    pred = np.round(((1.0 / d_true - t) / s) * 256 * 64) / (256 * 64)
    idx = rng.choice(H * W, K, replace=False)
    mask = np.zeros(H * W, bool)
    mask[idx] = True
    mask = mask.reshape(H, W)
    target = np.zeros((H, W))
    target[mask] = (1.0 / d_true[mask]) * (1 + 0.02 * rng.standard_normal(K))
    return pred.astype(np.float32), target.astype(np.float32), mask


def gen_estimator(ref):
    est = ref["estimator"]
    rng = np.random.default_rng(40)
    out = {}
    for tag, (H, W, K) in {"cfg1": (512, 640, 200), "small": (37, 53, 11), "one": (16, 16, 1), "none": (16, 16, 0)}.items():
        pred, target, mask = synth_alignment(rng, H, W, K)
        s32, t32 = est.compute_scale_and_shift_ls(pred, target, mask)
        s64, t64 = est.compute_scale_and_shift_ls(pred.astype(np.float64), target.astype(np.float64), mask)
        e = est.LeastSquaresEstimator(pred, target, mask)
        e.compute_scale_and_shift()
        e.apply_scale_and_shift()
        raw = e.output.copy()
        e.clamp_min_max(clamp_min=0.1, clamp_max=255.0)
        # the same apply+clamp driven by the float64 solution rounded to float32 (what the kernel does)
        e2 = est.LeastSquaresEstimator(pred, target, mask)
        e2.scale, e2.shift = np.float32(s64), np.float32(t64)
        e2.apply_scale_and_shift()
        e2.clamp_min_max(clamp_min=0.1, clamp_max=255.0)
        items = [("pred", pred), ("target", target), ("mask", mask), ("scale32", s32), ("shift32", t32),
                 ("scale64", s64), ("shift64", t64)]
        if H * W <= 4096:        # keep the fixture small: full maps only for the small cases
            items += [("out_raw32", raw), ("out_clamped32", e.output), ("out_clamped_from64", e2.output)]
        else:
            items += [("out_clamped_from64_sum", np.sum(e2.output, dtype=np.float64)),
                      ("out_clamped_from64_probe", e2.output[::37, ::41].copy())]
        for name, val in items:
            out["%s/%s" % (tag, name)] = np.asarray(val)
        if K >= 2:
            o = est.Optimizer(pred, target, mask, "inv")
            o.optimize_scale()
            out[tag + "/opt_scale"] = np.asarray(o.scale)
    np.savez_compressed(os.path.join(OUT, "estimator.npz"), **out)


def gen_interpolator(ref):
    itp = ref["interpolator"]
    rng = np.random.default_rng(50)
    out = {}
    for tag, (H, W, K) in {"n40": (64, 80, 40), "n3": (24, 32, 3), "n200": (128, 160, 200)}.items():
        pred, target, mask = synth_alignment(rng, H, W, K)
        pred = pred * 0.07 + 0.002          # aligned prior so the ratios are O(1)
        ip = itp.Interpolator2D(pred_inv=pred, sparse_depth_inv=target, valid=mask)
        ip.generate_interpolated_scale_map(interpolate_method="linear", fill_corners=False)
        for name, val in (("pred", pred), ("sparse", target), ("valid", mask), ("knot_coords", ip.knot_coords),
                          ("knot_scales", ip.knot_scales), ("knot_shifts", ip.knot_shifts),
                          ("scale_map", ip.interpolated_scale_map)):
            out["%s/%s" % (tag, name)] = np.asarray(val)
    np.savez_compressed(os.path.join(OUT, "interpolator.npz"), **out)


def gen_rcnet(ref):
    """Tiny RC-Net (same topology, narrow channels) through the reference's real forward_output."""
    rm, main = ref["rcnet_model"], ref["rcnet_main"]
    torch.manual_seed(60)
    cfg = dict(input_channels_image=3, input_channels_depth=3, input_patch_size_image=[64, 32],
               encoder_type=["rcnet", "batch_norm"], n_filters_encoder_image=[8, 16, 16, 16, 16],
               n_neurons_encoder_depth=[8, 16, 16, 16, 16], decoder_type=["multiscale", "batch_norm"],
               n_filters_decoder=[32, 16, 16, 8, 8], device=torch.device("cpu"))
    model = rm.RCNetModel(**cfg)
    model.eval()
    # make BatchNorm running stats non-trivial
    for m in list(model.encoder.modules()) + list(model.decoder.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    H, W, K = 72, 104, 5
    g = torch.Generator().manual_seed(61)
    image = torch.rand(1, 3, H, W, generator=g)
    xs = torch.tensor([0, W - 1, 30, 51, 77]).float()
    ys = torch.tensor([0, H - 1, 20, 36, 9]).float()
    pts = torch.stack([xs + 16, ys + 32, torch.rand(K, generator=g) * 90 + 2], 1)
    boxes = [torch.stack([pts[:, 0] - 16, pts[:, 1] - 32, pts[:, 0] + 16, pts[:, 1] + 32], 1)]
    with torch.no_grad():
        padded = torch.nn.functional.pad(image, (16, 16, 32, 32), mode="replicate")
        logits = model.forward(padded, pts, boxes, return_logits=True)
        thr = float(torch.sigmoid(logits).median())
        d, r = main.forward_output(model, image, pts.clone(), boxes, response_thr=thr, device=torch.device("cpu"))
    out = {"image": image.numpy(), "pts": pts.numpy(), "boxes": boxes[0].numpy(), "logits": logits.numpy(),
           "depth": d.numpy(), "response": r.numpy(), "thr": np.float32(thr)}
    out.update(state_np(model.encoder.state_dict(), "enc/"))
    out.update(state_np(model.decoder.state_dict(), "dec/"))
    np.savez_compressed(os.path.join(OUT, "rcnet_tiny.npz"), **out)


def sml_input_case(tag, H, W):
    """Inputs of the SML-input golden cases, regenerated (not stored) by the tests.  'zju' is a full-size frame whose
    pixels encode their own coordinates (pins cv2's nearest index tables at 640x512 -> 352x288 and compresses well);
    'small' is random data (pins the float32 arithmetic)."""
    if tag == "zju":
        yy, xx = np.mgrid[0:H, 0:W]
        code = ((yy * W + xx) / float(1 << 19)).astype(np.float32)                     # exact in float32
        image = np.stack([code, (xx / 1024.0).astype(np.float32), (yy / 1024.0).astype(np.float32)], -1)
        return image, code + np.float32(0.25), (np.float32(1.0) - code).astype(np.float32)
    rng = np.random.default_rng(H * 1000 + W)
    return (rng.uniform(0, 1, (H, W, 3)).astype(np.float32), rng.uniform(0.004, 1.2, (H, W)).astype(np.float32),
            rng.uniform(0, 1, (H, W)).astype(np.float32))


def gen_sml_glue(ref):
    """N2 / N4 rows and the sml_glue statements: the reference's own transforms, utils and eval functions."""
    import torchvision
    import rcnet_transforms
    from modules.midas import transforms as midas_transforms, utils as midas_utils
    from utils import eval_utils
    rng = np.random.default_rng(77)
    out = {}
    # --- SML input assembly: val_zju.py:181-196 executed literally around the reference's transform pipeline
    for tag, (H, W) in (("zju", (512, 640)), ("small", (96, 200))):
        tf = midas_transforms.get_transforms(288, 288, depth_predictor='midas_small')
        image, int_depth, int_scales = sml_input_case(tag, H, W)
        sample = tf({'image': image.copy(), 'int_depth': int_depth.copy(), 'int_scales': int_scales.copy(),
                     'int_depth_no_tf': int_depth.copy()})
        x = torch.cat([sample['int_depth'], sample['int_scales']], 0)
        image_gray = sample['image'][0] * 0.299 + sample['image'][1] * 0.587 + sample['image'][2] * 0.114
        x = torch.cat([x, image_gray.unsqueeze(0)], 0)
        xs = x.numpy() if tag != "zju" else x.numpy()[:, ::4, ::4]       # full-size case: d in full (index tables), x subsampled
        out.update({"sml_input_%s/shape" % tag: np.array([H, W]), "sml_input_%s/x" % tag: xs,
                    "sml_input_%s/d" % tag: sample['int_depth_no_tf'].numpy()})
    # --- RC-Net prologue: run_rcnet_zju.py:214-240 + Transforms + the pad of rcnet_main.forward_output
    for tag, hi in (("u8", 255.0), ("unit", 1.0)):
        H, W, ph, pw, K = 40, 56, 24, 10, 7
        image = torch.from_numpy(rng.uniform(0, hi, (1, 3, H, W)).astype(np.float32))
        pts0 = np.stack([rng.integers(0, W, K), rng.integers(0, H, K), rng.uniform(1.5, 100, K)], 1).astype(np.float32)
        radar_points = torch.from_numpy(pts0.copy())
        pad_size_x, pad_size_y = pw // 2, ph // 2
        bounding_boxes_list = []
        for i in range(radar_points.shape[0]):
            radar_points[i, 0] = radar_points[i, 0] + pad_size_x
            radar_points[i, 1] = radar_points[i, 1] + pad_size_y
            bounding_box = torch.zeros(4)
            bounding_box[0] = radar_points[i, 0] - pad_size_x
            bounding_box[1] = radar_points[i, 1] - pad_size_y
            bounding_box[2] = radar_points[i, 0] + pad_size_x
            bounding_box[3] = radar_points[i, 1] + pad_size_y
            bounding_boxes_list.append(bounding_box)
        bounding_boxes_list = [torch.stack(bounding_boxes_list, dim=0)]
        tr = rcnet_transforms.Transforms(normalized_image_range=[0, 1])
        [img_t], [pts_t], [bb_t] = tr.transform(images_arr=[image.clone()], points_arr=[radar_points],
                                                bounding_boxes_arr=[bounding_boxes_list], random_transform_probability=0.0)
        padded = torchvision.transforms.functional.pad(img_t, (pad_size_x, pad_size_y, pad_size_x, pad_size_y), padding_mode='edge')
        out.update({"prologue_%s/image" % tag: image.numpy(), "prologue_%s/points" % tag: pts0,
                    "prologue_%s/patch" % tag: np.array([ph, pw]), "prologue_%s/padded" % tag: padded.numpy(),
                    "prologue_%s/points_shifted" % tag: pts_t.numpy(), "prologue_%s/boxes" % tag: bb_t[0].numpy()})
    # --- pins for the statements sml_glue_ref restates: normalize_unit_range, eval_utils
    sc = rng.uniform(0.3, 2.5, (48, 64)).astype(np.float32)
    out["normalize/in"], out["normalize/out"] = sc, midas_utils.normalize_unit_range(sc.copy())
    gt = (rng.uniform(0, 60, (48, 64)) * (rng.uniform(0, 1, (48, 64)) < 0.4)).astype(np.float32)
    pred = (np.abs(gt + rng.normal(0, 2, gt.shape)) + 0.5).astype(np.float32)
    mask = np.where(np.logical_and(gt > 0, np.logical_and(gt > 0.0, gt < 50.0)))       # val_zju.py:212-218
    p, g = pred[mask], gt[mask]
    out["metrics/pred"], out["metrics/gt"] = pred, gt
    out["metrics/values"] = np.array([eval_utils.mean_abs_err(1000.0 * p, 1000.0 * g), eval_utils.root_mean_sq_err(1000.0 * p, 1000.0 * g),
                                      eval_utils.inv_mean_abs_err(0.001 * p, 0.001 * g), eval_utils.inv_root_mean_sq_err(0.001 * p, 0.001 * g),
                                      eval_utils.mean_abs_rel_err(1000.0 * p, 1000.0 * g), eval_utils.mean_sq_rel_err(1000.0 * p, 1000.0 * g),
                                      eval_utils.thr_acc(p, g), float(p.size)], np.float64)
    np.savez_compressed(os.path.join(OUT, "sml_glue.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    only = set(sys.argv[1:])                     # e.g. `python -m oracle.gen_golden gen_sml_glue` regenerates one file
    for fn in (gen_roi_pool, gen_loftr, gen_reduce, gen_estimator, gen_interpolator, gen_rcnet, gen_sml_glue):
        if only and fn.__name__ not in only:
            continue
        fn(ref)
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
