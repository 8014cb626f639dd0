"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference-generated
golden fixtures.  Integer / selection work must be bit-exact; floating point within the
tolerances of BASELINE.json's north_star, written next to each assert."""
import os

import numpy as np
import pytest
import torch

from oracle import (confidence_reduce_ref as cr, estimator_ref as er, linear_attention_ref as lr,
                    roi_pool_ref as rr, rcnet_ref)
from oracle import sml_glue_ref as sg
from riders_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype else t).to(DEV)


# ------------------------------------------------------------------------------------ A1 gather
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_gather_matches_golden_bit_exact(golden, layout):
    g = golden("roi_pool.npz")
    ph, pw = [int(v) for v in g["patch"]]
    feats = [cu(g["feat%d" % l]) for l in range(5)]
    if layout == "nhwc":
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    boxes = [cu(g["boxes0"]), cu(g["boxes1"])]
    for max_roi in (None, (pw, ph), (200, 200)):          # generic, TMA with in-kernel fallback, TMA
        outs = ops.roi_gather(feats, boxes, [float(g["scale%d" % l]) for l in range(5)],
                              [tuple(int(v) for v in g["osize%d" % l]) for l in range(5)], max_roi=max_roi)
        for l, o in enumerate(outs):
            assert o.shape == g["out%d" % l].shape
            assert np.array_equal(o.cpu().numpy(), g["out%d" % l]), "level %d layout %s max_roi %s" % (l, layout, max_roi)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_gather_rcnet_pyramid_vs_torchvision(layout):
    # the real pyramid shapes, K=24 points incl. the four frame corners
    rng = np.random.default_rng(3)
    H, W, patch = 512, 640, (240, 100)
    pts, boxes = synth.radar_points(3, 24, H, W, patch)
    corner = np.array([[0, 0], [W - 1, 0], [0, H - 1], [W - 1, H - 1]], np.float32) + [50, 120]
    boxes[:4] = np.concatenate([corner - [50, 120], corner + [50, 120]], 1)
    shapes = [(32, 376, 370), (64, 188, 185), (128, 94, 93), (128, 47, 47), (128, 24, 24)]
    scales, sizes = rr.rcnet_levels(patch)
    feats_np = [rng.standard_normal((2,) + s).astype(np.float32) for s in shapes]
    blist = [boxes[:14], boxes[14:]]
    rois = rr.boxes_list_to_rois(blist)
    feats = [cu(f) for f in feats_np]
    if layout == "nhwc":
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    outs, args = ops.roi_gather(feats, [cu(b) for b in blist], scales, sizes, want_argmax=True)
    fast = ops.roi_gather(feats, [cu(b) for b in blist], scales, sizes, max_roi=(100, 240))
    for l in range(5):
        ref = rr.roi_pool_torchvision(feats_np[l], rois, scales[l], sizes[l])
        assert np.array_equal(outs[l].cpu().numpy(), ref), "level %d" % l
        assert np.array_equal(fast[l].cpu().numpy(), ref), "level %d (tma)" % l
        a = args[l].cpu().numpy()
        flat = feats_np[l].reshape(2, shapes[l][0], -1)
        b = rois[:, 0].astype(int)
        picked = np.take_along_axis(flat[b], np.maximum(a, 0).reshape(a.shape[0], a.shape[1], -1), axis=2).reshape(a.shape)
        assert np.array_equal(np.where(a >= 0, picked, 0), ref)


def test_gather_backward_matches_torchvision_autograd():
    import torchvision
    rng = np.random.default_rng(5)
    feat = torch.from_numpy(rng.standard_normal((2, 8, 40, 36)).astype(np.float32))
    boxes = [torch.tensor([[2., 3., 30., 60.], [10., 0., 50., 70.]]), torch.tensor([[0., 0., 71., 79.]])]
    f = feat.clone().requires_grad_(True)
    out = torchvision.ops.roi_pool(f, boxes, output_size=(6, 5), spatial_scale=0.5)
    gout = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32))
    out.backward(gout)
    for nhwc in (False, True):
        fd = feat.to(DEV).contiguous(memory_format=torch.channels_last) if nhwc else feat.to(DEV)
        (o,), (a,) = ops.roi_gather([fd], [b.to(DEV) for b in boxes], [0.5], [(6, 5)], want_argmax=True)
        assert np.array_equal(o.cpu().numpy(), out.detach().numpy())
        (gf,) = ops.roi_gather_bwd([gout.to(DEV)], [a], [tuple(feat.shape)], [b.to(DEV) for b in boxes], [0.5], nhwc)
        np.testing.assert_allclose(gf.cpu().numpy(), f.grad.numpy(), rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------ A2-A4 LoFTR
def _layer_params(g, i):
    pre = "layers.%d." % i
    return {k[len(pre):]: cu(v) for k, v in g.items() if k.startswith(pre)}


@pytest.mark.parametrize("tag", ["small", "rcnet"])
def test_linear_attention_matches_reference(golden, tag):
    g = golden("loftr.npz").case(tag)
    out = ops.linear_attention(cu(g["q"]), cu(g["k"]), cu(g["v"]))
    np.testing.assert_allclose(out.cpu().numpy(), g["attn"], rtol=1e-5, atol=1e-6)      # fp32 re-association only
    out = ops.linear_attention(cu(g["q"]), cu(g["k"]), cu(g["v"]), cu(g["qm"]), cu(g["km"]))
    np.testing.assert_allclose(out.cpu().numpy(), g["attn_masked"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["small", "rcnet"])
def test_loftr_simt_matches_reference(golden, tag):
    g = golden("loftr.npz").case(tag)
    n_layers = int(g["n_layers"])
    one = ops.loftr_layer(cu(g["f0"]), cu(g["f1"]), _layer_params(g, 0), impl=ops.LOFTR_SIMT)
    np.testing.assert_allclose(one.cpu().numpy(), g["layer0"], rtol=1e-4, atol=2e-5)
    o0, o1 = ops.loftr_transformer(cu(g["f0"]), cu(g["f1"]), [_layer_params(g, i) for i in range(2 * n_layers)],
                                   ["self", "cross"] * n_layers, impl=ops.LOFTR_SIMT)
    # tolerance: 1e-4 relative (north_star) on the transformer output
    np.testing.assert_allclose(o0.cpu().numpy(), g["o0"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(o1.cpu().numpy(), g["o1"], rtol=1e-4, atol=5e-5)


# ------------------------------------------------------------------------------------ A5 reduce
@pytest.mark.parametrize("tag", ["basic", "single", "allbelow", "dups"])
def test_reduce_matches_golden(golden, tag):
    g = golden("confidence_reduce.npz").case(tag)
    H, W = [int(v) for v in g["hw"]]
    K = g["conf"].shape[0]
    d, r, s = ops.confidence_reduce(cu(g["conf"]), cu(g["pts"]), [0, K], H, W, float(g["thr"]))
    assert np.array_equal(r.cpu().numpy(), g["response"])                     # max: exact
    np.testing.assert_allclose(d.cpu().numpy(), g["depth"], rtol=1e-5, atol=0)  # 1e-4 rel allowed; sums re-associate
    d0, r0, s0 = cr.confidence_reduce_ref(g["conf"][:, 0], g["pts"], H, W, float(g["thr"]))
    margin = cr.top2_margin(g["conf"][:, 0], g["pts"], H, W, float(g["thr"]))
    sel = s.cpu().numpy()[0]
    assert np.array_equal(sel[margin > 1e-6], s0[margin > 1e-6])             # bit-exact where the margin is clear
    assert np.array_equal(sel, s0)                                            # and first-max everywhere else


@pytest.mark.parametrize("K", [128, 4096])
def test_stress_config_1280x1024_reduce_and_gather_vs_oracle(K):
    """BASELINE configs[4] (1280x1024 frames, 128..4096 radar points): the reduce against the oracle (response and select
    exact, depth 1e-5) and the five-level gather against the torchvision rules on the 1280x1024 pyramid (bit-exact), at the
    two ends of the density sweep tools/stress_sweep.py times."""
    H, W, patch = 1024, 1280, (240, 100)
    pts, boxes = synth.radar_points(K, K, H, W, patch)
    rng = np.random.default_rng(K)
    conf = (1.0 / (1.0 + np.exp(-rng.normal(-1.5, 2.0, (K,) + patch)))).astype(np.float32)
    d, r, s = ops.confidence_reduce(cu(conf), cu(pts), [0, K], H, W, 0.1)
    d0, r0, s0 = cr.confidence_reduce_ref(conf, pts, H, W, 0.1)
    assert np.array_equal(r[0].cpu().numpy(), r0) and np.array_equal(s[0].cpu().numpy(), s0)
    np.testing.assert_allclose(d[0].cpu().numpy(), d0, rtol=1e-5, atol=0)
    # gather on the pyramid of the padded 1264 x 1380 frame (ResNet strides 2..32), a subset of the points incl. the corners
    Hp, Wp = H + patch[0], W + patch[1]
    dims, h, w = [], Hp, Wp
    for _ in range(5):
        h, w = (h + 1) // 2, (w + 1) // 2
        dims.append((h, w))
    chans = [32, 64, 128, 128, 128]
    scales, sizes = rr.rcnet_levels(patch)
    sub = np.concatenate([np.arange(min(K, 24)), [K - 1]])
    bx = boxes[sub].copy()
    bx[0] = [0, 0, patch[1], patch[0]]                                   # top-left corner RoI (clipped by the map edge)
    bx[1] = [Wp - patch[1], Hp - patch[0], Wp, Hp]                       # bottom-right corner
    rois = rr.boxes_list_to_rois([bx])
    feats = [rng.standard_normal((1, c, hh, ww)).astype(np.float32) for c, (hh, ww) in zip(chans, dims)]
    got = ops.roi_gather([cu(f).contiguous(memory_format=torch.channels_last) for f in feats], cu(rois), scales, sizes,
                         max_roi=(patch[1], patch[0]))
    for f, sc, sz, g_ in zip(feats, scales, sizes, got):
        assert np.array_equal(g_.cpu().numpy(), rr.roi_pool_torchvision(f, rois, sc, sz))


def test_reduce_full_size_ragged_batch_and_logits():
    H, W, patch = 512, 640, (240, 100)
    Ks = [512, 1, 0, 300]
    pts, conf = [], []
    for f, K in enumerate(Ks):
        p, _ = synth.radar_points(f, max(K, 1), H, W, patch)
        pts.append(p[:K])
        conf.append(synth.confidences(f, max(K, 1), patch)[:K])
    offs = np.concatenate([[0], np.cumsum(Ks)]).tolist()
    P, Cn = np.concatenate(pts), np.concatenate(conf)
    d, r, s = ops.confidence_reduce(cu(Cn), cu(P), offs, H, W, 0.1)
    for f, K in enumerate(Ks):
        d0, r0, s0 = cr.confidence_reduce_ref(conf[f], pts[f], H, W, 0.1)
        assert np.array_equal(r[f].cpu().numpy(), r0)
        np.testing.assert_allclose(d[f].cpu().numpy(), d0, rtol=1e-5, atol=0)
        assert np.array_equal(s[f].cpu().numpy(), s0)
    # logits variant == sigmoid then reduce, within 1e-4 relative on the confidences
    logit = np.log(Cn / (1 - Cn)).astype(np.float32)
    d2, r2, _ = ops.confidence_reduce(cu(logit), cu(P), offs, H, W, 0.1, logits=True)
    sig = torch.sigmoid(cu(logit))
    d3, r3, _ = ops.confidence_reduce(sig, cu(P), offs, H, W, 0.1)
    rr_, r3_ = r2.cpu().numpy(), r3.cpu().numpy()
    same = (rr_ > 0) == (r3_ > 0)
    assert same.mean() > 0.9999
    np.testing.assert_allclose(rr_[same], r3_[same], rtol=1e-4)
    # properties: response is in {0} U [thr, 1]; depth within [min z, max z] where it fired
    rn = r.cpu().numpy()
    assert ((rn == 0) | (rn >= np.float32(0.1))).all()
    dn = d.cpu().numpy()
    assert ((dn[0] == 0) == (rn[0] == 0)).all()
    fired = rn[0] > 0
    assert dn[0][fired].min() >= pts[0][:, 2].min() * (1 - 1e-6) and dn[0][fired].max() <= pts[0][:, 2].max() * (1 + 1e-6)


# ------------------------------------------------------------------------------------ A6/A7 align
@pytest.mark.parametrize("tag", ["cfg1", "small", "one", "none"])
def test_lsq_matches_golden(golden, tag):
    g = golden("estimator.npz").case(tag)
    hi, lo = float(np.float32(1.0 / 0.1)), float(np.float32(1.0 / 255.0))
    sums, ss, out = ops.lsq_align(cu(g["pred"])[None], cu(g["target"])[None], cu(g["mask"])[None], hi, lo)
    ss = ss.cpu().numpy()[0]
    if tag in ("one", "none"):
        assert ss[0] == 0 and ss[1] == 0
    else:
        # 1e-6 relative vs the reference fed float64 inputs (north_star)
        np.testing.assert_allclose(ss[0], g["scale64"], rtol=1e-9)
        np.testing.assert_allclose(ss[1], g["shift64"], rtol=1e-9)
    ref_sums = er.normal_equation_sums(g["pred"], g["target"], g["mask"])
    np.testing.assert_allclose(sums.cpu().numpy()[0], ref_sums, rtol=1e-12, atol=0)
    exp = er.apply_clamp_ref(g["pred"], ss[0], ss[1], 0.1, 255.0)
    assert np.array_equal(out.cpu().numpy()[0], exp)                         # apply+clamp is bit-exact
    if "out_clamped_from64" in g and tag not in ("one", "none"):
        assert np.array_equal(out.cpu().numpy()[0], g["out_clamped_from64"])


def test_lsq_batched_deterministic_and_unaligned():
    B, H, W = 5, 67, 93                                   # odd sizes: scalar tail path
    data = [synth.alignment_inputs(s, 40, H, W) for s in range(B)]
    p, t, m = (np.stack([d[i] for d in data]) for i in range(3))
    a = ops.lsq_align(cu(p), cu(t), cu(m), 10.0, float(np.float32(1 / 255.0)))
    b = ops.lsq_align(cu(p), cu(t), cu(m), 10.0, float(np.float32(1 / 255.0)))
    for x, y in zip(a, b):
        assert torch.equal(x, y)                          # bit-reproducible
    for i in range(B):
        s64, t64 = er.scale_shift_ls_ref(p[i], t[i], m[i])
        np.testing.assert_allclose(a[1][i].cpu().numpy(), [s64, t64], rtol=1e-9)
        assert np.array_equal(a[2][i].cpu().numpy(), er.apply_clamp_ref(p[i], s64, t64, 0.1, 255.0))
    # linearity property: scaling the target scales (scale, shift)
    _, ss2, _ = ops.lsq_align(cu(p), cu(t * np.float32(2)), cu(m), want_out=False)
    np.testing.assert_allclose(ss2.cpu().numpy(), 2 * a[1].cpu().numpy(), rtol=1e-12)


def test_estimator_class_is_a_drop_in(golden):
    from riders_b200.modules.estimator import LeastSquaresEstimator, compute_scale_and_shift_ls
    g = golden("estimator.npz").case("small")
    e = LeastSquaresEstimator(estimate=g["pred"], target=g["target"], valid=g["mask"])
    e.compute_scale_and_shift()
    e.apply_scale_and_shift()
    e.clamp_min_max(clamp_min=0.1, clamp_max=255.0)
    assert isinstance(e.output, np.ndarray) and e.output.dtype == np.float32
    assert e.scale.shape == () and e.scale.dtype == np.float32
    assert np.array_equal(e.output, g["out_clamped_from64"])
    np.testing.assert_allclose(e.scale, g["scale32"], rtol=1e-5)             # vs the as-shipped fp32 reference
    x0, x1 = compute_scale_and_shift_ls(g["pred"], g["target"], g["mask"])
    np.testing.assert_allclose([x0, x1], [g["scale64"], g["shift64"]], rtol=1e-6)
    e2 = LeastSquaresEstimator(estimate=g["pred"], target=g["target"], valid=g["mask"])
    assert np.array_equal(e2.align(clamp_min=0.1, clamp_max=255.0), g["out_clamped_from64"])


# ------------------------------------------------------------------------------------ A0 end to end
def _tiny_model(g, device):
    from riders_b200.rcnet.rcnet_model import RCNetModel
    model = RCNetModel(3, 3, [64, 32], ["rcnet", "batch_norm"], [8, 16, 16, 16, 16], [8, 16, 16, 16, 16],
                       ["multiscale", "batch_norm"], [32, 16, 16, 8, 8], device=torch.device(device))
    model.encoder.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("enc/")})
    model.decoder.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("dec/")})
    model.eval()
    return model


@pytest.mark.parametrize("channels_last", [False, True])
def test_forward_output_matches_reference_golden(golden, channels_last):
    from riders_b200.rcnet.rcnet_main import forward_output, forward_output_batched
    g = golden("rcnet_tiny.npz")
    torch.backends.cudnn.allow_tf32 = False          # isolate the kernels under test from TF32 convs
    torch.backends.cuda.matmul.allow_tf32 = False
    model = _tiny_model(g, DEV)
    if channels_last:
        model.channels_last()
    image, pts, boxes = cu(g["image"]), cu(g["pts"]), [torch.from_numpy(g["boxes"])]   # boxes on the CPU like the reference
    with torch.no_grad():
        padded = torch.nn.functional.pad(image, (16, 16, 32, 32), mode="replicate")
        logits = model.forward(padded, pts, boxes, return_logits=True)
        np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], rtol=1e-4, atol=2e-5)
        d, r, s = forward_output(model, image, pts, boxes, response_thr=float(g["thr"]), return_select=True)
        d2, r2 = forward_output_batched(model, image, pts, [0, pts.shape[0]], float(g["thr"]), chunk_points=2)
    rn, gn = r.cpu().numpy(), g["response"]
    agree = (rn > 0) == (gn > 0)                         # threshold flips only for |conf-thr| < 1e-4*thr
    assert agree.mean() > 0.995
    np.testing.assert_allclose(rn[agree], gn[agree], rtol=1e-4)              # confidences: 1e-4 relative
    both = agree & (gn > 0)
    np.testing.assert_allclose(d.cpu().numpy()[both], g["depth"][both], rtol=2e-3)
    assert torch.allclose(r2, r, rtol=1e-4, atol=1e-6) and torch.allclose(d2, d, rtol=1e-3, atol=1e-5)


# ------------------------------------------------------------------------------------ A8 interpolation
def _interp_check(pred, sparse, valid, expected=None):
    from oracle import interpolator_ref as ir
    H, W = pred.shape
    n = int(valid.sum())
    xy, sc, sh, cnt = ops.extract_knots(cu(pred)[None], cu(sparse)[None], cu(valid)[None], max(n, 1))
    coords, scales, shifts = ir.extract_knots_ref(pred, sparse, valid)
    assert int(cnt[0]) == n
    assert np.array_equal(xy[0, :n].cpu().numpy(), coords.T)                 # row-major scan order, (x, y)
    assert np.array_equal(sc[0, :n].cpu().numpy(), scales) and np.array_equal(sh[0, :n].cpu().numpy(), shifts)
    out, status = ops.interp_linear(xy, sc, cnt, H, W, 1.0)
    assert int(status[0]) == 0
    got = out[0].cpu().numpy()
    ref = expected if expected is not None else ir.interpolated_scale_map_ref(pred, sparse, valid)
    ref64 = ir.interpolate_knots_ref((H, W), coords, scales)
    outside = np.isclose(ref64, 1.0, rtol=0, atol=0) & (ref == 1.0)
    degenerate = ir.degenerate_pixel_mask(coords.T, H, W)
    ok = ~degenerate
    # exactly fill outside the hull (hull membership is exact: integer predicates)
    hull_out = ok & (got == 1.0) & (ref == 1.0)
    inside = ok & ~hull_out
    np.testing.assert_allclose(got[inside], ref[inside], rtol=1e-5, atol=0)  # 1e-5 relative inside the hull
    assert np.array_equal(got[ok] == 1.0, ref[ok] == 1.0)                     # strict: exactly the fill value outside, and only there
    # knots reproduce their values exactly
    assert np.array_equal(got[coords[1], coords[0]], scales.astype(np.float32))
    return degenerate.mean()


@pytest.mark.parametrize("tag", ["n40", "n3", "n200"])
def test_interp_matches_reference_golden(golden, tag):
    g = golden("interpolator.npz").case(tag)
    frac = _interp_check(g["pred"], g["sparse"], g["valid"], g["scale_map"])
    assert frac < 0.05


@pytest.mark.parametrize("K,H,W", [(200, 512, 640), (512, 512, 640), (4096, 1024, 1280)])
def test_interp_full_size_vs_scipy(K, H, W):
    pred, target, mask = synth.alignment_inputs(11, K, H, W)
    pred = (pred * np.float32(0.07) + np.float32(0.002)).astype(np.float32)
    frac = _interp_check(pred, target, mask)
    assert frac < 0.02                       # co-circular quads are rare and masked


def test_interp_triangulation_is_delaunay_and_batched():
    from scipy.spatial import Delaunay
    B, K, H, W = 6, 300, 512, 640
    xs = []
    for b in range(B):
        rng = np.random.default_rng(70 + b)
        flat = np.sort(rng.choice(H * W, K - 7 * b, replace=False))
        xs.append(np.stack([flat % W, flat // W], 1).astype(np.int32))
    xy = np.zeros((B, K, 2), np.int32)
    val = np.zeros((B, K), np.float32)
    for b in range(B):
        xy[b, :len(xs[b])] = xs[b]
        val[b, :len(xs[b])] = np.random.default_rng(b).uniform(0.5, 2.0, len(xs[b]))
    cnt = torch.tensor([len(x) for x in xs], dtype=torch.int32)
    out, status, tri, tcount = ops.interp_linear(cu(xy), cu(val), cnt.to(DEV), H, W, 1.0, want_triangles=True)
    assert not status.cpu().numpy().any()
    for b in range(B):
        mine = set(tuple(sorted(int(v) for v in t)) for t in tri[b, :int(tcount[b])].cpu().numpy())
        theirs = set(tuple(sorted(int(v) for v in t)) for t in Delaunay(xs[b].astype(np.float64)).simplices)
        assert len(mine) == len(theirs) and len(mine ^ theirs) <= 6
    # frames are independent: frame b alone gives the same map bit for bit
    solo, _ = ops.interp_linear(cu(xy[2:3]), cu(val[2:3]), cnt[2:3].to(DEV), H, W, 1.0)
    assert torch.equal(solo[0], out[2])
    again, _ = ops.interp_linear(cu(xy), cu(val), cnt.to(DEV), H, W, 1.0)
    assert torch.equal(again, out)           # deterministic despite the atomics inside


def test_interpolator_class_and_qhull_errors(golden):
    from riders_b200.modules.interpolator import Interpolator2D, QhullError
    g = golden("interpolator.npz").case("n40")
    ip = Interpolator2D(pred_inv=g["pred"], sparse_depth_inv=g["sparse"], valid=g["valid"])
    assert ip.num_knots == 40 and np.array_equal(ip.knot_coords, g["knot_coords"])
    assert np.array_equal(ip.knot_scales, g["knot_scales"]) and np.array_equal(ip.knot_shifts, g["knot_shifts"])
    ip.generate_interpolated_scale_map(interpolate_method="linear", fill_corners=False)
    m = ip.interpolated_scale_map
    assert m.dtype == np.float32 and m.shape == g["scale_map"].shape
    diff = np.abs(m - g["scale_map"]) / np.abs(g["scale_map"])
    assert np.mean(diff > 1e-5) < 0.05
    # 2 knots / collinear knots: the reference's SciPy call raises QhullError (SURVEY 8a edge cases)
    valid = np.zeros((16, 16), bool)
    valid[3, 4] = valid[9, 2] = True
    with pytest.raises(QhullError):
        Interpolator2D(np.ones((16, 16), np.float32), np.ones((16, 16), np.float32), valid).generate_interpolated_scale_map("linear")
    valid[:] = False
    valid[5, [1, 4, 8, 12]] = True
    with pytest.raises(QhullError):
        Interpolator2D(np.ones((16, 16), np.float32), np.ones((16, 16), np.float32), valid).generate_interpolated_scale_map("linear")


def test_lsq_launch_forms_are_bit_identical():
    """riders_lsq_align has four launch forms - two plain kernels, one cooperative launch (partials -> grid barrier ->
    solve + apply), ONE ordinary launch with a per-frame flag barrier (nonce-tagged flags, no zeroed scratch), and the
    default: pass 2 as a programmatic dependent launch that prefetches `pred` under pass 1.  Same slices and summation
    order: bit-identical outputs."""
    B, Hh, Ww = 5, 512, 640
    ins = [synth.alignment_inputs(20 + b, 300 + 40 * b, Hh, Ww) for b in range(B)]
    p, t, m = (cu(np.stack([x[i] for x in ins])) for i in range(3))
    hi, lo = 10.0, float(np.float32(1 / 255.0))
    res, launches = {}, {}
    try:
        for mode in (0, 1, 2, 3):
            ops.debug_set("lsq_fused", mode)
            n0 = ops.launch_count()
            res[mode] = ops.lsq_align(p, t, m, clamp_hi=hi, clamp_lo=lo)
            launches[mode] = ops.launch_count() - n0
            res[(mode, "noout")] = ops.lsq_align(p, t, m, want_out=False)[1]
            res[(mode, "odd")] = ops.lsq_align(p[:, 1:, 3:], t[:, 1:, 3:], m[:, 1:, 3:], clamp_hi=hi, clamp_lo=lo)    # unaligned
            if mode == 3:                                                  # same workspace pattern twice in a row: the nonce changes
                again = ops.lsq_align(p * 1.5, t, m, clamp_hi=hi, clamp_lo=lo)
                ops.debug_set("lsq_fused", 0)
                for a, b in zip(again, ops.lsq_align(p * 1.5, t, m, clamp_hi=hi, clamp_lo=lo)):
                    assert torch.equal(a, b)
    finally:
        ops.debug_set("lsq_fused", 2)
    assert launches == {0: 2, 1: 1, 2: 2, 3: 1}
    for mode in (1, 2, 3):
        for a, b in zip(res[mode], res[0]):
            assert torch.equal(a, b)
        assert torch.equal(res[(mode, "noout")], res[0][1])
        for a, b in zip(res[(mode, "odd")], res[(0, "odd")]):
            assert torch.equal(a, b)


def test_interp_nearest_matches_griddata_where_the_nearest_knot_is_unique():
    """Interpolator2D.generate_interpolated_scale_map('nearest') (griddata nearest = cKDTree query): identical wherever one
    knot is strictly nearest; among equidistant knots SciPy's pick is undefined, so those pixels are only checked to carry
    the value of ONE of the nearest knots."""
    from scipy.interpolate import griddata
    from riders_b200.modules.interpolator import Interpolator2D
    H, W, K = 120, 160, 60
    pred, target, mask = synth.alignment_inputs(5, K, H, W)
    pred = (pred * np.float32(0.07) + np.float32(0.002)).astype(np.float32)
    it = Interpolator2D(pred, target, mask)
    it.generate_interpolated_scale_map('nearest')
    got = it.interpolated_scale_map
    ys, xs = np.nonzero(mask)
    vals = (target[mask] / pred[mask]).astype(np.float32)
    rows, cols = np.mgrid[0:H, 0:W]
    ref = griddata(np.stack([xs, ys], 1), vals, (cols, rows), method='nearest').astype(np.float32)
    d2 = (cols[..., None] - xs) ** 2 + (rows[..., None] - ys) ** 2
    dmin = d2.min(-1)
    unique = (d2 == dmin[..., None]).sum(-1) == 1
    assert unique.mean() > 0.9 and np.array_equal(got[unique], ref[unique])
    tie_ok = (np.abs(got[..., None] - vals) < 1e-12) & (d2 == dmin[..., None])
    assert tie_ok.any(-1).all()
    with pytest.raises(NotImplementedError):
        it.generate_interpolated_scale_map('cubic')


def test_ransac_estimator_follows_the_reference_draw_for_draw():
    """LeastSquaresEstimator.compute_scale_and_shift_ran (estimator.py:32-87,101-106): with the same np.random seed the mirror
    consumes the same draws and returns the reference's scale / shift (restated here on the full masked maps, as shipped)."""
    from riders_b200.modules.estimator import LeastSquaresEstimator
    pred, target, mask = synth.alignment_inputs(9, 200)
    pred = pred.astype(np.float32)

    def reference(num_iterations=60, sample_size=5, thr=0.02, ratio=0.8):
        best = (0.0, 0.0, 0)
        vi = np.where(mask)
        n = len(vi[0])
        for _ in range(num_iterations):
            idx = np.random.choice(n, size=sample_size, replace=False)
            ms = np.zeros_like(mask)
            ms[vi[0][idx], vi[1][idx]] = 1
            a00, a01, a11 = np.sum(ms * pred * pred), np.sum(ms * pred), np.sum(ms)
            b0, b1 = np.sum(ms * pred * target), np.sum(ms * target)
            det = a00 * a11 - a01 * a01
            x0 = (a11 * b0 - a01 * b1) / det if det > 0 else 0.0
            x1 = (-a01 * b0 + a00 * b1) / det if det > 0 else 0.0
            res = np.abs(mask * pred * x0 + x1 - mask * target)[mask]
            cnt = np.sum(res < thr)
            if cnt > best[2]:
                best = (x0, x1, cnt)
                if cnt / n > ratio:
                    break
        return best

    np.random.seed(4)
    r0, r1, _ = reference()
    np.random.seed(4)
    est = LeastSquaresEstimator(pred, target, mask)
    est.compute_scale_and_shift_ran()
    np.testing.assert_allclose([est.scale, est.shift], [r0, r1], rtol=1e-4)
    est.apply_scale_and_shift()
    np.testing.assert_allclose(est.output, pred * np.float32(est.scale) + np.float32(est.shift), rtol=1e-6)


# ------------------------------------------------------------------------------------ whole chain
def _rand_model(device, patch, filt, dec, seed=0):
    from riders_b200.rcnet.rcnet_model import RCNetModel
    torch.manual_seed(seed)
    m = RCNetModel(3, 3, list(patch), ["rcnet", "batch_norm"], filt, filt, ["multiscale", "batch_norm"], dec,
                   device=torch.device("cpu"))
    m.eval()
    return m


def _to_cuda(m):
    import copy
    g = copy.copy(m)
    g.encoder, g.decoder = copy.deepcopy(m.encoder).to(DEV), copy.deepcopy(m.decoder).to(DEV)
    g.device = torch.device(DEV)
    return g


def test_pipeline_matches_oracle_chain():
    """RadarDepthPipeline (device + host entry points) against oracle.chain_ref frame by frame."""
    from oracle import chain_ref, interpolator_ref as ir
    from riders_b200.pipeline import RadarDepthPipeline
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    Hh, Ww, patch, K, B = 96, 128, (64, 32), 14, 3
    cpu = _rand_model("cpu", patch, [8, 16, 16, 16, 16], [32, 16, 16, 8, 8])
    gpu = _to_cuda(cpu).channels_last()
    batch = synth.chain_batch(40, B, K, Hh, Ww, patch)
    pipe = RadarDepthPipeline(gpu, response_thr=0.45, chunk_points=16,
                              host_keys=("depth", "int_depth", "scale_map", "int_scales", "scale_shift", "interp_status"))
    out = pipe.run_device(cu(batch["images"]), cu(batch["points"]), batch["offsets"], cu(batch["mono"]), cu(batch["sparse"]))
    hout, h2d, d2h = pipe.run_host(batch["images"], batch["points"], batch["offsets"], batch["mono"], batch["sparse"])
    torch.cuda.synchronize()
    assert h2d == sum(batch[k].nbytes for k in ("images", "points", "mono", "sparse")) and d2h > 3 * B * Hh * Ww * 4
    for k in ("depth", "int_depth", "scale_map"):
        assert np.array_equal(hout[k].numpy(), out[k].cpu().numpy())          # host entry point == device entry point
    # pipelined host entry point: two steps in flight (copies on their own streams, alternating pinned result buffers)
    batch2 = synth.chain_batch(77, B, K, Hh, Ww, patch)
    ref2 = pipe.run_device(cu(batch2["images"]), cu(batch2["points"]), batch2["offsets"], cu(batch2["mono"]), cu(batch2["sparse"]))
    h1 = pipe.submit_host(batch["images"], batch["points"], batch["offsets"], batch["mono"], batch["sparse"])
    h2 = pipe.submit_host(batch2["images"], batch2["points"], batch2["offsets"], batch2["mono"], batch2["sparse"])
    (o1, b_in, b_out), (o2, _, _) = h1.wait(), h2.wait()
    assert (b_in, b_out) == (h2d, d2h)
    for k in ("depth", "int_depth", "scale_map", "int_scales"):
        assert np.array_equal(o1[k].numpy(), out[k].cpu().numpy()) and np.array_equal(o2[k].numpy(), ref2[k].cpu().numpy())
    for b in range(B):
        fr = synth.chain_frame(40 + b, K, Hh, Ww, patch)
        ref = chain_ref.chain_frame_ref(cpu, fr, 0.45)
        r, d = out["response"][b].cpu().numpy(), out["depth"][b].cpu().numpy()
        agree = (r > 0) == (ref["response"] > 0)
        assert agree.mean() > 0.99                                             # flips only for |conf-thr| < 1e-4
        np.testing.assert_allclose(r[agree], ref["response"][agree], rtol=1e-4)   # confidences 1e-4 relative
        both = agree & (ref["response"] > 0)
        np.testing.assert_allclose(d[both], ref["depth"][both], rtol=2e-3)
        ss = out["scale_shift"][b].cpu().numpy()
        np.testing.assert_allclose(ss, [ref["scale"], ref["shift"]], rtol=1e-9)  # alignment 1e-6 relative (fp64 both sides)
        assert np.array_equal(out["int_depth"][b].cpu().numpy(), ref["int_depth"])
        coords, _, _ = ir.extract_knots_ref(ref["int_depth"], ref["sparse_inv"], ref["valid"])
        ok = ~ir.degenerate_pixel_mask(coords.T, Hh, Ww)
        sm = out["scale_map"][b].cpu().numpy()
        assert np.array_equal(sm[ok] == 1.0, ref["scale_map"][ok] == 1.0)      # exactly fill outside the hull
        np.testing.assert_allclose(sm[ok], ref["scale_map"][ok], rtol=1e-5)     # 1e-5 relative inside
        from oracle import sml_glue_ref as gr                                   # scaffold of the GPU's own quasi-dense depth
        assert np.array_equal(out["int_scales"][b].cpu().numpy(),
                              gr.scale_scaffold_ref(ref["int_depth"], d, ref["sparse_inv"], ref["valid"]))


@pytest.mark.parametrize("impl,optimized", [("simt", False), ("tcgen05", False), ("tcgen05", True)])
def test_full_size_rcnet_forward_vs_oracle(impl, optimized, monkeypatch):
    """The real RC-Net configuration (d_model 128, 21 tokens, patch 240x100, 752x740 padded frame):
    gather on the real pyramid through the TMA path + the LoFTR kernels, against the CPU oracle."""
    if impl == "tcgen05" and not ops.loftr_impl_available(ops.LOFTR_TCGEN05, 128, 8, 21):
        pytest.skip("tcgen05 LoFTR layer not built into this library")
    monkeypatch.setenv("RIDERS_LOFTR_IMPL", impl)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    patch, K = (240, 100), 9
    cpu = _rand_model("cpu", patch, [32, 64, 128, 128, 128], [256, 128, 64, 32, 16], seed=1)
    gpu = _to_cuda(cpu)
    gpu = gpu.optimize_for_inference() if optimized else gpu.channels_last()   # optimized: folded BN + fused decoder glue + head
    fr = synth.chain_frame(7, K, 512, 640, patch)
    fr["points"][0, :2] = [50, 120]                       # a corner point: RoI clipped by the map edge
    fr["boxes"][0] = [0, 0, 100, 240]
    image = torch.from_numpy(fr["image"])[None]
    padded = torch.nn.functional.pad(image, (50, 50, 120, 120), mode="replicate")
    with torch.no_grad():
        ref = rcnet_ref.model_forward_ref(cpu, padded, torch.from_numpy(fr["points"]), [torch.from_numpy(fr["boxes"])])
        got = gpu.forward(padded.to(DEV), cu(fr["points"]), [cu(fr["boxes"])], return_logits=True)
    # logits -> confidences within 1e-4 relative (north_star)
    np.testing.assert_allclose(torch.sigmoid(got).cpu().numpy(), torch.sigmoid(ref).numpy(), rtol=1e-4, atol=1e-6)


# Tolerance on the confidences in the TF32 flow (cuDNN TF32 convolutions + TF32 fused decoder tail: torch's and the
# reference's GPU default).  TF32 keeps 10 mantissa bits, so 1e-4 is out of reach by construction; the bench's `parity`
# block reports the measured error of every run (profiles/r02_bench_*.json: ~1e-3), this is the ceiling the test enforces.
TF32_CONF_RTOL = 5e-3


@pytest.mark.parametrize("mode,align", [("strict_fp32", "st"), ("tf32", "st"), ("strict_fp32", "s")])
def test_benched_configuration_k512_vs_oracle_chain(mode, align):
    """The bench workload itself - one 640x512 frame, 512 radar points, optimize_for_inference(), raw inputs, retry,
    SML input - through RadarDepthPipeline against oracle.chain_ref, in BOTH arithmetic modes bench.py can time:
    strict fp32 (north_star tolerances) and the TF32 flow (TF32 tolerance on the confidences, everything downstream of
    the alignment unchanged because that half never sees a TF32 operation)."""
    import bench
    from oracle import chain_ref, confidence_reduce_ref as cr, interpolator_ref as ir
    from riders_b200.pipeline import RadarDepthPipeline
    tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    K, Hh, Ww, patch, thr = 512, 512, 640, (240, 100), 0.1
    cpu = bench.build_model(torch.device("cpu"))
    gpu = bench.build_model(torch.device(DEV))
    fr = synth.chain_frame(0, K, Hh, Ww, patch)
    ref = chain_ref.chain_frame_ref(cpu, fr, thr, want_conf=True, global_alignment=align)
    images = cu(np.rint(fr["image"] * 255.0).astype(np.float32))[None]
    points = cu((fr["points"] - np.array([patch[1] // 2, patch[0] // 2, 0], np.float32)).astype(np.float32))
    pipe = RadarDepthPipeline(gpu, response_thr=thr, chunk_points=2048, raw_inputs=True, retry_empty=True, sml_input=True,
                              global_alignment=align, return_select=True)
    out = pipe.run_device(images, points, [0, K], cu(fr["mono"])[None], cu(fr["sparse"])[None])
    assert gpu.decoder.last_head_sigmoid and (getattr(gpu.decoder, "last_used_tail", tf32) == tf32)
    r, d, sel = out["response"][0].cpu().numpy(), out["depth"][0].cpu().numpy(), out["select"][0].cpu().numpy()
    fired, fired_ref = r > 0, ref["response"] > 0
    assert (fired == fired_ref).mean() > 0.999
    both = fired & fired_ref
    rtol = TF32_CONF_RTOL if tf32 else 1e-4
    np.testing.assert_allclose(r[both], ref["response"][both], rtol=rtol)
    if not tf32:
        margin = cr.top2_margin(ref["conf"], fr["points"], Hh, Ww, thr)
        m = (margin > 1e-6) & both
        # north_star: select bit-exact where the margin exceeds 1e-6 - read as 1e-6 beyond the 1e-4 confidence tolerance
        # (a GPU confidence may legitimately differ from the CPU's by 1e-4 relative, i.e. more than such a margin)
        safe = m & (margin > 2e-4)
        assert np.array_equal(sel[safe], ref["select"][safe])
        assert (sel[m] == ref["select"][m]).mean() > 0.99
    np.testing.assert_allclose(d[both], ref["depth"][both], rtol=20 * rtol)
    ss = out["scale_shift"][0].cpu().numpy()
    np.testing.assert_allclose(ss[0], ref["scale"], rtol=1e-6)
    if align == "st":
        np.testing.assert_allclose(ss[1], ref["shift"], rtol=1e-6)
        assert np.array_equal(out["int_depth"][0].cpu().numpy(), ref["int_depth"])
    else:
        np.testing.assert_allclose(out["int_depth"][0].cpu().numpy(), ref["int_depth"], rtol=2e-6)
    coords, _, _ = ir.extract_knots_ref(ref["int_depth"], ref["sparse_inv"], ref["valid"])
    ok = ~ir.degenerate_pixel_mask(coords.T, Hh, Ww)
    sm = out["scale_map"][0].cpu().numpy()
    assert np.array_equal(sm[ok] == 1.0, ref["scale_map"][ok] == 1.0)
    np.testing.assert_allclose(sm[ok], ref["scale_map"][ok], rtol=1e-5 if align == "st" else 2e-5)


# ------------------------------------------------------------------------------------ N1 L1 scale (Optimizer)
@pytest.mark.parametrize("depth_type,s_true", [("inv", 0.07), ("inv", 0.2), ("pos", 1.1)])
def test_l1_scale_matches_scipy_bounded_brent(depth_type, s_true):
    from scipy.optimize import minimize_scalar
    B, H, W, K = 3, 120, 160, 200
    bounds = (0.01, 0.3) if depth_type == "inv" else (0.5, 1.6)
    P, T, M = [], [], []
    for b in range(B):
        pred, target, mask = synth.alignment_inputs(20 + b, K, H, W, s=s_true, t=0.0)
        P.append(pred); T.append(target); M.append(mask)
    P, T, M = np.stack(P), np.stack(T), np.stack(M)
    scale, fval, nfev, out = ops.l1_scale_align(cu(P), cu(T), cu(M), bounds=bounds,
                                                clamp_hi=10.0, clamp_lo=float(np.float32(1 / 255.0)))
    for b in range(B):
        p64, t64, m = P[b].astype(np.float64), T[b].astype(np.float64), M[b]
        res = minimize_scalar(er.l1_objective_ref, args=(p64, t64, m), bounds=bounds)       # the reference's call, fp64-fed
        # same algorithm, same fp64 objective up to summation order: 1e-6 relative on the scale (north_star)
        np.testing.assert_allclose(scale[b].item(), res.x, rtol=1e-6)
        assert abs(int(nfev[b]) - res.nfev) <= 2
        np.testing.assert_allclose(fval[b].item(), res.fun, rtol=1e-9)
        exp = P[b] * np.float32(scale[b].item())
        exp = np.clip(exp, np.float32(1 / 255.0), np.float32(10.0))
        assert np.array_equal(out[b].cpu().numpy(), exp)
        res32 = minimize_scalar(er.l1_objective_ref, args=(P[b], T[b], M[b]), bounds=bounds)  # as shipped (float32 maps)
        assert abs(scale[b].item() - res32.x) < 2e-5                                         # within Brent's own xatol


def test_optimizer_class_is_a_drop_in():
    from riders_b200.modules.estimator import Optimizer
    pred, target, mask = synth.alignment_inputs(5, 300, 96, 128, s=0.05, t=0.0)
    o = Optimizer(estimate=pred, target=target, valid=mask, depth_type="inv")
    o.optimize_scale()
    assert isinstance(o.scale, np.float64) and 0.01 <= o.scale <= 0.3
    o.apply_scale()
    o.clamp_min_max(clamp_min=0.1, clamp_max=255.0)
    assert isinstance(o.output, np.ndarray) and o.output.dtype == np.float32 and o.output.shape == pred.shape
    assert o.output.max() <= np.float32(10.0) and o.output.min() >= np.float32(1 / 255.0)
    o2 = Optimizer(estimate=pred, target=target, valid=mask, depth_type="inv")
    assert np.array_equal(o2.align(clamp_min=0.1, clamp_max=255.0), o.output) and o2.scale == o.scale


# ------------------------------------------------------------------------------------ decoder glue
@pytest.mark.parametrize("shape,size", [((3, 32, 120, 50), (240, 100)), ((2, 256, 7, 3), (15, 6)), ((2, 128, 30, 12), (60, 25)),
                                        ((1, 8, 13, 5), (5, 13)), ((2, 4, 2, 1), (7, 9))])
def test_upsample_nearest_is_bit_exact_with_torch(shape, size):
    x = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    got = ops.upsample_nearest(x, size)
    ref = torch.nn.functional.interpolate(x, size=size)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, ref)


def test_optimize_for_inference_keeps_the_tolerance():
    """BatchNorm folding + channels_last + the upsample kernel: confidences stay within 1e-4 of the oracle."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    patch, K = (64, 32), 10
    cpu = _rand_model("cpu", patch, [8, 16, 16, 16, 16], [32, 16, 16, 8, 8], seed=3)
    with torch.no_grad():                                  # non-trivial running statistics so that folding matters
        for net in (cpu.encoder, cpu.decoder):
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.uniform_(-0.3, 0.3)
                    m.running_var.uniform_(0.5, 1.5)
                    m.weight.uniform_(0.7, 1.3)
                    m.bias.uniform_(-0.2, 0.2)
    gpu = _to_cuda(cpu).optimize_for_inference()
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in gpu.decoder.modules())
    fr = synth.chain_frame(9, K, 96, 128, patch)
    padded = torch.nn.functional.pad(torch.from_numpy(fr["image"])[None], (16, 16, 32, 32), mode="replicate")
    with torch.no_grad():
        ref = rcnet_ref.model_forward_ref(cpu, padded, torch.from_numpy(fr["points"]), [torch.from_numpy(fr["boxes"])])
        got = gpu.forward(padded.to(DEV), cu(fr["points"]), [cu(fr["boxes"])], return_logits=True)
    np.testing.assert_allclose(torch.sigmoid(got).cpu().numpy(), torch.sigmoid(ref).numpy(), rtol=1e-4, atol=1e-6)


def test_leaky_concat_is_bit_exact_with_torch():
    a = torch.randn((5, 32, 30, 12), device=DEV).contiguous(memory_format=torch.channels_last)
    b = torch.randn((5, 64, 30, 12), device=DEV).contiguous(memory_format=torch.channels_last)
    got = ops.leaky_concat(a, b, 0.2)
    ref = torch.cat([torch.nn.functional.leaky_relu(a, 0.2), b], dim=1)
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, ref)
    assert torch.equal(ops.leaky_concat(a, b, 0.2, act_on_a=False), torch.cat([a, b], dim=1))


def test_bias_act_and_biased_concat_match_torch():
    x = torch.randn((3, 16, 20, 10), device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(16, device=DEV)
    ref = torch.nn.functional.leaky_relu(x + bias.view(1, -1, 1, 1), 0.2)
    got = ops.bias_act_(x.clone(memory_format=torch.channels_last), bias, 0.2)
    assert torch.equal(got, ref)
    assert torch.equal(ops.bias_act_(x.clone(memory_format=torch.channels_last), None, 1.0), x)
    assert torch.equal(ops.bias_act_(x.clone(memory_format=torch.channels_last), bias, 0.0), torch.relu(x + bias.view(1, -1, 1, 1)))
    one = torch.randn((2, 1, 9, 7), device=DEV)                               # C == 1 (output head): scalar path
    assert torch.equal(ops.bias_act_(one.clone(), torch.tensor([0.5], device=DEV), 0.2),
                       torch.nn.functional.leaky_relu(one + 0.5, 0.2))
    skip = torch.randn((3, 32, 20, 10), device=DEV).contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.leaky_concat(x, skip, 0.2, bias=bias), torch.cat([ref, skip], dim=1))


def test_conv3x3_head_matches_torch():
    torch.backends.cudnn.allow_tf32 = False
    x = torch.randn((5, 16, 37, 53), device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(16, device=DEV)
    w = torch.randn((1, 16, 3, 3), device=DEV) * 0.2
    got = ops.conv3x3_head(x, bias, 0.2, w)
    ref = torch.nn.functional.conv2d(torch.nn.functional.leaky_relu(x + bias.view(1, -1, 1, 1), 0.2), w, padding=1)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-5)
    got2 = ops.conv3x3_head(x, None, 1.0, w)                                   # no bias, identity activation
    ref2 = torch.nn.functional.conv2d(x, w, padding=1)
    np.testing.assert_allclose(got2.cpu().numpy(), ref2.cpu().numpy(), rtol=1e-5, atol=1e-5)
    big = torch.randn((3, 16, 240, 100), device=DEV).contiguous(memory_format=torch.channels_last)   # the RC-Net patch
    np.testing.assert_allclose(ops.conv3x3_head(big, bias, 0.2, w).cpu().numpy(),
                               torch.nn.functional.conv2d(torch.nn.functional.leaky_relu(big + bias.view(1, -1, 1, 1), 0.2), w, padding=1).cpu().numpy(),
                               rtol=1e-5, atol=1e-5)
    # fused sigmoid (what forward_output_batched feeds the reduce kernel): torch.sigmoid of the SAME logits, 1e-6
    logit = ops.conv3x3_head(big, bias, 0.2, w)
    np.testing.assert_allclose(ops.conv3x3_head(big, bias, 0.2, w, sigmoid=True).cpu().numpy(), torch.sigmoid(logit).cpu().numpy(),
                               rtol=2e-6, atol=1e-7)


# ------------------------------------------------------------------------------------ N2/N3 scaffold, N4 metrics
def test_scale_scaffold_is_bit_exact_with_the_reference_statements():
    from oracle import sml_glue_ref as gr
    B, H, W = 3, 96, 128
    ID, RC, SI, SV = [], [], [], []
    for b in range(B):
        rng = np.random.default_rng(300 + b)
        pred, target, mask = synth.alignment_inputs(30 + b, 60, H, W)
        int_depth = er.apply_clamp_ref(pred, 0.07, 0.002, 0.1, 255.0)
        rc = rng.uniform(0.0, 120.0, (H, W)).astype(np.float32)              # some beyond max_depth, dense like A5's output
        rc[rng.uniform(size=(H, W)) < 0.3] = 0.0                               # uncovered pixels
        ID.append(int_depth); RC.append(rc); SI.append(target); SV.append(mask)
    ID, RC, SI, SV = (np.stack(a) for a in (ID, RC, SI, SV))
    out, minmax, status = ops.scale_scaffold(cu(ID), cu(RC), cu(SI), cu(SV))
    assert not status.cpu().numpy().any()
    for b in range(B):
        ref = gr.scale_scaffold_ref(ID[b], RC[b], SI[b], SV[b])
        assert np.array_equal(out[b].cpu().numpy(), ref)                       # float32 statements in the same order
    assert np.array_equal(gr.png_roundtrip_ref(RC[0]), np.floor(RC[0] * 256) / 256)
    # no RC-Net map (interp without 'rcnet', val_zju.py:175-178) and the zero-range error case
    out2, _, _ = ops.scale_scaffold(cu(ID), None, cu(SI), cu(SV))
    assert np.array_equal(out2[1].cpu().numpy(), gr.scale_scaffold_ref(ID[1], None, SI[1], SV[1]))
    _, _, st = ops.scale_scaffold(cu(ID), None, cu(SI), cu(np.zeros_like(SV)))
    assert st.cpu().numpy().all()                                              # all ones -> the reference raises ValueError


def test_depth_metrics_match_eval_utils():
    from oracle import sml_glue_ref as gr
    B, H, W = 4, 120, 160
    rng = np.random.default_rng(9)
    gt = rng.uniform(0.5, 80.0, (B, H, W)).astype(np.float32)
    gt[rng.uniform(size=gt.shape) < 0.9] = 0.0                                  # sparse lidar ground truth
    pred = (gt * rng.uniform(0.8, 1.3, gt.shape) + rng.uniform(0.1, 1.0, gt.shape)).astype(np.float32)
    pred[gt == 0] = rng.uniform(1, 50, (gt == 0).sum()).astype(np.float32)
    got = ops.depth_metrics(cu(pred), cu(gt), 0.0, 50.0).cpu().numpy()
    for b in range(B):
        ref = gr.depth_metrics_ref(pred[b], gt[b], 0.0, 50.0)
        assert got[b, 7] == ref[7]
        np.testing.assert_allclose(got[b, :7], ref[:7], rtol=2e-6)              # float32 pairwise mean vs fp64 tree
    again = ops.depth_metrics(cu(pred), cu(gt), 0.0, 50.0)
    assert torch.equal(again, ops.depth_metrics(cu(pred), cu(gt), 0.0, 50.0))   # deterministic


def test_d2s_bias_act_concat_matches_torch():
    N, Ca, Cb, h, w = 3, 8, 12, 5, 7
    y4 = torch.randn((N, 4 * Ca, h, w), device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(Ca, device=DEV)
    skip = torch.randn((N, Cb, 2 * h, 2 * w), device=DEV).contiguous(memory_format=torch.channels_last)
    y = y4.view(N, 2, 2, Ca, h, w).permute(0, 3, 4, 1, 5, 2).reshape(N, Ca, 2 * h, 2 * w)     # channel (2a+b)*Ca + c -> pixel (2i+a, 2j+b)
    ref = torch.cat([torch.nn.functional.leaky_relu(y + bias.view(1, -1, 1, 1), 0.2), skip], dim=1)
    got = ops.d2s_bias_act_concat(y4, bias, skip, 0.2)
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, ref)
    assert torch.equal(ops.d2s_bias_act_concat(y4, bias, None, 0.2), ref[:, :Ca])
    # 2x2-tap form: phase (a, b) of pixel (i, j) sits at (i + a, j + b) of an (h+1, w+1) map
    y2 = torch.randn((N, 4 * Ca, h + 1, w + 1), device=DEV).contiguous(memory_format=torch.channels_last)
    ref2 = torch.empty((N, Ca, 2 * h, 2 * w), device=DEV)
    for a in (0, 1):
        for b in (0, 1):
            ref2[:, :, a::2, b::2] = y2[:, (2 * a + b) * Ca:(2 * a + b + 1) * Ca, a:a + h, b:b + w]
    ref2 = torch.cat([torch.nn.functional.leaky_relu(ref2 + bias.view(1, -1, 1, 1), 0.2), skip], dim=1)
    assert torch.equal(ops.d2s_bias_act_concat(y2, bias, skip, 0.2, taps=2), ref2)


@pytest.mark.parametrize("sy,sx", [(0, 1), (1, 0), (0, 0)])
def test_bias_act_pad_is_bit_exact_with_torch(sy, sx):
    x = torch.randn((3, 32, 9, 7), device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(32, device=DEV)
    act = torch.nn.functional.leaky_relu(x + bias.view(1, -1, 1, 1), 0.2)
    ref = torch.nn.functional.pad(act, (sx, 0, sy, 0), mode="replicate")
    got = ops.bias_act_pad(x, bias, 0.2, sy, sx)
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, ref)
    assert torch.equal(got[:, :, sy:, sx:], act)


@pytest.mark.parametrize("taps", [2, 3])
def test_subpixel_decoder_block_equals_resize_then_conv(taps):
    """DecoderBlock fast path (exact 2x): sub-pixel convolution == nearest resize + 3x3 convolution (fp32 rounding only)."""
    from riders_b200.rcnet import networks
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(4)
    blk = networks.DecoderBlock(32, 16, 16, 'kaiming_uniform', 'leaky_relu', True).to(DEV).eval()
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.3, 0.3); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.uniform_(-0.2, 0.2)
    x = torch.randn((4, 32, 15, 6), device=DEV).contiguous(memory_format=torch.channels_last)
    skip = torch.randn((4, 16, 30, 12), device=DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ref = blk(x, skip)                                           # reference order: interpolate, conv, BN, act, cat, conv, BN, act
        for m in blk.modules():
            if isinstance(m, networks.Conv2d):
                m.fold_batch_norm_()
        blk.to(memory_format=torch.channels_last)
        blk.fold_upsample_()
        assert blk._w4 is not None and tuple(blk._w4.shape) == (64, 32, 3, 3) and tuple(blk._w2.shape) == (64, 32, 2, 2)
        blk.subpixel_taps = taps
        got = blk(x, skip)
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("hw,target", [((30, 12), (60, 25)), ((7, 3), (15, 6)), ((5, 4), (10, 9))])
@pytest.mark.parametrize("placed", [False, True])
@pytest.mark.parametrize("taps", [2, 3])
def test_subpixel_decoder_block_odd_targets(hw, target, placed, taps, monkeypatch):
    """DecoderBlock fast path for 2n+1 targets along one axis (RC-Net: 30x12 -> 60x25, 7x3 -> 15x6): the shifted sub-pixel
    convolution + first-column / first-row phase convolution == F.interpolate(nearest) + 3x3 convolution (fp32 rounding only),
    and the resized map is never produced."""
    from riders_b200.rcnet import networks
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(7)
    blk = networks.DecoderBlock(32, 16, 24, 'kaiming_uniform', 'leaky_relu', True).to(DEV).eval()
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.3, 0.3); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.uniform_(-0.2, 0.2)
    x = torch.randn((3, 32) + hw, device=DEV).contiguous(memory_format=torch.channels_last)
    skip = torch.randn((3, 16) + target, device=DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ref = blk(x, skip)
        for m in blk.modules():
            if isinstance(m, networks.Conv2d):
                m.fold_batch_norm_()
        blk.to(memory_format=torch.channels_last)
        blk.fold_upsample_()
        if placed:
            view = ops.placed_skip_buffers(3, [24], [16], [target], x.device)[0]
            view.copy_(skip)
            skip = view
        def no_resize(*a, **k):
            raise AssertionError("the resized map must not be materialised on this path")
        monkeypatch.setattr(ops, "upsample_nearest", no_resize)
        blk.subpixel_taps = taps
        got = blk(x, skip)
        # the padded operand handed over by the previous block's bias + activation pass (MultiScaleDecoder.forward)
        sy, sx = int(target[0] == 2 * hw[0] + 1), int(target[1] == 2 * hw[1] + 1)
        xp = torch.nn.functional.pad(x, (sx, 0, sy, 0), mode="replicate").contiguous(memory_format=torch.channels_last)
        got2 = blk(xp[:, :, sy:, sx:], skip, xp=xp)
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert torch.equal(got, got2)


def test_decoder_hands_padded_maps_to_odd_blocks(monkeypatch):
    """MultiScaleDecoder at RC-Net's sizes (7x3 -> 15x6 -> 30x12 -> 60x25 -> 120x50 -> 240x100): the block in front of the
    60x25 one returns its output replicate-padded (bias_act_pad), nothing calls F.pad, same logits."""
    from riders_b200.rcnet import networks
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(11)
    dec = networks.MultiScaleDecoder(64, 1, 1, [32, 32, 16, 16, 16], [32, 32, 16, 8, 0], 'kaiming_uniform', 'leaky_relu', 'linear',
                                     True, 'up').to(DEV).eval()
    N = 3
    x = torch.randn((N, 64, 7, 3), device=DEV)
    skips = [torch.randn((N, c) + hw, device=DEV) for c, hw in ((8, (120, 50)), (16, (60, 25)), (32, (30, 12)), (32, (15, 6)))]
    with torch.no_grad():
        ref = dec(x, skips, shape=(240, 100))[-1]
        for m in dec.modules():
            if isinstance(m, networks.Conv2d):
                m.fold_batch_norm_()
        dec.to(memory_format=torch.channels_last)
        for m in dec.modules():
            if isinstance(m, networks.DecoderBlock):
                m.fold_upsample_()
        xc = x.contiguous(memory_format=torch.channels_last)
        sc = [t.contiguous(memory_format=torch.channels_last) for t in skips]
        calls = []
        real_pad, real_bap = torch.nn.functional.pad, ops.bias_act_pad
        monkeypatch.setattr(networks.F, "pad", lambda t, *a, **k: (calls.append(("pad", tuple(t.shape))), real_pad(t, *a, **k))[1])
        monkeypatch.setattr(ops, "bias_act_pad", lambda t, *a, **k: (calls.append(("bap", tuple(t.shape))), real_bap(t, *a, **k))[1])
        got = dec(xc, sc, shape=(240, 100))[-1]
    assert ("bap", (N, 32, 30, 12)) in calls, calls
    assert not [c for c in calls if c[0] == "pad"], calls                                   # torch's pad kernel is never used
    assert ("bap", (N, 64, 7, 3)) in calls, calls                                          # the latent: identity bias_act_pad
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=2e-4, atol=2e-5)


def _tf32_rna(x):
    """cvt.rna.tf32.f32 in torch: round the magnitude to 10 mantissa bits, ties away from zero."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("taps", [2, 3])
@pytest.mark.parametrize("hw", [(120, 50), (7, 9), (13, 30)])
def test_decoder_tail_matches_the_three_passes(hw, taps):
    """ops.decoder_tail (d2s + bias + act -> 16->16 conv on TF32 tensor cores -> bias + act -> 16->1 head, TF32 too) against
    torch: against fp32 convolutions of the TF32-ROUNDED operands (same rounding points as the kernel) the mean error is at
    fp32 level and only values whose z sits on a TF32 rounding boundary move (by one TF32 ulp of a product); against the
    pure fp32 passes it is within TF32 accuracy (2e-3 of the logit scale)."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] * 7 + taps)
    N, (h, w), t = 3, hw, (1 if taps == 2 else 0)
    y4 = torch.randn((N, 64, h + t, w + t), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=taps).contiguous()           # bit-exact pass (tested above)
    z32 = F.leaky_relu(F.conv2d(y, wb, bb, 1, 1), 0.2)
    ref32 = F.conv2d(z32, wh, None, 1, 1)
    zt = F.leaky_relu(F.conv2d(_tf32_rna(y), _tf32_rna(wb), None, 1, 1) + bb.view(1, -1, 1, 1), 0.2)
    reft = F.conv2d(_tf32_rna(zt), _tf32_rna(wh), None, 1, 1)               # the head runs on the tensor cores too
    got = ops.decoder_tail(y4, ba, wb, bb, wh, 0.2, taps=taps)
    sg = ops.decoder_tail(y4, ba, wb, bb, wh, 0.2, taps=taps, sigmoid=True)
    assert got.shape == ref32.shape
    scale = float(ref32.abs().max())
    err = (got - reft).abs()
    assert float(err.mean()) < 2e-6 * scale and float(err.max()) < 5e-4 * scale
    assert float((err > 2e-5 * scale).float().mean()) < 0.05
    assert float((got - ref32).abs().max()) < 2e-3 * scale
    np.testing.assert_allclose(sg.cpu().numpy(), torch.sigmoid(got).cpu().numpy(), rtol=2e-6, atol=1e-7)


def _fp16(x):
    return x.half().float()


@pytest.mark.parametrize("taps", [2, 3])
@pytest.mark.parametrize("hw,N", [((120, 50), 3), ((7, 9), 2), ((13, 30), 5), ((2, 3), 1), ((40, 70), 2), ((120, 50), 40), ((9, 130), 1)])
def test_decoder_tail_tcgen05_matches_fp16_operand_convolutions(hw, N, taps):
    """ops.decoder_tail_tc (tcgen05.mma kind::f16, accumulators in TMEM, streaming row pipeline): fp16 operands (11-bit
    significand, like TF32) with fp32 accumulation.  Against fp32 convolutions of the fp16-ROUNDED operands (the kernel's own
    rounding points: y, Wb, z, Wh) only values whose z sits on an fp16 rounding boundary move (by one fp16 ulp of a product);
    against the pure fp32 passes it is within fp16 / TF32 operand accuracy.  Shapes: RC-Net's 240 x 100, tiny maps, several
    column strips (2 w > 124), rows split over CTAs (few points), many strips per SM (N = 40); deterministic."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] * 7 + taps + N)
    (h, w), t = hw, (1 if taps == 2 else 0)
    y4 = torch.randn((N, 64, h + t, w + t), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=taps).contiguous()           # bit-exact pass (tested above)
    ref32 = F.conv2d(F.leaky_relu(F.conv2d(y, wb, bb, 1, 1), 0.2), wh, None, 1, 1)
    zh = F.leaky_relu(F.conv2d(_fp16(y), _fp16(wb), None, 1, 1) + bb.view(1, -1, 1, 1), 0.2)
    refh = F.conv2d(_fp16(zh), _fp16(wh), None, 1, 1)
    got = ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=taps)
    sg = ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=taps, sigmoid=True)
    assert got.shape == ref32.shape
    scale = float(ref32.abs().max())
    err = (got - refh).abs()
    assert float(err.mean()) < 4e-6 * scale and float(err.max()) < 5e-4 * scale
    assert float((err > 2e-5 * scale).float().mean()) < 0.05
    assert float((got - ref32).abs().max()) < 2e-3 * scale
    np.testing.assert_allclose(sg.cpu().numpy(), torch.sigmoid(got).cpu().numpy(), rtol=2e-6, atol=1e-7)
    assert torch.equal(got, ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=taps))
    # the two tail kernels agree to operand accuracy (TF32 and fp16 keep the same number of significand bits)
    assert float((got - ops.decoder_tail(y4, ba, wb, bb, wh, 0.2, taps=taps)).abs().max()) < 2e-3 * scale


@pytest.mark.parametrize("hw,N", [((120, 50), 3), ((7, 9), 2), ((13, 30), 5), ((2, 3), 1), ((120, 50), 40), ((31, 54), 2)])
def test_upconv_and_tail_on_tcgen05_through_the_fp16_operand_map(hw, N):
    """ops.subpixel_conv_tc_half (tcgen05 phase convolution whose epilogue finishes the up-convolution: depth-to-space + bias +
    LeakyReLU + fp16, written as the tail's operand image) and ops.decoder_tail_tc_half (rows bulk-copied into the operand
    ring).  The map equals fp16(act(d2s(conv2x2(TF32-truncated x, TF32-rounded W2)) + bias)) up to the fp32 summation order
    (one fp16 ulp where the sum sits on a rounding boundary); the logits follow the same reference as the float-input tail."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] + 3 * N)
    h, w = hw
    x = torch.randn((N, 32, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    x_trunc = (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    y4 = F.conv2d(x_trunc, _tf32_rna(w2), None, 1, 1).contiguous(memory_format=torch.channels_last)
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=2).contiguous()
    yh, flag = ops.subpixel_conv_tc_half(x, w2, ba, 0.2)
    assert yh.shape == (N, 2 * h, 2, 2 * w, 8) and yh.dtype == torch.float16 and int(flag.item()) == 0
    yref = y.view(N, 2, 8, 2 * h, 2 * w).permute(0, 3, 1, 4, 2)                      # [n][Y][plane][X][k]
    ysc = float(yref.abs().max())
    ey = (yh.float() - yref).abs()
    assert float(ey.max()) <= 2.0 ** -10 * ysc and float((ey > 2.0 ** -11 * yref.abs() + 1e-7).float().mean()) < 0.02
    y16 = yh.float().permute(0, 2, 4, 1, 3).reshape(N, 16, 2 * h, 2 * w)           # the map the tail really reads
    zh = F.leaky_relu(F.conv2d(y16, _fp16(wb), None, 1, 1) + bb.view(1, -1, 1, 1), 0.2)
    refh = F.conv2d(_fp16(zh), _fp16(wh), None, 1, 1)
    ref32 = F.conv2d(F.leaky_relu(F.conv2d(y, wb, bb, 1, 1), 0.2), wh, None, 1, 1)
    out, _ = ops.decoder_tail_tc_half(yh, wb, bb, wh, 0.2, flag=flag)
    got = ops.decoder_upconv_tail_tc(x, w2, ba, wb, bb, wh, 0.2, producer="tma")
    assert torch.equal(out, got) and int(flag.item()) == 0
    scale = float(ref32.abs().max())
    err = (got - refh).abs()
    assert float(err.mean()) < 4e-6 * scale and float(err.max()) < 5e-4 * scale
    assert float((err > 2e-5 * scale).float().mean()) < 0.05
    # against the library flow it replaces (cuDNN fp32 phase convolution + the three fp32 passes): operand accuracy
    y32 = ops.d2s_bias_act_concat(F.conv2d(x, w2, None, 1, 1).contiguous(memory_format=torch.channels_last), ba, None, 0.2, taps=2)
    full32 = F.conv2d(F.leaky_relu(F.conv2d(y32.contiguous(), wb, bb, 1, 1), 0.2), wh, None, 1, 1)
    assert float((got - full32).abs().max()) < 3e-3 * float(full32.abs().max())
    sg = ops.decoder_upconv_tail_tc(x, w2, ba, wb, bb, wh, 0.2, sigmoid=True, producer="tma")
    np.testing.assert_allclose(sg.cpu().numpy(), torch.sigmoid(got).cpu().numpy(), rtol=2e-6, atol=1e-7)
    # range guard: an activation beyond fp16 -> the TF32 library flow answers
    xb = x.clone()
    xb[0, :, 0, 0] = 3.0e5
    refb = ops.decoder_tail(F.conv2d(xb, w2, None, 1, 1), ba, wb, bb, wh, 0.2, taps=2)
    torch.backends.cudnn.allow_tf32 = False
    assert torch.equal(ops.decoder_upconv_tail_tc(xb, w2, ba, wb, bb, wh, 0.2, producer="tma"), refb)


@pytest.mark.parametrize("hw,N,raw", [((120, 50), 3, True), ((7, 9), 2, False), ((13, 30), 5, True), ((2, 3), 1, True),
                                      ((120, 50), 200, True), ((31, 54), 2, True), ((9, 8), 160, False)])
def test_upconv_tcgen05_from_the_raw_map_to_the_fp16_operand_map(hw, N, raw):
    """ops.upconv_tc_half (csrc/upconv_tc.cu): previous bias + LeakyReLU, 2x2-tap phase convolution on tcgen05 (fp16 operands staged
    once, taps as shifted views), depth-to-space + bias + LeakyReLU + fp16.  Against torch on the fp16-rounded operands the map
    differs by at most one fp16 rounding; through the tail it matches the library flow to operand accuracy; several points per
    persistent CTA (N = 200 / 160), partial batches and pairs (h + 1 not a multiple of 8), w at the limit (54)."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] + 3 * N)
    h, w = hw
    x = torch.randn((N, 32, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
    bp = torch.randn(32, device=DEV, generator=g) * 0.3 if raw else None
    sp = 0.2 if raw else 1.0
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    xa = F.leaky_relu(x + bp.view(1, -1, 1, 1), sp) if raw else x
    y4 = F.conv2d(_fp16(xa), _fp16(w2), None, 1, 1).contiguous(memory_format=torch.channels_last)
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=2).contiguous()
    yref = y.view(N, 2, 8, 2 * h, 2 * w).permute(0, 3, 1, 4, 2)                      # [n][Y][plane][X][k]
    yh, flag = ops.upconv_tc_half(x, bp, sp, w2, ba, 0.2)
    assert yh.shape == (N, 2 * h, 2, 2 * w, 8) and yh.dtype == torch.float16 and int(flag.item()) == 0
    ysc = float(yref.abs().max())
    ey = (yh.float() - yref).abs()
    assert float(ey.max()) <= 2.0 ** -10 * ysc and float((ey > 2.0 ** -11 * yref.abs() + 1e-7).float().mean()) < 0.02
    assert torch.equal(yh, ops.upconv_tc_half(x, bp, sp, w2, ba, 0.2)[0])            # deterministic
    # through the tail, against the library flow it replaces (fp32 activation pass, phase convolution, three passes)
    got = ops.decoder_upconv_tail_tc(x, w2, ba, wb, bb, wh, 0.2, bias_p=bp, slope_p=sp)
    assert torch.equal(got, ops.decoder_tail_tc_half(yh, wb, bb, wh, 0.2)[0])
    y32 = ops.d2s_bias_act_concat(F.conv2d(xa, w2, None, 1, 1).contiguous(memory_format=torch.channels_last), ba, None, 0.2, taps=2)
    full32 = F.conv2d(F.leaky_relu(F.conv2d(y32.contiguous(), wb, bb, 1, 1), 0.2), wh, None, 1, 1)
    assert float((got - full32).abs().max()) < 3e-3 * float(full32.abs().max())
    other = ops.decoder_upconv_tail_tc(x, w2, ba, wb, bb, wh, 0.2, bias_p=bp, slope_p=sp, producer="tma")
    assert float((got - other).abs().max()) < 3e-3 * float(full32.abs().max())      # the TMA-fed TF32 producer agrees
    # range guard: an activation beyond fp16 -> the TF32 library flow answers
    xb = x.clone()
    xb[0, :, 0, 0] = 3.0e5
    xab = F.leaky_relu(xb + bp.view(1, -1, 1, 1), sp) if raw else xb
    refb = ops.decoder_tail(F.conv2d(xab, w2, None, 1, 1), ba, wb, bb, wh, 0.2, taps=2)
    assert torch.equal(ops.decoder_upconv_tail_tc(xb, w2, ba, wb, bb, wh, 0.2, bias_p=bp, slope_p=sp), refb)


@pytest.mark.parametrize("hw,N,Cs,raw,placed", [((60, 25), 3, 32, False, True), ((8, 5), 1, 32, False, False), ((7, 9), 3, 0, True, False),
                                                ((13, 26), 2, 8, True, False), ((60, 25), 300, 32, False, True)])
def test_upconv_tcgen05_float_output_for_the_64_to_32_block(hw, N, Cs, raw, placed):
    """ops.upconv_tc_f32 (csrc/upconv_tc.cu, the 64 -> 32 channel instance: RC-Net's deconv1): phase convolution on tcgen05 +
    depth-to-space + bias + LeakyReLU written as float32 into the leading channels of the concat buffer, the skip behind them
    untouched (placed) or copied.  Against the library passes on fp16-rounded operands: fp32 summation order only; against
    the pure fp32 passes: operand accuracy.  Several points per persistent CTA (N = 300), w at the limit (26)."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] + 5 * N)
    h, w = hw
    x = torch.randn((N, 64, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((128, 64, 2, 2), device=DEV, generator=g) * 0.08
    bp = torch.randn(64, device=DEV, generator=g) * 0.3 if raw else None
    sp = 0.2 if raw else 1.0
    ba = torch.randn(32, device=DEV, generator=g) * 0.3
    if placed:
        parent = torch.randn((N, 32 + Cs, 2 * h, 2 * w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
        skip = parent[:, 32:]
        assert ops.placed_parent(skip, 32) is not None
        skip_ref = skip.clone()
    else:
        skip = torch.randn((N, Cs, 2 * h, 2 * w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last) if Cs else None
        skip_ref = skip
    xa = F.leaky_relu(x + bp.view(1, -1, 1, 1), sp) if raw else x
    y4 = F.conv2d(_fp16(xa), _fp16(w2), None, 1, 1).contiguous(memory_format=torch.channels_last)
    ref = ops.d2s_bias_act_concat(y4, ba, skip_ref, 0.2, taps=2)
    ref32 = ops.d2s_bias_act_concat(F.conv2d(xa, w2, None, 1, 1).contiguous(memory_format=torch.channels_last), ba, skip_ref, 0.2, taps=2)
    got, flag = ops.upconv_tc_f32(x, bp, sp, w2, ba, 0.2, skip)
    assert got.shape == ref.shape and ops._is_channels_last(got) and int(flag.item()) == 0
    if placed:
        assert got.data_ptr() == parent.data_ptr()
    sc = float(ref[:, :32].abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=2e-5 * sc)
    if Cs:
        assert torch.equal(got[:, 32:], skip_ref)
    assert float((got - ref32).abs().max()) < 3e-3 * sc
    assert torch.equal(got, ops.upconv_tc_f32(x, bp, sp, w2, ba, 0.2, skip)[0])
    xb = x.clone()
    xb[0, 3, 0, 0] = 1.0e6
    assert int(ops.upconv_tc_f32(xb, bp, sp, w2, ba, 0.2, skip)[1].item()) == 1            # staged value beyond the fp16 range


@pytest.mark.parametrize("hw,N", [((120, 50), 3), ((8, 5), 1), ((13, 54), 2), ((120, 50), 300)])
def test_conv3x3_tcgen05_instance_matches_torch(hw, N):
    """ops.conv3x3_tc_f32: the 3x3 / nine-tap instance of the up-convolution kernel template (64 -> 32 channels, raw float32
    out).  Not on the chain - it measures 1.43 ms against cuDNN's 1.21 ms per 2048 points (36 MMAs of N = 32 per tile re-read
    the A operand) - but it is exported and therefore tested: fp32 summation order against torch on fp16-rounded operands."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] + N)
    h, w = hw
    x = torch.randn((N, 64, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    wt = torch.randn((32, 64, 3, 3), device=DEV, generator=g) * 0.05
    ref = F.conv2d(_fp16(x), _fp16(wt), None, 1, 1)
    got, flag = ops.conv3x3_tc_f32(x, wt)
    assert got.shape == ref.shape and ops._is_channels_last(got) and int(flag.item()) == 0
    sc = float(ref.abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=3e-5 * sc)
    assert float((got - F.conv2d(x, wt, None, 1, 1)).abs().max()) < 3e-3 * sc
    b = torch.randn(32, device=DEV, generator=g) * 0.3
    act, _ = ops.conv3x3_tc_f32(x, wt, bias=b, slope=0.2)
    np.testing.assert_allclose(act.cpu().numpy(), F.leaky_relu(ref + b.view(1, -1, 1, 1), 0.2).cpu().numpy(), rtol=1e-5, atol=3e-5 * sc)


def test_decoder_tail_tcgen05_fp16_range_guard():
    """Activations beyond the fp16 range (a limit neither the fp32 reference nor the TF32 kernel has) raise the device flag of
    riders_decoder_tail_tc_nhwc: ops.decoder_tail_tc then returns the mma.sync TF32 kernel's result (bit-identical to calling
    it), immediately or - with defer=True - once the caller has read the pending flags."""
    g = torch.Generator(device=DEV).manual_seed(3)
    y4 = torch.randn((2, 64, 13, 10), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    ok = ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=2)
    assert torch.isfinite(ok).all()
    big = y4.clone()
    big[1, 5, 4, 3] = 1.0e5                                              # one activation out of range
    ref = ops.decoder_tail(big, ba, wb, bb, wh, 0.2, taps=2)
    assert torch.equal(ops.decoder_tail_tc(big, ba, wb, bb, wh, 0.2, taps=2), ref)
    zbig, wbp = y4.clone(), wb.abs()                                     # y in range, z = conv(y) beyond it
    zbig[0, :, 3:6, 2:5] = 3.0e4
    refz = ops.decoder_tail(zbig, ba, wbp, bb, wh, 0.2, taps=2)
    assert torch.equal(ops.decoder_tail_tc(zbig, ba, wbp, bb, wh, 0.2, taps=2), refz)
    ops.LoftrOverflow.pending.clear()
    ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=2, defer=True)
    assert len(ops.LoftrOverflow.pending) == 1 and not ops.loftr_overflow_pending()
    ops.decoder_tail_tc(big, ba, wb, bb, wh, 0.2, taps=2, defer=True)
    assert ops.loftr_overflow_pending() and not ops.LoftrOverflow.pending


@pytest.mark.parametrize("hw", [(120, 50), (7, 9), (13, 30)])
def test_subpixel_conv_act_matches_torch(hw):
    """ops.subpixel_conv_act (mma.sync): conv2x2(act(x + bias), W2, padding 1) on the TF32 tensor cores == fp32 convolution
    of the TF32-rounded operands (one rounding point: 1e-5 of the scale), and within TF32 accuracy of the pure fp32 result."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0])
    N, (h, w) = 3, hw
    x = torch.randn((N, 32, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    b = torch.randn(32, device=DEV, generator=g) * 0.3
    w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
    xa = F.leaky_relu(x + b.view(1, -1, 1, 1), 0.2)
    ref32 = F.conv2d(xa, w2, None, 1, 1)
    reft = F.conv2d(_tf32_rna(xa), _tf32_rna(w2), None, 1, 1)
    got = ops.subpixel_conv_act(x, b, 0.2, w2)
    assert got.shape == ref32.shape == (N, 64, h + 1, w + 1) and got.is_contiguous(memory_format=torch.channels_last)
    scale = float(ref32.abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), reft.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)
    assert float((got - ref32).abs().max()) < 2e-3 * scale
    got1 = ops.subpixel_conv_act(xa, None, 1.0, w2)                      # already activated input: identity "activation"
    np.testing.assert_allclose(got1.cpu().numpy(), reft.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)


@pytest.mark.parametrize("hw,N", [((120, 50), 3), ((7, 9), 2), ((13, 30), 5), ((120, 50), 40)])
def test_subpixel_conv_tc_matches_torch(hw, N):
    """ops.subpixel_conv_tc (TMA -> tcgen05.mma kind::tf32 -> TMEM): conv2x2(x, W2, padding 1).  The tensor core truncates
    the fp32 activations to TF32 (W2 is rounded when packed): exactly the fp32 convolution of those operands (1e-5 of the
    scale) and within TF32 accuracy of the pure fp32 result; several tiles per persistent CTA in the N = 40 case."""
    F = torch.nn.functional
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(hw[0] + N)
    h, w = hw
    x = torch.randn((N, 32, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
    ref32 = F.conv2d(x, w2, None, 1, 1)
    x_trunc = (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    reft = F.conv2d(x_trunc, _tf32_rna(w2), None, 1, 1)
    got = ops.subpixel_conv_tc(x, w2)
    assert got.shape == ref32.shape == (N, 64, h + 1, w + 1) and got.is_contiguous(memory_format=torch.channels_last)
    scale = float(ref32.abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), reft.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)
    assert float((got - ref32).abs().max()) < 3e-3 * scale
    assert torch.equal(got, ops.subpixel_conv_tc(x, w2))                 # deterministic


def test_decoder_uses_the_fused_tail_only_where_tf32_is_allowed(monkeypatch):
    """MultiScaleDecoder: with torch.backends.cudnn.allow_tf32 the tail runs in ops.decoder_tail and agrees with the cuDNN
    TF32 passes to TF32 accuracy; without it the fp32 passes run (the 1e-4 parity tests above are untouched)."""
    from riders_b200.rcnet import networks
    torch.manual_seed(12)
    dec = networks.MultiScaleDecoder(64, 1, 1, [32, 32, 16, 32, 16], [32, 32, 16, 8, 0], 'kaiming_uniform', 'leaky_relu', 'linear',
                                     True, 'up').to(DEV).eval()
    N = 3
    x = torch.randn((N, 64, 7, 3), device=DEV).contiguous(memory_format=torch.channels_last)
    skips = [torch.randn((N, c) + hw, device=DEV).contiguous(memory_format=torch.channels_last)
             for c, hw in ((8, (120, 50)), (16, (60, 25)), (32, (30, 12)), (32, (15, 6)))]
    with torch.no_grad():
        for m in dec.modules():
            if isinstance(m, networks.Conv2d):
                m.fold_batch_norm_()
        dec.to(memory_format=torch.channels_last)
        for m in dec.modules():
            if isinstance(m, networks.DecoderBlock):
                m.fold_upsample_()
        calls, subs, tcs = [], [], []
        real, real_sub, real_tc, real_tail_tc = ops.decoder_tail, ops.subpixel_conv_act, ops.subpixel_conv_tc, ops.decoder_tail_tc
        tail_tcs = []
        monkeypatch.setattr(ops, "subpixel_conv_tc", lambda *a, **k: (tcs.append(1), real_tc(*a, **k))[1])
        monkeypatch.setattr(ops, "decoder_tail", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
        monkeypatch.setattr(ops, "decoder_tail_tc", lambda *a, **k: (calls.append(1), tail_tcs.append(1), real_tail_tc(*a, **k))[2])
        assert dec.tail_impl == os.environ.get("RIDERS_TAIL_IMPL", "tcgen05")
        monkeypatch.setattr(ops, "subpixel_conv_act", lambda *a, **k: (subs.append(1), real_sub(*a, **k))[1])
        torch.backends.cudnn.allow_tf32 = False
        ref = dec(x, skips, shape=(240, 100))[-1]
        assert not calls and not subs
        torch.backends.cudnn.allow_tf32 = True
        try:
            dec.tail_fp16_map = True                                  # default: tcgen05 phase convolution -> fp16 operand map -> tcgen05 tail
            if dec.tail_impl == "tcgen05":
                got5 = dec(x, skips, shape=(240, 100))[-1]
                assert dec.last_tail_impl == "tcgen05+fp16map" and dec.last_used_tail and not calls and not tcs
                assert float((got5 - ref).abs().max()) < 5e-3 * float(ref.abs().max())
            dec.tail_fp16_map = False                                 # the float y4 interface (RIDERS_TAIL_FP16_MAP=0)
            dec.fused_sub = True                                      # opt-in (RIDERS_FUSED_SUBCONV=1)
            got = dec(x, skips, shape=(240, 100))[-1]                 # mma.sync phase convolution (+ fused activation) + tail
            assert len(calls) == 1 and len(subs) == 1 and not tcs
            dec.fused_sub = False
            dec.subconv_tc = True                                     # opt-in (RIDERS_SUBCONV_TC=1)
            got2 = dec(x, skips, shape=(240, 100))[-1]                # tcgen05 phase convolution + tail
            assert len(calls) == 2 and len(subs) == 1 and len(tcs) == 1
            dec.subconv_tc = False
            got3 = dec(x, skips, shape=(240, 100))[-1]                # cuDNN phase convolution + tail
            assert len(calls) == 3 and len(subs) == 1 and len(tcs) == 1
            n_tc = len(tail_tcs)
            dec.tail_impl = "mma_sync" if dec.tail_impl == "tcgen05" else "tcgen05"       # the other tail kernel
            got4 = dec(x, skips, shape=(240, 100))[-1]
            assert len(calls) == 4 and len(tail_tcs) == n_tc + (dec.tail_impl == "tcgen05") and dec.last_tail_impl == dec.tail_impl
            calls.pop(); calls.pop()
            dec.fused_tail = False
            sep = dec(x, skips, shape=(240, 100))[-1]                 # the separate passes
            assert len(calls) == 2 and len(subs) == 1
        finally:
            torch.backends.cudnn.allow_tf32 = False
    scale = float(ref.abs().max())
    for name, t in (("mma.sync phase conv + tail", got), ("tcgen05 phase conv + tail", got2), ("cuDNN phase conv + tail", got3),
                    ("other tail kernel", got4), ("separate passes", sep)):
        assert float((t - ref).abs().max()) < 5e-3 * scale, name
    assert float((got - sep).abs().max()) < 5e-3 * scale


# ------------------------------------------------------------------------------------ config 4: training step
def test_training_step_gradients_match_the_reference_path():
    """rcnet_main.train's inner step (forward with logits, masked BCE, backward) through the mirror on the GPU
    (deterministic gather backward + differentiable LoFTR) against the oracle path (torchvision roi_pool) on the CPU."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    patch, K = (64, 32), 8
    cpu = _rand_model("cpu", patch, [8, 16, 16, 16, 16], [32, 16, 16, 8, 8], seed=5)
    gpu = _to_cuda(cpu)
    cpu.train(); gpu.train()
    fr = synth.chain_frame(11, K, 96, 128, patch)
    padded = torch.nn.functional.pad(torch.from_numpy(fr["image"])[None], (16, 16, 32, 32), mode="replicate")
    pts, boxes = torch.from_numpy(fr["points"]), torch.from_numpy(fr["boxes"])
    rng = np.random.default_rng(0)
    gt = torch.from_numpy((rng.uniform(size=(K, 1) + patch) > 0.6).astype(np.float32))
    valid = torch.from_numpy((rng.uniform(size=(K, 1) + patch) > 0.2).astype(np.float32))
    # reference path on the CPU with autograd (oracle functions without their no_grad decorators)
    latent, skips = rcnet_ref.encoder_forward_ref.__wrapped__(cpu.encoder, padded, pts, [boxes])
    logits_ref = cpu.decoder(x=latent, skips=skips, shape=cpu.input_patch_size_image)[-1]
    loss_ref, _ = cpu.compute_loss(logits_ref, gt, valid, w_positive_class=2.5)
    loss_ref.backward()
    logits = gpu.forward(padded.to(DEV), pts.to(DEV), [boxes.to(DEV)], return_logits=True)
    loss, _ = gpu.compute_loss(logits, gt.to(DEV), valid.to(DEV), w_positive_class=2.5)
    loss.backward()
    np.testing.assert_allclose(loss.item(), loss_ref.item(), rtol=1e-5)
    checked = 0
    for (name, pc), pg in zip(list(cpu.encoder.named_parameters()) + list(cpu.decoder.named_parameters()),
                              list(gpu.encoder.parameters()) + list(gpu.decoder.parameters())):
        if pc.grad is None:
            assert pg.grad is None or float(pg.grad.abs().max()) == 0.0, name
            continue
        scale = float(pc.grad.abs().max()) + 1e-12
        # train-mode BatchNorm amplifies fp32 re-association, and cuDNN's weight-gradient kernels are not run-to-run deterministic
        # (measured 1e-3 .. 4.2e-3 of the gradient scale on the stem convolution across visits)
        assert float((pg.grad.cpu() - pc.grad).abs().max()) <= 1e-2 * scale + 1e-7, name
        checked += 1
    assert checked > 50


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("cross", [False, True])
def test_loftr_layer_backward_kernels_match_autograd(dtype, cross):
    """riders_loftr_layer_fwd_train / riders_loftr_layer_bwd (ops.LoftrLayerFunction) at RC-Net's real layer shape (d_model
    128, 8 heads, 21 tokens) against torch autograd through the reference formulation of the layer (oracle.rcnet_ref
    .loftr_layer_torch, fp32): output, d_x, d_source and the gradients of all ten parameters.  float32 tokens: 2e-4 of each
    gradient's scale (different summation order only); bfloat16 tokens (fp32 master weights and accumulation): 3e-2, the
    bf16 mantissa (2^-8) through ~10 dependent roundings."""
    from riders_b200.rcnet.linear_attention import LoFTREncoderLayer
    torch.manual_seed(7)
    N = 37
    layer = LoFTREncoderLayer(128, 8).to(DEV)
    for p_ in layer.parameters():
        if p_.dim() > 1:
            torch.nn.init.xavier_uniform_(p_)
    with torch.no_grad():
        layer.norm1.weight.uniform_(0.5, 1.5); layer.norm1.bias.uniform_(-0.2, 0.2)
        layer.norm2.weight.uniform_(0.5, 1.5); layer.norm2.bias.uniform_(-0.2, 0.2)
    x0 = torch.randn(N, 21, 128, device=DEV)
    s0 = torch.randn(N, 21, 128, device=DEV) if cross else None
    g = torch.randn(N, 21, 128, device=DEV)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    # reference: fp32 autograd on the (dtype-rounded) inputs
    xr = x0.to(tdt).float().clone().requires_grad_(True)          # clone: .to() of a float32 tensor is the tensor itself
    sr = s0.to(tdt).float().clone().requires_grad_(True) if cross else xr
    ref = rcnet_ref.loftr_layer_torch(layer, xr, sr)
    ref.backward(g.to(tdt).float())
    ref_grads = {n: p_.grad.clone() for n, p_ in layer.named_parameters()}
    ref_dx, ref_ds = xr.grad.clone(), (sr.grad.clone() if cross else None)
    layer.zero_grad()
    # kernels
    xk = x0.to(tdt).clone().requires_grad_(True)
    sk = s0.to(tdt).clone().requires_grad_(True) if cross else xk
    out = ops.loftr_layer_autograd(xk, sk, layer.kernel_params(), 8, layer.attention.eps, layer.norm1.eps)
    assert out.dtype == tdt
    out.backward(g.to(tdt))
    tol = 3e-2 if dtype == "bf16" else 2e-4

    def close(a, b, name):
        scale = float(b.abs().max()) + 1e-12
        assert float((a.float() - b).abs().max()) <= tol * scale, (name, float((a.float() - b).abs().max()) / scale)

    close(out.detach(), ref.detach(), "out")
    close(xk.grad, ref_dx, "d_x")
    if cross:
        close(sk.grad, ref_ds, "d_source")
    for n, p_ in layer.named_parameters():
        assert p_.grad is not None and p_.grad.dtype == torch.float32, n
        close(p_.grad, ref_grads[n], n)
    # the module itself takes this path whenever gradients are required, and it is deterministic
    layer.zero_grad()
    x2 = x0.to(tdt).clone().requires_grad_(True)
    o2 = layer(x2, s0.to(tdt) if cross else x2)
    o2.backward(g.to(tdt))
    assert torch.equal(o2, out) and torch.equal(x2.grad, xk.grad)


def test_gather_into_placed_concat_buffers_is_bit_exact_and_never_copied():
    """Skips gathered straight into the decoder's concat buffers (riders_roi_gather_fwd_placed) + the d2s / bias-leaky
    kernels that fill the leading channels == dense gather followed by the concat kernels, bit for bit."""
    rng = np.random.default_rng(11)
    H, W, patch = 512, 640, (240, 100)
    pts, boxes = synth.radar_points(5, 20, H, W, patch)
    corner = np.array([[0, 0], [W - 1, H - 1]], np.float32) + [50, 120]
    boxes[:2] = np.concatenate([corner - [50, 120], corner + [50, 120]], 1)
    shapes = [(32, 376, 370), (64, 188, 185), (128, 94, 93), (128, 47, 47)]
    scales, sizes = rr.rcnet_levels(patch)
    scales, sizes = scales[:4], sizes[:4]
    feats = [cu(rng.standard_normal((2,) + s).astype(np.float32)).contiguous(memory_format=torch.channels_last) for s in shapes]
    rois = cu(rr.boxes_list_to_rois([boxes[:9], boxes[9:]]))
    dense = ops.roi_gather(feats, rois, scales, sizes, max_roi=(100, 240))
    lead = [32, 64, 128, 128]
    views = ops.placed_skip_buffers(rois.shape[0], lead, [s[0] for s in shapes], sizes, feats[0].device)
    for v in views:
        ops.placed_parent(v, v.stride(3) - v.shape[1]).fill_(float("nan"))
    placed = ops.roi_gather(feats, rois, scales, sizes, max_roi=(100, 240), outs=views)
    for l in range(4):
        assert placed[l].data_ptr() == views[l].data_ptr() and torch.equal(placed[l], dense[l]), "level %d" % l
        parent = ops.placed_parent(placed[l], lead[l])
        assert parent is not None and parent.shape[1] == lead[l] + shapes[l][0] and torch.isnan(parent[:, :lead[l]]).all()
        assert ops.placed_parent(placed[l], lead[l] + 4) is None and ops.placed_parent(dense[l], lead[l]) is None
    # generic kernel (no max_roi) honours the pixel stride too
    views2 = ops.placed_skip_buffers(rois.shape[0], lead, [s[0] for s in shapes], sizes, feats[0].device)
    for a, b in zip(ops.roi_gather(feats, rois, scales, sizes, outs=views2), dense):
        assert torch.equal(a, b)
    # exact-2x block (level 1/2: 60x25 -> 120x50): depth-to-space into the leading channels
    R, Ca = rois.shape[0], lead[0]
    y4 = torch.randn((R, 4 * Ca, 60, 25), device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(Ca, device=DEV)
    ref = ops.d2s_bias_act_concat(y4, bias, dense[0], 0.2)
    got = ops.d2s_bias_act_concat(y4, bias, placed[0], 0.2)
    assert got.data_ptr() == ops.placed_parent(placed[0], Ca).data_ptr() and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, ref)
    # non-2x block (level 1/4: 60x25): bias + LeakyReLU into the leading channels
    y = torch.randn((R, lead[1], 60, 25), device=DEV).contiguous(memory_format=torch.channels_last)
    b1 = torch.randn(lead[1], device=DEV)
    assert torch.equal(ops.leaky_concat(y, placed[1], 0.2, bias=b1), ops.leaky_concat(y, dense[1], 0.2, bias=b1))


@pytest.mark.parametrize("tag", ["u8", "unit"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_rcnet_prologue_matches_the_reference_statements(golden, tag, channels_last):
    """N4: shift / boxes / int()-and-/255 normalisation / edge pad against fixtures made by the reference's own
    Transforms + torchvision pad + the script's box loop (oracle/gen_golden.py gen_sml_glue)."""
    g = golden("sml_glue.npz").case("prologue_" + tag)
    patch = tuple(int(v) for v in g["patch"])
    # two frames: the golden one and a [0,1]-valued one (the max > 1 decision is per frame)
    other = np.random.default_rng(3).uniform(0, 1, g["image"].shape).astype(np.float32)
    images = np.concatenate([g["image"], other], 0)
    pts = np.concatenate([g["points"], g["points"][:3]], 0)
    K = g["points"].shape[0]
    padded, shifted, rois = ops.rcnet_prologue(cu(images), cu(pts), [0, K, K + 3], patch, channels_last=channels_last)
    assert padded.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    assert np.array_equal(padded[0:1].cpu().numpy(), g["padded"])
    ref_other, _, _ = sg.rcnet_prologue_ref(other, pts[:1], patch)
    assert np.array_equal(padded[1:2].cpu().numpy(), ref_other)
    assert np.array_equal(shifted[:K].cpu().numpy(), g["points_shifted"])
    r = rois.cpu().numpy()
    assert np.array_equal(r[:K, 1:], g["boxes"]) and np.all(r[:K, 0] == 0) and np.all(r[K:, 0] == 1)
    assert np.array_equal(r[K:, 1:], g["boxes"][:3])


def test_sml_input_assemble_matches_the_reference_transform_pipeline(golden):
    """N2: val_zju.py:181-196 (Resize INTER_NEAREST + NormalizeIntermediate + gray) against the reference's transform
    pipeline (golden) and the oracle at full size, both image layouts, bit for bit."""
    from oracle.gen_golden import sml_input_case
    for tag in ("zju", "small"):
        g = golden("sml_glue.npz").case("sml_input_" + tag)
        H, W = (int(v) for v in g["shape"])
        image, int_depth, int_scales = sml_input_case(tag, H, W)
        x, d = ops.sml_input_assemble(cu(int_depth[None]), cu(int_scales[None]), cu(image[None]), image_hwc=True)
        assert np.array_equal(d[0].cpu().numpy(), g["d"])
        xs = x[0].cpu().numpy()
        assert np.array_equal(xs if tag != "zju" else xs[:, ::4, ::4], g["x"])
    rng = np.random.default_rng(8)
    B, H, W = 3, 512, 640
    img = rng.uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    dep = rng.uniform(0.004, 10, (B, H, W)).astype(np.float32)
    sc = rng.uniform(0, 1, (B, H, W)).astype(np.float32)
    x, d = ops.sml_input_assemble(cu(dep), cu(sc), cu(np.ascontiguousarray(img.transpose(0, 3, 1, 2))))
    assert tuple(x.shape) == (B, 3, 288, 352) and tuple(d.shape) == (B, 1, 288, 352)
    for b in range(B):
        rx, rd = sg.sml_input_ref(img[b], dep[b], sc[b])
        assert np.array_equal(x[b].cpu().numpy(), rx) and np.array_equal(d[b].cpu().numpy(), rd)


def test_pipeline_raw_inputs_retry_and_sml_input():
    """The pipeline fed 'as loaded' inputs ([0,255] frames, unshifted points) == the pipeline fed the transformed ones;
    the threshold-lowering retry of run_rcnet_zju.py:254-265; SML inputs against the oracle."""
    from riders_b200.pipeline import RadarDepthPipeline
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    Hh, Ww, patch, K, B = 96, 128, (64, 32), 14, 3
    gpu = _to_cuda(_rand_model("cpu", patch, [8, 16, 16, 16, 16], [32, 16, 16, 8, 8])).channels_last()
    batch = synth.chain_batch(60, B, K, Hh, Ww, patch)
    args = (batch["offsets"], cu(batch["mono"]), cu(batch["sparse"]))
    base = RadarDepthPipeline(gpu, response_thr=0.45, chunk_points=16).run_device(cu(batch["images"]), cu(batch["points"]), *args)
    raw_images = np.rint(batch["images"] * 255.0).astype(np.float32)
    raw_points = batch["points"] - np.array([patch[1] // 2, patch[0] // 2, 0], np.float32)
    pipe = RadarDepthPipeline(gpu, response_thr=0.45, chunk_points=16, raw_inputs=True, retry_empty=True, sml_input=True)
    out = pipe.run_device(cu(raw_images), cu(raw_points), *args)
    for k in ("depth", "response", "int_depth", "scale_map", "int_scales"):
        assert torch.equal(out[k], base[k]), k
    for b in range(B):
        rx, rd = sg.sml_input_ref(batch["images"][b].transpose(1, 2, 0), out["int_depth"][b].cpu().numpy(), out["int_scales"][b].cpu().numpy())
        x = out["sml_x"][b].cpu().numpy()
        assert np.array_equal(x[:2], rx[:2]) and np.array_equal(out["sml_d"][b].cpu().numpy(), rd)
        np.testing.assert_allclose(x[2], rx[2], rtol=1e-6)            # raw / 255 on the device vs the loader's uint8 / 255
    # nothing fires at thr = 0.99 with random weights (sigmoid ~ 0.5): the retry walks the threshold down until it does
    from riders_b200.rcnet.rcnet_main import forward_output_batched
    with torch.no_grad():
        d0, r0 = forward_output_batched(gpu, cu(batch["images"]), cu(batch["points"]), batch["offsets"], 0.99, chunk_points=16)
        d1, r1 = forward_output_batched(gpu, cu(batch["images"]), cu(batch["points"]), batch["offsets"], 0.99, chunk_points=16,
                                        retry_empty=True)
    assert float(d0.abs().sum()) == 0.0
    assert all(float(d1[b].amax()) > 0 for b in range(B))
