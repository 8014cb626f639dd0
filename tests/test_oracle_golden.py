"""The CPU oracle against fixtures produced by executing the unmodified reference
(oracle/gen_golden.py).  This is what pins the oracle (the reference ships no tests)."""
import numpy as np
import pytest
import torch

from oracle import (confidence_reduce_ref as cr, estimator_ref as er, interpolator_ref as ir,
                    linear_attention_ref as lr, roi_pool_ref as rr, rcnet_ref)


def test_roi_pool_restatement_bit_exact(golden):
    g = golden("roi_pool.npz")
    rois = rr.boxes_list_to_rois([g["boxes0"], g["boxes1"]])
    for l in range(5):
        out = rr.roi_pool_ref(g["feat%d" % l], rois, float(g["scale%d" % l]), tuple(g["osize%d" % l]))
        assert np.array_equal(out, g["out%d" % l]), "level %d" % l
        assert np.array_equal(rr.roi_pool_torchvision(g["feat%d" % l], rois, float(g["scale%d" % l]),
                                                      tuple(g["osize%d" % l])), g["out%d" % l])


def test_roi_pool_argmax_points_at_max(golden):
    g = golden("roi_pool.npz")
    rois = rr.boxes_list_to_rois([g["boxes0"], g["boxes1"]])
    feat = g["feat2"]
    out, arg = rr.roi_pool_ref(feat, rois, float(g["scale2"]), tuple(g["osize2"]), return_argmax=True)
    R, C = out.shape[:2]
    flat = feat.reshape(feat.shape[0], feat.shape[1], -1)
    for r in range(R):
        b = int(rois[r, 0])
        for c in range(C):
            a = arg[r, c]
            sel = a >= 0
            assert np.array_equal(out[r, c][sel], flat[b, c][a[sel]])
            assert np.all(out[r, c][~sel] == 0)


@pytest.mark.parametrize("tag", ["small", "rcnet"])
def test_loftr_oracle_matches_reference(golden, tag):
    g = golden("loftr.npz").case(tag)
    state = {k: v for k, v in g.items() if k.startswith("layers.")}
    n_layers = int(g["n_layers"])
    o0, o1 = lr.transformer_ref(g["f0"], g["f1"], state, ("self", "cross") * n_layers)
    np.testing.assert_allclose(o0, g["o0"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(o1, g["o1"], rtol=2e-5, atol=2e-5)
    w0 = lr.split_layer_weights(state, 2 * n_layers)[0]
    np.testing.assert_allclose(lr.loftr_layer_ref(g["f0"], g["f1"], w0), g["layer0"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(lr.linear_attention_ref(g["q"], g["k"], g["v"]), g["attn"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lr.linear_attention_ref(g["q"], g["k"], g["v"], g["qm"], g["km"]),
                               g["attn_masked"], rtol=1e-5, atol=1e-6)
    # fp64 ground truth stays within the fp32 reference's own rounding noise
    d0, _ = lr.transformer_ref(g["f0"], g["f1"], state, ("self", "cross") * n_layers, dtype=np.float64)
    assert np.abs(d0 - g["o0"]).max() / np.abs(g["o0"]).max() < 1e-5


@pytest.mark.parametrize("tag", ["basic", "single", "allbelow", "dups"])
def test_confidence_reduce_oracle_matches_reference(golden, tag):
    g = golden("confidence_reduce.npz").case(tag)
    H, W = [int(v) for v in g["hw"]]
    d, r, s = cr.confidence_reduce_ref(g["conf"][:, 0], g["pts"], H, W, float(g["thr"]))
    assert np.array_equal(r, g["response"][0])                      # max is exact
    np.testing.assert_allclose(d, g["depth"][0], rtol=1e-6, atol=0)
    assert np.array_equal(d == 0, g["depth"][0] == 0)
    if tag == "allbelow":
        assert not d.any() and not r.any()
    # select really is an argmax of the thresholded tiles
    K = g["conf"].shape[0]
    assert s.min() >= 0 and s.max() < K


@pytest.mark.parametrize("tag", ["cfg1", "small", "one", "none"])
def test_estimator_oracle_matches_reference(golden, tag):
    g = golden("estimator.npz").case(tag)
    s64, t64 = er.scale_shift_ls_ref(g["pred"], g["target"], g["mask"], np.float64)
    s32, t32 = er.scale_shift_ls_ref(g["pred"], g["target"], g["mask"], np.float32)
    if tag in ("one", "none"):          # det <= 0 -> (0, 0)   (estimator.py:22-27)
        assert s64 == 0 and t64 == 0 and g["scale64"] == 0 and g["shift64"] == 0
    else:
        np.testing.assert_allclose(s64, g["scale64"], rtol=1e-12)
        np.testing.assert_allclose(t64, g["shift64"], rtol=1e-10)
        np.testing.assert_allclose(s32, g["scale32"], rtol=1e-6)
        np.testing.assert_allclose(t32, g["shift32"], rtol=1e-4)
    out = er.apply_clamp_ref(g["pred"], np.float32(g["scale64"]), np.float32(g["shift64"]), 0.1, 255.0)
    if "out_clamped_from64" in g:
        assert np.array_equal(out, g["out_clamped_from64"])
        assert np.array_equal(er.apply_clamp_ref(g["pred"], g["scale32"], g["shift32"], 0.1, 255.0), g["out_clamped32"])
    else:
        assert np.array_equal(out[::37, ::41], g["out_clamped_from64_probe"])
        assert np.sum(out, dtype=np.float64) == float(g["out_clamped_from64_sum"])
    if "opt_scale" in g:
        np.testing.assert_allclose(er.optimize_scale_ref(g["pred"], g["target"], g["mask"]), g["opt_scale"], rtol=1e-9)


@pytest.mark.parametrize("tag", ["n40", "n3", "n200"])
def test_interpolator_oracle_matches_reference(golden, tag):
    g = golden("interpolator.npz").case(tag)
    coords, scales, shifts = ir.extract_knots_ref(g["pred"], g["sparse"], g["valid"])
    assert np.array_equal(coords, g["knot_coords"])
    assert np.array_equal(scales, g["knot_scales"]) and np.array_equal(shifts, g["knot_shifts"])
    m = ir.interpolated_scale_map_ref(g["pred"], g["sparse"], g["valid"])
    assert m.dtype == np.float32 and np.array_equal(m, g["scale_map"])


def _tiny_model(g):
    from riders_b200.rcnet.rcnet_model import RCNetModel
    model = RCNetModel(3, 3, [64, 32], ["rcnet", "batch_norm"], [8, 16, 16, 16, 16], [8, 16, 16, 16, 16],
                       ["multiscale", "batch_norm"], [32, 16, 16, 8, 8], device=torch.device("cpu"))
    enc = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("enc/")}
    dec = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("dec/")}
    model.encoder.load_state_dict(enc)          # strict: parameter names match the reference's
    model.decoder.load_state_dict(dec)
    model.eval()
    return model


def test_rcnet_oracle_matches_reference_end_to_end(golden):
    g = golden("rcnet_tiny.npz")
    model = _tiny_model(g)
    image, pts, boxes = torch.from_numpy(g["image"]), torch.from_numpy(g["pts"]), [torch.from_numpy(g["boxes"])]
    padded = torch.nn.functional.pad(image, (16, 16, 32, 32), mode="replicate")
    logits = rcnet_ref.model_forward_ref(model, padded, pts, boxes)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    for literal in (True, False):
        d, r, _ = rcnet_ref.forward_output_ref(model, image, pts, boxes, float(g["thr"]), literal=literal)
        np.testing.assert_allclose(r.numpy(), g["response"], rtol=1e-5, atol=1e-6)
        agree = (r.numpy() > 0) == (g["response"] > 0)
        np.testing.assert_allclose(d.numpy()[agree], g["depth"][agree], rtol=1e-4, atol=1e-5)
        assert agree.mean() > 0.999


# ---------------------------------------------------------------------------------------------
# N2 / N4 rows: oracle.sml_glue_ref against the reference's own transforms / utils
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["zju", "small"])
def test_sml_input_oracle_matches_the_reference_transform_pipeline(golden, tag):
    from oracle import sml_glue_ref as sg
    from oracle.gen_golden import sml_input_case
    g = golden("sml_glue.npz").case("sml_input_" + tag)
    H, W = (int(v) for v in g["shape"])
    image, int_depth, int_scales = sml_input_case(tag, H, W)
    x, d = sg.sml_input_ref(image, int_depth, int_scales)
    assert d.shape == g["d"].shape and np.array_equal(d, g["d"])              # cv2.INTER_NEAREST index tables
    ref_x = g["x"]
    assert np.array_equal(x if tag != "zju" else x[:, ::4, ::4], ref_x)       # float32 statements, bit for bit
    assert sg.sml_resize_size_ref(640, 512) == (352, 288)


def test_sml_nearest_indices_equal_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import sml_glue_ref as sg
    for src, dst in ((640, 352), (512, 288), (1280, 352), (200, 608), (96, 288), (37, 64)):
        ramp = np.arange(src, dtype=np.float32)[None, :].repeat(2, 0)
        got = cv2.resize(ramp, (dst, 2), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
        assert np.array_equal(sg.nearest_indices_ref(dst, src), got), (src, dst)


@pytest.mark.parametrize("tag", ["u8", "unit"])
def test_rcnet_prologue_oracle_matches_the_reference_statements(golden, tag):
    from oracle import sml_glue_ref as sg
    g = golden("sml_glue.npz").case("prologue_" + tag)
    padded, pts, boxes = sg.rcnet_prologue_ref(g["image"], g["points"], g["patch"])
    assert np.array_equal(padded, g["padded"]) and np.array_equal(pts, g["points_shifted"]) and np.array_equal(boxes, g["boxes"])


def test_sml_glue_statements_are_pinned_by_the_reference_utils(golden):
    from oracle import sml_glue_ref as sg
    z = golden("sml_glue.npz")
    n = z.case("normalize")
    assert np.array_equal(sg.normalize_unit_range_ref(n["in"].copy()), n["out"])
    m = z.case("metrics")
    np.testing.assert_allclose(sg.depth_metrics_ref(m["pred"], m["gt"], 0.0, 50.0), m["values"], rtol=1e-12)
