"""The mirrors resolve under the reference's import names and keep its class / parameter names
(train_zju.py, val_zju.py, run_rcnet_zju.py import by bare module name)."""
import sys

import torch


def test_install_registers_reference_module_names():
    from riders_b200 import dropin
    done = dropin.install()
    try:
        import linear_attention, networks, rcnet_model                  # noqa: E401
        from modules.estimator import LeastSquaresEstimator, compute_scale_and_shift_ls   # noqa: F401
        from modules.interpolator import Interpolator2D, interpolate_knots                 # noqa: F401
        from rcnet_main import forward_output                                              # noqa: F401
        assert networks.RCNetEncoder.__module__.startswith("riders_b200")
        assert "rcnet_main.forward_output" in done
    finally:
        dropin.uninstall()
        sys.modules.pop("rcnet_main", None)


def test_mirror_keeps_reference_parameter_names_and_counts():
    from riders_b200.rcnet.rcnet_model import RCNetModel
    m = RCNetModel(3, 3, [240, 100], ["rcnet", "batch_norm"], [32, 64, 128, 128, 128], [32, 64, 128, 128, 128],
                   ["multiscale", "batch_norm"], [256, 128, 64, 32, 16], device=torch.device("cpu"))
    n = sum(p.numel() for p in m.parameters())
    assert n == 5892272                                   # SURVEY section 6 [measured on the reference]
    keys = m.encoder.state_dict().keys()
    for name in ("q_proj", "k_proj", "v_proj", "merge", "mlp.0", "mlp.2"):
        assert "attention.layers.0.%s.weight" % name in keys
    assert "attention.layers.7.norm2.bias" in keys and len(m.encoder.attention.layers) == 8
    assert m.encoder.attention.layer_names == ["self", "cross"] * 4
