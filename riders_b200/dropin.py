"""Install the B200 host mirrors under the reference's module names.

The reference scripts import by bare module name after ``sys.path`` edits
(``import networks`` RCNet/rcnet_model.py:3, ``from linear_attention import LocalFeatureTransformer``
RCNet/networks.py:4, ``from rcnet_main import forward_output`` RCNet/run_rcnet_zju.py:9,
``from modules.estimator import LeastSquaresEstimator`` val_zju.py:19, ``from modules.interpolator import
Interpolator2D`` val_zju.py:20).  ``install()`` registers the mirrors in ``sys.modules`` so those imports
resolve to the CUDA path; the reference's drivers (train / validate loops, data loaders) are untouched.
"""
import sys

_NAMES = {
    "linear_attention": "riders_b200.rcnet.linear_attention",
    "networks": "riders_b200.rcnet.networks",
    "rcnet_model": "riders_b200.rcnet.rcnet_model",
    "modules.estimator": "riders_b200.modules.estimator",
    "modules.interpolator": "riders_b200.modules.interpolator",
}


def install(patch_rcnet_main=True):
    """Returns the dict name -> module that was installed.  With ``patch_rcnet_main`` the REFERENCE's own ``rcnet_main``
    module (already imported, or importable from ``sys.path`` - the scripts run with ``RCNet/`` on the path) keeps all its
    names (``train``, ``validate``, ``log_network_settings``, ``log_evaluation_results`` ... that
    ``run_rcnet_zju.py:9`` / ``train_rcnet_zju.py`` import) and only gets its ``forward_output`` replaced, so
    ``rcnet_main.validate`` (RCNet/rcnet_main.py:558) and the run scripts use the reduce kernel.  The module is never
    swapped for the mirror; only when no reference ``rcnet_main`` can be imported at all does the name resolve to the
    mirror (which then offers ``forward_output`` / ``forward_output_batched`` only)."""
    import importlib
    done = {}
    for alias, target in _NAMES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        done[alias] = mod
    from .rcnet import rcnet_main as ours
    if patch_rcnet_main:
        ref = sys.modules.get("rcnet_main")
        if ref is None:
            try:
                ref = importlib.import_module("rcnet_main")          # the reference's, found through the caller's sys.path
            except Exception:                                         # not importable here (missing deps / not on the path)
                ref = None
        if ref is not None and ref is not ours:
            if not hasattr(ref, "_riders_reference_forward_output"):
                ref._riders_reference_forward_output = getattr(ref, "forward_output", None)
            ref.forward_output = ours.forward_output
            done["rcnet_main"] = ref
        else:
            sys.modules["rcnet_main"] = ours
            done["rcnet_main"] = ours
        done["rcnet_main.forward_output"] = ours.forward_output
    return done


def uninstall():
    for alias, target in _NAMES.items():
        if sys.modules.get(alias) is sys.modules.get(target):
            del sys.modules[alias]
    ref = sys.modules.get("rcnet_main")
    if ref is not None and hasattr(ref, "_riders_reference_forward_output"):
        ref.forward_output = ref._riders_reference_forward_output
        del ref._riders_reference_forward_output
    elif ref is not None and ref.__name__.startswith("riders_b200."):
        del sys.modules["rcnet_main"]
