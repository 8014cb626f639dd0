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
    """Returns the dict name -> module that was installed.  With ``patch_rcnet_main`` an already
    imported (or importable) reference ``rcnet_main`` gets its ``forward_output`` replaced, so
    ``rcnet_main.validate`` (RCNet/rcnet_main.py:558) uses the reduce kernel too."""
    import importlib
    done = {}
    for alias, target in _NAMES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        done[alias] = mod
    from .rcnet import rcnet_main as ours
    if patch_rcnet_main:
        ref = sys.modules.get("rcnet_main")
        if ref is not None and ref is not ours:
            ref.forward_output = ours.forward_output
        else:
            sys.modules["rcnet_main"] = ours
        done["rcnet_main.forward_output"] = ours.forward_output
    return done


def uninstall():
    for alias, target in _NAMES.items():
        if sys.modules.get(alias) is sys.modules.get(target):
            del sys.modules[alias]
