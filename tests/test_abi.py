"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/riders_b200.h declares, and refuses to compute without a GPU (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "riders_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(riders_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for needed in ("riders_roi_gather_fwd", "riders_linear_attention_fwd", "riders_loftr_layer_fwd",
                   "riders_loftr_transformer_fwd", "riders_confidence_reduce", "riders_lsq_align"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    from riders_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), "libriders_b200.so does not export %s" % name
    # and the ctypes table covers the header, so nothing is bound with guessed argtypes
    assert set(declared_symbols()) == set(_lib.SIGNATURES)
    assert _lib.lib().riders_abi_version() >= 1


def test_signatures_have_no_torch_types():
    text = open(os.path.join(ROOT, "include", "riders_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # prose may mention torch; code may not
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from riders_b200 import _lib, ops
    with pytest.raises(_lib.RidersError):
        ops.confidence_reduce(torch.zeros(1, 4, 4), torch.zeros(1, 3), [0, 1], 8, 8, 0.1)
    with pytest.raises(_lib.RidersError):
        ops.lsq_align(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 4, 4, dtype=torch.bool))
    with pytest.raises(_lib.RidersError):
        ops.roi_pool(torch.zeros(1, 4, 8, 8), [torch.zeros(1, 4)], (2, 2), 1.0)
    assert _lib.lib().riders_device_arch() < 0


def test_argument_validation_returns_einval():
    from riders_b200 import _lib
    L = _lib.lib()
    rc = L.riders_confidence_reduce(None, 0, None, None, 1, 8, 8, 3, 4, 0.1, None, None, None, None)
    assert rc == _lib.RIDERS_EINVAL and "confidence_reduce" in _lib.last_error()
    rc = L.riders_lsq_align(None, None, None, 1, 0, 8, None, None, None, 0.0, 0.0, 0, None, None)
    assert rc == _lib.RIDERS_EINVAL
