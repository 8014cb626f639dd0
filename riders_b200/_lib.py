"""ctypes binding of ``libriders_b200.so`` (the C ABI declared in include/riders_b200.h).

There is deliberately no fallback: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C riders_b200/csrc``)
every op raises.  PyTorch only supplies device memory and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libriders_b200.so")

RIDERS_OK = 0
RIDERS_EINVAL, RIDERS_ECUDA, RIDERS_ENODEVICE, RIDERS_EQHULL, RIDERS_EINTERNAL = -1, -2, -3, -4, -5
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
LOFTR_SIMT, LOFTR_TCGEN05 = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
MAX_LEVELS = 8

vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class LoftrWeights(C.Structure):
    """riders_loftr_weights (include/riders_b200.h)."""
    _fields_ = [(n, vp) for n in ("wq", "wk", "wv", "wm", "w1", "w2", "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


# name -> (restype, argtypes); must list every symbol of include/riders_b200.h
SIGNATURES = {
    "riders_last_error_string": (C.c_char_p, []),
    "riders_abi_version": (ci, []),
    "riders_device_arch": (ci, []),
    "riders_launch_count": (C.c_ulonglong, []),
    "riders_roi_gather_fwd": (ci, [ci, vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, ci, ci, ci, vp, vp, vp]),
    "riders_roi_gather_fwd_placed": (ci, [ci, vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, ci, ci, ci, vp, vp, vp, vp]),
    "riders_roi_gather_bwd": (ci, [ci, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, ci, vp, vp]),
    "riders_linear_attention_fwd": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, vp, vp]),
    "riders_loftr_workspace_bytes": (sz, [ci, ci, ci, ci]),
    "riders_loftr_impl_available": (ci, [ci, ci, ci, ci]),
    "riders_loftr_layer_fwd": (ci, [vp, vp, C.POINTER(LoftrWeights), ci, ci, ci, ci, ci, cf, cf, ci, vp, vp, vp]),
    "riders_loftr_transformer_fwd": (ci, [vp, vp, C.POINTER(LoftrWeights), C.POINTER(ci), ci,
                                          ci, ci, ci, ci, ci, cf, cf, ci, vp, vp, vp, vp]),
    "riders_loftr_saved_bytes": (sz, [ci, ci, ci, ci]),
    "riders_loftr_train_workspace_bytes": (sz, [ci, ci, ci, ci]),
    "riders_loftr_layer_fwd_train": (ci, [vp, vp, C.POINTER(LoftrWeights), ci, ci, ci, ci, ci, cf, cf, ci, vp, vp, vp, vp]),
    "riders_loftr_layer_bwd": (ci, [vp, vp, C.POINTER(LoftrWeights), vp, vp, ci, ci, ci, ci, ci, cf, cf, ci,
                                    vp, vp, C.POINTER(LoftrWeights), vp, vp]),
    "riders_confidence_reduce": (ci, [vp, ci, vp, vp, ci, ci, ci, ci, ci, cf, vp, vp, vp, vp]),
    "riders_lsq_workspace_bytes": (sz, [ci, ci, ci]),
    "riders_lsq_align": (ci, [vp, vp, vp, ci, ci, ci, vp, vp, vp, cf, cf, ci, vp, vp]),
    "riders_l1_scale_workspace_bytes": (sz, [ci, ci]),
    "riders_l1_scale_align": (ci, [vp, vp, vp, ci, ci, ci, ci, C.c_double, C.c_double, C.c_double, ci,
                                   vp, vp, vp, vp, cf, cf, ci, vp, vp]),
    "riders_glue_workspace_bytes": (sz, [ci, ci, ci]),
    "riders_scale_scaffold": (ci, [vp, vp, vp, vp, ci, ci, ci, cf, cf, ci, vp, vp, vp, vp, vp]),
    "riders_depth_metrics": (ci, [vp, vp, ci, ci, ci, cf, cf, vp, vp, vp]),
    "riders_extract_knots": (ci, [vp, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, vp]),
    "riders_interp_workspace_bytes": (sz, [ci, ci, ci, ci]),
    "riders_interp_nearest": (ci, [vp, vp, vp, ci, ci, ci, ci, cf, vp, vp]),
    "riders_interp_triangulate": (ci, [vp, vp, vp, ci, ci, vp, vp, vp, vp, vp]),
    "riders_interp_fill": (ci, [ci, ci, ci, ci, cf, vp, vp, vp]),
    "riders_upsample_nearest_nhwc": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, vp]),
    "riders_leaky_concat_nhwc": (ci, [vp, ci, vp, ci, C.c_longlong, cf, ci, vp, vp]),
    "riders_bias_leaky_concat_nhwc": (ci, [vp, vp, ci, vp, ci, C.c_longlong, cf, vp, vp]),
    "riders_bias_leaky_place_nhwc": (ci, [vp, vp, ci, vp, ci, C.c_longlong, cf, vp, ci, vp]),
    "riders_bias_act_nhwc": (ci, [vp, vp, ci, C.c_longlong, cf, vp, vp]),
    "riders_conv3x3_head_nhwc": (ci, [vp, ci, ci, ci, ci, vp, cf, vp, vp, vp]),
    "riders_conv3x3_head_sigmoid_nhwc": (ci, [vp, ci, ci, ci, ci, vp, cf, vp, ci, vp, vp]),
    "riders_d2s_bias_act_concat_nhwc": (ci, [vp, vp, ci, vp, ci, ci, ci, ci, cf, vp, vp]),
    "riders_d2s_bias_act_place_nhwc": (ci, [vp, vp, ci, vp, ci, ci, ci, ci, cf, vp, ci, vp]),
    "riders_d2s_shift_bias_act_place_nhwc": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, cf, vp, ci, vp]),
    "riders_subpixel_d2s_place_nhwc": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, cf, vp, ci, vp]),
    "riders_subpixel_d2s_shift_place_nhwc": (ci, [vp, ci, vp, vp, ci, ci, ci, ci, ci, ci, cf, vp, ci, vp]),
    "riders_bias_act_pad_nhwc": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, cf, vp, vp]),
    "riders_decoder_tail_nhwc": (ci, [vp, ci, ci, ci, ci, vp, vp, vp, vp, cf, ci, vp, vp]),
    "riders_decoder_tail_tc_nhwc": (ci, [vp, ci, ci, ci, ci, vp, vp, vp, cf, ci, vp, vp, vp]),
    "riders_subpixel_conv_tc_half_nhwc": (ci, [vp, vp, ci, ci, ci, vp, cf, vp, vp, vp]),
    "riders_decoder_tail_tc_half_nhwc": (ci, [vp, ci, ci, ci, vp, vp, cf, ci, vp, vp, vp]),
    "riders_upconv_tc_half_nhwc": (ci, [vp, vp, cf, vp, ci, ci, ci, vp, cf, vp, vp, vp]),
    "riders_upconv_tc_f32_nhwc": (ci, [vp, vp, cf, vp, ci, ci, ci, vp, cf, vp, ci, vp, vp]),
    "riders_conv3x3_tc_f32_nhwc": (ci, [vp, vp, cf, vp, ci, ci, ci, vp, cf, vp, ci, vp, vp]),
    "riders_subpixel_conv_act_nhwc": (ci, [vp, vp, cf, vp, ci, ci, ci, vp, vp]),
    "riders_subpixel_conv_tc_nhwc": (ci, [vp, vp, ci, ci, ci, vp, vp]),
    "riders_subpixel_conv_smem_bytes": (C.c_size_t, [ci]),
    "riders_debug_mma_sync_rate": (ci, [ci, ci, ci, vp, vp, vp]),
    "riders_debug_set": (ci, [C.c_char_p, ci]),
    "riders_rcnet_prologue": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, vp, vp, vp]),
    "riders_sml_input_assemble": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, cf, cf, cf, cf, vp, vp, vp]),
    "riders_debug_loftr_trace": (ci, [vp]),
    "riders_debug_tc_gemm": (ci, [vp, vp, vp, ci, ci, vp]),
    "riders_debug_tc_mma_rate": (ci, [ci, ci, ci, vp, vp]),
    "riders_interp_linear": (ci, [vp, vp, vp, ci, ci, ci, ci, cf, vp, vp, vp, vp, vp, vp]),
}

_lib = None


class RidersError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "riders_b200: %s is missing - build the CUDA extension first "
                "(`make -C riders_b200/csrc` or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so is stale
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def last_error():
    return lib().riders_last_error_string().decode("utf-8", "replace")


def check(rc, what):
    if rc != RIDERS_OK:
        raise RidersError("%s failed (code %d): %s" % (what, rc, last_error()))
