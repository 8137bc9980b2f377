"""ctypes binding of libhctr_b200.so (the C ABI declared in include/hctr_b200.h).

There is deliberately no fallback: if the library is missing or the device is not sm_100a the
calls raise.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_uint, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libhctr_b200.so")

HCTR_OK = 0
HCTR_ERR_INVALID = -1
HCTR_ERR_CUDA = -2
HCTR_ERR_UNSUPPORTED = -3
HCTR_ERR_INDEX = -4
HCTR_F32 = 0
HCTR_BF16 = 1

_P = c_void_p
_I = c_int
_L = c_longlong

# name -> (restype, argtypes); mirrors include/hctr_b200.h one to one
SIGNATURES = {
    "hctr_last_error": (c_char_p, []),
    "hctr_abi_version": (_I, []),
    "hctr_device_supported": (_I, [_I]),
    "hctr_stem_conv_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "hctr_conv_bn_act_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "hctr_conv_se_slices": (_I, [_I, _I, _I]),
    "hctr_conv_bn_se_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "hctr_conv_sum_slices": (_I, [_I, _I, _I, _I, _I]),
    "hctr_conv_bn_act_sum_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "hctr_conv_stats_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "hctr_se_gate_workspace_bytes": (_L, [_I, _I]),
    "hctr_se_gate_from_input": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _L, _P]),
    "hctr_conv_bn_gate_res_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "hctr_se_slices": (_I, [_I, _I]),
    "hctr_se_squeeze": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "hctr_se_excite": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _P]),
    "hctr_se_scale_residual_relu": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "hctr_classifier_fwd": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _I, _P]),
    "hctr_classifier_lse_fwd": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _I, _P, _P, _L, _P]),
    "hctr_classifier_lse_workspace_bytes": (_L, [_I, _I, _I]),
    "hctr_classifier_greedy_fwd": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _L, _P]),
    "hctr_classifier_greedy_workspace_bytes": (_L, [_I, _I, _I]),
    "hctr_ctc_greedy_decode": (_I, [_P, _I, _I, _I, _I, _L, _L, _P, _P, _P, _P]),
    "hctr_ctc_collapse": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "hctr_ctc_topk_logsoftmax": (_I, [_P, _I, _I, _I, _I, _L, _L, _I, _P, _P, _P, _P]),
    "hctr_ctc_prefix_beam_search": (_I, [_P, _P, _I, _I, _I, _I, _I, c_double, c_double, _P, _P, _P, _P, _P, _L, _P]),
    "hctr_ctc_beam_workspace_bytes": (_L, [_I, _I, _I]),
    "hctr_ctc_prefix_beam_search_lm": (_I, [_P, _P, _I, _I, _I, _I, _I, c_double, c_double, _P, _P, _P, _P, _P, _P, _L, _P]),
    "hctr_ngram_score": (_I, [_P, _P, _P, _I, _P, _P]),
    "hctr_ctc_skip_beam_search": (_I, [_P, _I, _I, _I, _I, _L, _L, _I, c_double, c_double, _P, _P, _P, _P, _P, _L, _P]),
    "hctr_ctc_skip_beam_search_lm": (_I, [_P, _I, _I, _I, _I, _L, _L, _I, c_double, c_double, _P, _P, _P, _P, _P, _P, _L, _P]),
    "hctr_ctc_skip_workspace_bytes": (_L, [_I, _I]),
    "hctr_ctc_skip_max_candidates": (_I, []),
    "hctr_ctc_skip_workspace_bytes_ex": (_L, [_I, _I, _I]),
    "hctr_ctc_skip_beam_search_ex": (_I, [_P, _I, _I, _I, _I, _L, _L, _I, c_double, c_double, _P, _P, _I, _P, _P, _P, _P, _L, _P]),
    "hctr_ctc_loss_fwd_bwd": (_I, [_P, _I, _I, _I, _I, _L, _L, _P, _P, _P, _I, _P, _P, _P, _P, c_float, _P, _L, _P]),
    "hctr_ctc_loss_workspace_bytes": (_L, [_I, _I, _I]),
    "hctr_ctc_loss_flag_offset": (_L, [_I, _I, _I]),
    "hctr_stat_slices": (_I, [_I, _I, _I]),
    "hctr_chan_stats": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "hctr_bn_finalize_train": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hctr_se_excite_train": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "hctr_train_apply_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, c_float, c_uint, _P]),
    "hctr_train_bwd_reduce": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, c_float, _P]),
    "hctr_train_bwd_finalize": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P,
                                     _P, _P, _P, _P, _P, _P, _P]),
    "hctr_train_bwd_apply": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, c_float, _P]),
    "hctr_conv_dgrad": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "hctr_conv_wgrad": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "hctr_wgrad_workspace_bytes": (_L, [_I, _I, _I, _I, _I, _I]),
    "hctr_classifier_dgrad": (_I, [_P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "hctr_linear_wgrad": (_I, [_P, _L, _P, _P, _I, _I, _I, _I, _I, _P, _L, _P]),
    "hctr_linear_wgrad_workspace_bytes": (_L, [_I, _I, _I, _I, _I]),
    "hctr_colsum_bf16": (_I, [_P, _L, _I, _L, _P, _P, _L, _P]),
    "hctr_colsum_workspace_bytes": (_L, [_L, _I]),
    "hctr_stem_wgrad": (_I, [_P, _P, _P, _I, _I, _I, _P, _L, _P]),
    "hctr_stem_wgrad_workspace_bytes": (_L, [_I, _I, _I]),
    "hctr_sgd_clip_step": (_I, [_P, _P, _P, _L, c_float, c_float, c_float, c_float, c_float, _I, _P, _P, _P]),
    "hctr_sgd_workspace_bytes": (_L, []),
    "hctr_pack_weights": (_I, [_P, _I, _L, _P]),
    "hctr_edit_distance": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _P]),
    "hctr_normalize_pad": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "hctr_resize_area_u8": (_I, [_P, _I, _I, _L, _P, _I, _I, _L, _P]),
}

# include/hctr_b200_testing.h: hooks for tests/, refused by the library unless HCTR_TEST_HOOKS=1
TESTING_SIGNATURES = {
    "hctr_testing_set_conv_variant": (_I, [_I, _I]),
}



class PackDesc(ctypes.Structure):
    """hctr_pack_desc of include/hctr_b200.h."""
    _fields_ = [("src", c_void_p), ("dst_fwd", c_void_p), ("dst_bwd", c_void_p), ("cout", c_int), ("cin", c_int),
                ("taps", c_int), ("bwd_mode", c_int), ("bwd_pitch", c_longlong), ("tile_start", c_longlong)]


_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raise loudly if the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "hctr_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/PyTorch fallback for this path)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(TESTING_SIGNATURES.items()):
            fn = getattr(handle, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    msg = lib().hctr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    """Map C error codes to the exception types the reference would raise."""
    if rc == HCTR_OK:
        return
    msg = "%s: %s" % (what, last_error()) if what else last_error()
    if rc == HCTR_ERR_INDEX:
        raise IndexError(msg)
    if rc == HCTR_ERR_INVALID:
        raise RuntimeError(msg)          # torch raises RuntimeError on shape mismatches
    raise RuntimeError(msg)


c_void_p = c_void_p  # re-exported for callers that offset raw pointers


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
