"""ctypes binding of the C-ABI in include/speech_inpainting_b200.h.

There is no CPU fallback: if `libsib_b200.so` is missing the import fails loudly
(build it with `python speech-inpainting_b200/build.py` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SIB_LIB_VARIANT=<tag> loads libsib_b200_<tag>.so instead: same-box A/B of two kernel builds (e.g. a baseline build of an
# older tree kept beside the current one); unset in normal use
LIB_PATH = os.path.join(HERE, "libsib_b200" + ("_" + os.environ["SIB_LIB_VARIANT"] if os.environ.get("SIB_LIB_VARIANT") else "") + ".so")

SIB_MAX_TAPS = 128
ACT_NONE, ACT_GELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3


class ConvDesc(C.Structure):
    """mirror of `sib_conv_desc`"""
    _fields_ = [
        ("batch", C.c_int32), ("t_in", C.c_int32), ("t_out", C.c_int32), ("c_in", C.c_int32),
        ("c_out", C.c_int32), ("groups", C.c_int32), ("n_taps", C.c_int32), ("stride", C.c_int32),
        ("tap_offset", C.c_int32 * SIB_MAX_TAPS),
        ("x_batch_stride", C.c_int64), ("y_batch_stride", C.c_int64), ("r_batch_stride", C.c_int64),
        ("x_row_stride", C.c_int32), ("y_row_stride", C.c_int32), ("r_row_stride", C.c_int32),
        ("pre_act", C.c_int32), ("pre_slope", C.c_float), ("post_act", C.c_int32), ("post_slope", C.c_float),
        ("out_scale", C.c_float), ("accumulate", C.c_int32), ("res_after_act", C.c_int32),
        ("act2_slope", C.c_float),
    ]


class ResUnitDesc(C.Structure):
    """mirror of `sib_resunit_desc`"""
    _fields_ = [
        ("batch", C.c_int32), ("t", C.c_int32), ("c", C.c_int32), ("k", C.c_int32), ("dilation", C.c_int32),
        ("accumulate", C.c_int32),
        ("slope_in", C.c_float), ("slope_mid", C.c_float), ("out_scale", C.c_float), ("act2_slope", C.c_float),
        ("x_batch_stride", C.c_int64), ("y_batch_stride", C.c_int64),
        ("x_row_stride", C.c_int32), ("y_row_stride", C.c_int32),
    ]


SIB_LN_SLOTS, SIB_LN_APPLY, SIB_LN_RESIDUAL = 32, 1, 2


class LnFold(C.Structure):
    """mirror of `sib_ln_fold`"""
    _fields_ = [("stats_in", C.c_void_p), ("colsum", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("stats_out", C.c_void_p), ("mode", C.c_int32), ("n_norm", C.c_int32), ("eps", C.c_float)]


class Flow(C.Structure):
    """mirror of `sib_flow`"""
    _fields_ = [("wait", C.c_void_p), ("signal", C.c_void_p), ("wait_target", C.c_int32), ("wait_target_last", C.c_int32)]


class SibError(RuntimeError):
    pass


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGS = {
    "sib_abi_version": ([], C.c_int),
    "sib_last_error": ([], C.c_char_p),
    "sib_launch_count": ([], C.c_longlong),
    "sib_set_pdl": ([_I], _I),
    "sib_conv1d_f32": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P], _I),
    "sib_conv1d_cout1_f32": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P], _I),
    "sib_conv0_f32": ([_I, _P, _I, _I, _L, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P], _I),
    "sib_conv0_num_tiles": ([_I], _I),
    "sib_gn_finalize_f32": ([_P, _I, _I, _I, _I, _F, _P, _P, _P], _I),
    "sib_conv0_gn_stats_f32": ([_P, _I, _I, _L, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P], _I),
    "sib_layernorm_f32": ([_P, _P, _P, _P, _P, _L, _I, _F, _I, _P], _I),
    "sib_attention_f32": ([_P, _P, _P, _I, _I, _I, _I, _P], _I),
    "sib_zero_padded_frames_f32": ([_P, _P, _I, _I, _I, _P], _I),
    "sib_zero_ranges_f32": ([_P, _I, _I, _P, _P, _F, _P], _I),
    "sib_mask_peak_normalize_f32": ([_P, _P, _I, _I, _P, _P, _F, _P], _I),
    "sib_znorm_f32": ([_P, _P, _I, _I, _P, _F, _P], _I),
    "sib_gather_frames_f32": ([_P, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "sib_cos_argmax_f32": ([_P, _P, _I, _I, _I, _P, _P], _I),
    "sib_l2_argmin_f32": ([_P, _P, _I, _I, _I, _P, _P], _I),
    "sib_linear_skinny_f32": ([_P, _P, _P, _P, _I, _I, _I, _P], _I),
    "sib_row_sqnorm_f32": ([_P, _I, _I, _F, _P, _P], _I),
    "sib_row_argmax_f32": ([_P, _I, _I, _P, _P], _I),
    "sib_paste_centroids_f32": ([_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P], _I),
    "sib_extend_mel_f32": ([_P, _P, _I, _I, _I, _I, _I, _P], _I),
    "sib_transpose_f32": ([_P, _P, _I, _I, _I, _P], _I),
    "sib_embed_concat_f32": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "sib_weight_norm_fold_f32": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "sib_pack_int16_f32": ([_P, _P, _L, _P], _I),
    "sib_mel_spectrogram_f32": ([_P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P], _I),
    "sib_mel_workspace_bytes": ([_I], C.c_size_t),
    "sib_resample": ([_P, _I, _I, _I, _L, _P, _P, _I, _I, _I, _I, _P, _I, _L, _P, _P], _I),
    "sib_resample_smem_bytes": ([_I, _I], C.c_size_t),
    "sib_si_sdr_f32": ([_P, _P, _I, _I, _P, _F, _P, _P, _P], _I),
    "sib_si_sdr_workspace_bytes": ([_I], C.c_size_t),
    "sib_abs_diff_mean_f32": ([_P, _P, _I, _L, _P, _P, _P], _I),
    "sib_abs_diff_workspace_bytes": ([_I], C.c_size_t),
    "sib_conv1d_bf16": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P], _I),
    "sib_linear_ln_bf16": ([C.POINTER(ConvDesc), C.POINTER(LnFold), _P, _P, _P, _P, _P, _P], _I),
    "sib_linear_flow_bf16": ([C.POINTER(ConvDesc), C.POINTER(Flow), _P, _P, _P, _P, _P, _P], _I),
    "sib_layernorm_flow_bf16": ([_P, _P, _P, _P, _P, _L, _I, _F, C.POINTER(Flow), _P], _I),
    "sib_attention_flow_bf16": ([_P, _P, _P, _I, _I, _I, _I, C.POINTER(Flow), _P], _I),
    "sib_fill_zero": ([_P, _L, _P], _I),
    "sib_resunit_bf16": ([C.POINTER(ResUnitDesc), _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "sib_resunit_bf16_supported": ([_I, _I, _I, _I, _I], _I),
    "sib_conv1d_bf16_pre_act_supported": ([C.POINTER(ConvDesc), _I, _I], _I),
    "sib_conv1d_bf16_kblock": ([_I, C.POINTER(C.c_int), C.POINTER(C.c_int)], _I),
    "sib_layernorm": ([_P, _I, _P, _I, _P, _P, _P, _I, _L, _I, _F, _I, _P], _I),
    "sib_attention": ([_P, _I, _P, _P, _I, _I, _I, _I, _P], _I),
    "sib_conv0": ([_I, _P, _I, _I, _L, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P], _I),
    "sib_zero_padded_frames": ([_P, _I, _P, _I, _I, _I, _P], _I),
    "sib_conv1d_cout1": ([_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P], _I),
    "sib_extend_mel": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "sib_cast_f32_to_bf16": ([_P, _P, _L, _P], _I),
    "sib_cast_bf16_to_f32": ([_P, _P, _L, _P], _I),
}

_lib = None


def lib():
    """The loaded shared library (raises if it has not been built - no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SibError(f"{LIB_PATH} is missing: build it with `python speech-inpainting_b200/build.py` "
                           "(there is no CPU / PyTorch fallback for this path)")
        l = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the .so is stale
            fn.argtypes = args
            fn.restype = res
        if l.sib_abi_version() != 2:
            raise SibError("libsib_b200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGS)


def check(rc: int, what: str = ""):
    if rc != 0:
        raise SibError(f"{what} failed (status {rc}): {lib().sib_last_error().decode()}")
