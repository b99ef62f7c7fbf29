"""ctypes binding of libuorf_b200.so (the C ABI declared in include/uorf_b200.h).

There is no fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libuorf_b200.so")

UORF_OK = 0
PREC_BF16, PREC_FP16, PREC_BF16X3, PREC_FP16X3 = 1, 2, 3, 4
PRECISIONS = {"bf16": PREC_BF16, "fp16": PREC_FP16, "bf16x3": PREC_BF16X3, "fp16x3": PREC_FP16X3}

c_float_p = C.POINTER(C.c_float)


class UorfError(RuntimeError):
    pass


class Frustum(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("D", C.c_int32),
                ("near_plane", C.c_float), ("far_plane", C.c_float),
                ("spixel2cam", C.c_float * 16), ("nss_scale", C.c_float)]


class DecoderWeights(C.Structure):
    _fields_ = [("f_before_w", C.c_void_p * 3), ("f_before_b", C.c_void_p * 3),
                ("f_after_w", C.c_void_p * 3), ("f_after_b", C.c_void_p * 3),
                ("f_after_latent_w", C.c_void_p), ("f_after_latent_b", C.c_void_p),
                ("f_after_shape_w", C.c_void_p), ("f_after_shape_b", C.c_void_p),
                ("f_color0_w", C.c_void_p), ("f_color0_b", C.c_void_p),
                ("f_color2_w", C.c_void_p), ("f_color2_b", C.c_void_p),
                ("b_before_w", C.c_void_p * 3), ("b_before_b", C.c_void_p * 3),
                ("b_after_w", C.c_void_p * 4), ("b_after_b", C.c_void_p * 4)]


class DecoderArgs(C.Structure):
    _fields_ = [("coor_bg", C.c_void_p), ("coor_fg", C.c_void_p), ("fg_slot_stride", C.c_int64),
                ("fg_transform", C.c_void_p), ("noise", C.c_void_p),
                ("dens_noise", C.c_float), ("locality_ratio", C.c_float),
                ("locality", C.c_int32), ("fixed_locality", C.c_int32),
                ("P", C.c_int64), ("K", C.c_int32), ("precision", C.c_int32),
                ("image", C.c_void_p), ("workspace", C.c_void_p),
                ("raws", C.c_void_p), ("masked_raws", C.c_void_p), ("unmasked_raws", C.c_void_p),
                ("masks", C.c_void_p), ("outside", C.c_void_p), ("fg_slot_offset", C.c_void_p)]


class DecoderBwdArgs(C.Structure):
    _fields_ = [("coor_bg", C.c_void_p), ("coor_fg", C.c_void_p), ("fg_slot_stride", C.c_int64),
                ("fg_transform", C.c_void_p), ("noise", C.c_void_p),
                ("dens_noise", C.c_float), ("locality_ratio", C.c_float),
                ("locality", C.c_int32), ("fixed_locality", C.c_int32),
                ("P", C.c_int64), ("K", C.c_int32), ("Z", C.c_int32),
                ("z_slots", C.c_void_p), ("image", C.c_void_p), ("unmasked_raws", C.c_void_p), ("grad_raws", C.c_void_p),
                ("workspace", C.c_void_p), ("grad_z_slots", C.c_void_p)]


_SA_FIELDS = ["slots_mu", "slots_logsigma", "slots_mu_bg", "slots_logsigma_bg",
              "norm_feat_w", "norm_feat_b", "to_k_w", "to_v_w",
              "to_q_ln_w", "to_q_ln_b", "to_q_w", "to_q_bg_ln_w", "to_q_bg_ln_b", "to_q_bg_w",
              "gru_w_ih", "gru_w_hh", "gru_b_ih", "gru_b_hh",
              "gru_bg_w_ih", "gru_bg_w_hh", "gru_bg_b_ih", "gru_bg_b_hh",
              "to_res_ln_w", "to_res_ln_b", "to_res_w1", "to_res_b1", "to_res_w2", "to_res_b2",
              "to_res_bg_ln_w", "to_res_bg_ln_b", "to_res_bg_w1", "to_res_bg_b1", "to_res_bg_w2", "to_res_bg_b2"]


class SlotAttentionWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _SA_FIELDS]


# state-dict key of each SlotAttentionWeights field (reference models/model.py:165-203)
SA_STATE_KEYS = {
    "slots_mu": "slots_mu", "slots_logsigma": "slots_logsigma", "slots_mu_bg": "slots_mu_bg",
    "slots_logsigma_bg": "slots_logsigma_bg", "norm_feat_w": "norm_feat.weight", "norm_feat_b": "norm_feat.bias",
    "to_k_w": "to_k.weight", "to_v_w": "to_v.weight",
    "to_q_ln_w": "to_q.0.weight", "to_q_ln_b": "to_q.0.bias", "to_q_w": "to_q.1.weight",
    "to_q_bg_ln_w": "to_q_bg.0.weight", "to_q_bg_ln_b": "to_q_bg.0.bias", "to_q_bg_w": "to_q_bg.1.weight",
    "gru_w_ih": "gru.weight_ih", "gru_w_hh": "gru.weight_hh", "gru_b_ih": "gru.bias_ih", "gru_b_hh": "gru.bias_hh",
    "gru_bg_w_ih": "gru_bg.weight_ih", "gru_bg_w_hh": "gru_bg.weight_hh", "gru_bg_b_ih": "gru_bg.bias_ih",
    "gru_bg_b_hh": "gru_bg.bias_hh",
    "to_res_ln_w": "to_res.0.weight", "to_res_ln_b": "to_res.0.bias", "to_res_w1": "to_res.1.weight",
    "to_res_b1": "to_res.1.bias", "to_res_w2": "to_res.3.weight", "to_res_b2": "to_res.3.bias",
    "to_res_bg_ln_w": "to_res_bg.0.weight", "to_res_bg_ln_b": "to_res_bg.0.bias", "to_res_bg_w1": "to_res_bg.1.weight",
    "to_res_bg_b1": "to_res_bg.1.bias", "to_res_bg_w2": "to_res_bg.3.weight", "to_res_bg_b2": "to_res_bg.3.bias",
}

class EncoderPrepLayer(C.Structure):      # struct uorf_encoder_prep_layer
    _fields_ = [("weight", C.c_void_p), ("fwd_t", C.c_void_p), ("dgrad_t", C.c_void_p), ("c_out", C.c_int32), ("c_in", C.c_int32)]


_SIGNATURES = {
    "uorf_abi_version": (C.c_int, []),
    "uorf_last_error": (C.c_char_p, []),
    "uorf_init": (C.c_int, [C.c_int]),
    "uorf_launch_count": (C.c_int64, []),
    "uorf_debug_flag": (C.c_int, []),
    "uorf_encoder_conv3x3": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "uorf_encoder_conv3x3_dgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_void_p, C.c_void_p]),
    "uorf_encoder_upsample2x_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "uorf_encoder_wgrad_workspace_floats": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "uorf_encoder_conv3x3_wgrad": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_conv3x3_stem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "uorf_conv3x3_stem_dgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p]),
    "uorf_encoder_weight_prep": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "uorf_conv_split_pool2x2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "uorf_conv_split_unpool2x2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p]),
    "uorf_invert_small": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "uorf_image_loss_forward": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32] + [C.c_void_p] * 7),
    "uorf_image_loss_backward": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "uorf_feature_mse_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_feature_mse_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "uorf_adam_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "uorf_conv_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "uorf_conv3x3_workspace_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "uorf_conv3x3_x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "uorf_sample_coords": (C.c_int, [C.POINTER(Frustum), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_sample_coords_window": (C.c_int, [C.POINTER(Frustum), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_decoder_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "uorf_decoder_prepare": (C.c_int, [C.POINTER(DecoderWeights), C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "uorf_decoder_forward_workspace_floats": (C.c_size_t, [C.c_int64, C.c_int32]),
    "uorf_decoder_forward": (C.c_int, [C.POINTER(DecoderArgs), C.c_void_p]),
    "uorf_decoder_backward_workspace_floats": (C.c_size_t, [C.c_int64, C.c_int32]),
    "uorf_decoder_backward": (C.c_int, [C.POINTER(DecoderWeights), C.POINTER(DecoderWeights), C.POINTER(DecoderBwdArgs), C.c_void_p]),
    "uorf_composite_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_composite_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                          C.c_void_p]),
    "uorf_slot_attention_workspace_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "uorf_slot_attention_forward": (C.c_int, [C.POINTER(SlotAttentionWeights), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_slot_attention_backward_workspace_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "uorf_slot_attention_backward": (C.c_int, [C.POINTER(SlotAttentionWeights), C.POINTER(SlotAttentionWeights), C.c_void_p, C.c_int64,
                                               C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}
_OPTIONAL = {}
# libuorf_b200_probe.so (include/uorf_b200_probe.h): test / diagnostic hooks, a separate library loaded only by tests/ and scripts/
PROBE_LIB_PATH = os.path.join(_HERE, "_build", "libuorf_b200_probe.so")
_PROBE_SIGNATURES = {
    "uorf_probe_gemm_tn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uorf_probe_timing": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "uorf_probe_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (python -m uorf_b200.build)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                # UORF_B200_LIB: a tuning variant of the same library (uorf_b200/build.py, --tag); never a different backend
                path = os.environ.get("UORF_B200_LIB") or LIB_PATH
                if not os.path.exists(path):
                    raise UorfError(f"{path} not found: build it with `python -m uorf_b200.build` "
                                    "(uorf_b200 has no CPU or PyTorch fallback)")
                l = C.CDLL(path)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype, fn.argtypes = res, args
                for name, (res, args) in _OPTIONAL.items():
                    if hasattr(l, name):
                        fn = getattr(l, name)
                        fn.restype, fn.argtypes = res, args
                _lib = l
    return _lib


_probe = None


def probe_lib() -> C.CDLL:
    """The test-hook library (tcgen05 layout probes, hand-off timing); built next to the product library."""
    global _probe
    if _probe is None:
        with _lock:
            if _probe is None:
                if not os.path.exists(PROBE_LIB_PATH):
                    raise UorfError(f"{PROBE_LIB_PATH} not found: build it with `python -m uorf_b200.build`")
                l = C.CDLL(PROBE_LIB_PATH)
                for name, (res, args) in _PROBE_SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype, fn.argtypes = res, args
                _probe = l
    return _probe


def launch_count() -> int:
    """kernels launched by libuorf_b200.so since it was loaded (uorf_launch_count)"""
    return int(lib().uorf_launch_count())


def check(rc: int, what: str) -> None:
    if rc != UORF_OK:
        msg = lib().uorf_last_error().decode(errors="replace")
        raise UorfError(f"{what} failed ({rc}): {msg}")


def exported_symbols():
    return list(_SIGNATURES) + list(_OPTIONAL)
