"""ctypes binding of the C-ABI in include/disco_b200.h.  No CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdisco_b200.so")

PREC_FP16 = 0
PREC_BF16X3 = 1
OUT_ACT = 0
OUT_F32 = 1


class DiscoError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """struct disco_conv_desc (include/disco_b200.h)."""
    _fields_ = [
        ("src", C.c_void_p * 2),
        ("src_lo_off", C.c_longlong * 2),
        ("src_c", C.c_int * 2),
        ("src_up", C.c_int * 2),
        ("n", C.c_int), ("h_in", C.c_int), ("w_in", C.c_int),
        ("h_out", C.c_int), ("w_out", C.c_int),
        ("stride", C.c_int), ("taps", C.c_int), ("c_blk", C.c_int),
        ("c_out", C.c_int), ("block_n", C.c_int),
        ("wpack", C.c_void_p), ("wpack_stacked", C.c_int), ("wref", C.c_void_p), ("bias", C.c_void_p),
        ("relu", C.c_int), ("precision", C.c_int),
        ("out_mode", C.c_int),
        ("out", C.c_void_p * 2),
        ("out_lo_off", C.c_longlong),
        ("out_split", C.c_int),
        ("chain_wpack", C.c_void_p), ("chain_bias", C.c_void_p), ("chain_c_out", C.c_int), ("chain_relu", C.c_int),
        ("src_lo_nonzero", C.c_void_p),
        ("subpix", C.c_int), ("sub_py", C.c_int), ("sub_px", C.c_int),
    ]


class FusionDesc(C.Structure):
    """struct disco_fusion_desc (include/disco_b200.h)."""
    _fields_ = [
        ("feat_hi", C.c_void_p), ("feat_lo_off", C.c_longlong), ("precision", C.c_int),
        ("en", C.c_void_p), ("hid", C.c_int),
        ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("w4", C.c_void_p), ("b4", C.c_void_p),
        ("trans", C.c_void_p), ("num_agent", C.c_void_p),
        ("B", C.c_int), ("A", C.c_int), ("h", C.c_int), ("w", C.c_int), ("C", C.c_int),
        ("only_v2i", C.c_int), ("trans_scale", C.c_float),
        ("out_hi", C.c_void_p), ("out_lo_off", C.c_longlong),
        ("weights", C.c_void_p),
        ("row_begin", C.c_int), ("row_end", C.c_int),
        ("outage", C.c_void_p),
        ("wpre", C.c_void_p),
    ]


class GradSrc(C.Structure):
    """struct disco_grad_src."""
    _fields_ = [("ptr", C.c_void_p), ("c_total", C.c_int), ("c_off", C.c_int), ("pool", C.c_int)]


class BnDesc(C.Structure):
    """struct disco_bn_desc (training-mode BatchNorm forward/backward)."""
    _fields_ = [
        ("z", C.c_void_p), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
        ("momentum", C.c_float), ("eps", C.c_float),
        ("sums", C.c_void_p), ("stats", C.c_void_p),
        ("out_hi", C.c_void_p), ("out_lo_off", C.c_longlong), ("relu", C.c_int),
        ("g", GradSrc * 3), ("n_g", C.c_int),
        ("dz_hi", C.c_void_p), ("dz_lo_off", C.c_longlong),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
    ]


class PackDesc(C.Structure):
    """struct disco_pack_desc."""
    _fields_ = [
        ("w", C.c_void_p), ("co_src", C.c_int), ("ci_src", C.c_int), ("taps", C.c_int),
        ("transpose", C.c_int), ("c0", C.c_int), ("n_real", C.c_int), ("k_pad", C.c_int),
        ("block_n", C.c_int), ("c_blk", C.c_int), ("n_tiles", C.c_int), ("stacked", C.c_int),
        ("wpack", C.c_void_p), ("bias_src", C.c_void_p), ("bias", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    """struct disco_wgrad_desc."""
    _fields_ = [
        ("src", C.c_void_p * 2), ("src_lo_off", C.c_longlong * 2), ("src_c", C.c_int * 2), ("src_up", C.c_int * 2),
        ("n", C.c_int), ("h_in", C.c_int), ("w_in", C.c_int), ("h_out", C.c_int), ("w_out", C.c_int),
        ("stride", C.c_int), ("taps", C.c_int),
        ("dz_hi", C.c_void_p), ("dz_lo_off", C.c_longlong), ("c_out", C.c_int),
        ("partial", C.c_void_p), ("splits", C.c_int), ("dw", C.c_void_p), ("c_in_real", C.c_int), ("passes", C.c_int),
    ]


class PwfTrainDesc(C.Structure):
    """struct disco_pwf_train_desc."""
    _fields_ = [
        ("feat_hi", C.c_void_p), ("feat_lo_off", C.c_longlong), ("en", C.c_void_p), ("hid", C.c_int),
        ("g1", C.c_void_p), ("be1", C.c_void_p),
        ("w2", C.c_void_p), ("b2", C.c_void_p), ("g2", C.c_void_p), ("be2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p), ("g3", C.c_void_p), ("be3", C.c_void_p),
        ("w4", C.c_void_p), ("b4", C.c_void_p),
        ("eps", C.c_float), ("momentum", C.c_float),
        ("rm1", C.c_void_p), ("rv1", C.c_void_p), ("rm2", C.c_void_p), ("rv2", C.c_void_p), ("rm3", C.c_void_p), ("rv3", C.c_void_p),
        ("nbt1", C.c_void_p), ("nbt2", C.c_void_p), ("nbt3", C.c_void_p),
        ("trans", C.c_void_p), ("num_agent", C.c_void_p), ("outage", C.c_void_p),
        ("B", C.c_int), ("A", C.c_int), ("h", C.c_int), ("w", C.c_int), ("C", C.c_int),
        ("only_v2i", C.c_int), ("trans_scale", C.c_float),
        ("psum", C.c_void_p), ("wlogit", C.c_void_p),
        ("dfused", C.c_void_p), ("dwlogit", C.c_void_p), ("dfeat", C.c_void_p), ("den", C.c_void_p), ("gsum", C.c_void_p), ("dparams", C.c_void_p),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "disco_version": (C.c_int, []),
    "disco_last_error": (C.c_int, [C.c_char_p, C.c_size_t]),
    "disco_device_check": (C.c_int, []),
    "disco_conv_forward": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "disco_conv_reference": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "disco_conv_smem_bytes": (C.c_int, [C.POINTER(ConvDesc)]),
    "disco_bev_pack": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "disco_act_unpack_nchw": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "disco_voxelize_occupy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "disco_voxelize_occupy_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "disco_bev_scatter": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_void_p]),
    "disco_bev_scatter_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_longlong,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "disco_fusion_forward": (C.c_int, [C.POINTER(FusionDesc), C.c_void_p]),
    "disco_det_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_float, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "disco_nms_workspace_bytes": (C.c_longlong, [C.c_int, C.c_int]),
    "disco_nms_rotated": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                    C.c_double, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "disco_corner_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    # segmentation U-Net data movement
    "disco_maxpool2": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "disco_upsample_bilinear2x": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "disco_maxpool2_backward": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "disco_upsample_bilinear2x_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "disco_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    # training mode
    "disco_bn_train_forward": (C.c_int, [C.POINTER(BnDesc), C.c_void_p]),
    "disco_bn_train_backward": (C.c_int, [C.POINTER(BnDesc), C.c_void_p]),
    "disco_pack_weights": (C.c_int, [C.POINTER(PackDesc), C.c_void_p]),
    "disco_grad_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "disco_channel_sum": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "disco_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "disco_add_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "disco_kd_kl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "disco_focal_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_float, C.c_int, C.c_void_p,
                                   C.c_longlong, C.c_void_p, C.c_void_p]),
    "disco_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "disco_conv_wgrad_reference": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "disco_conv_wgrad_splits": (C.c_int, [C.POINTER(WgradDesc)]),
    "disco_pwf_train_forward": (C.c_int, [C.POINTER(PwfTrainDesc), C.c_void_p]),
    "disco_fusion_combine_backward": (C.c_int, [C.POINTER(PwfTrainDesc), C.c_void_p]),
    "disco_pwf_train_backward": (C.c_int, [C.POINTER(PwfTrainDesc), C.c_void_p]),
}

ABI_VERSION = 200    # == disco_version() of the library these ctypes structures describe (csrc/capi.cu)
_lib = None


def load():
    """Load libdisco_b200.so (built by `python -m disconet_b200.build`).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DiscoError(
            f"{LIB_PATH} is missing: build it with `python -m disconet_b200.build` "
            "(or __graft_entry__.build()).  disconet_b200 has no CPU / PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if an exported symbol is missing
        fn.restype = res
        fn.argtypes = args
    got = lib.disco_version()
    if got != ABI_VERSION:
        raise DiscoError(f"{LIB_PATH} reports ABI version {got}, this package needs {ABI_VERSION}: the library is stale -- rebuild it "
                         "with `python -m disconet_b200.build --force`")
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().disco_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc < 0:
        raise DiscoError(f"{what or 'libdisco_b200'} failed (code {rc}): {last_error()}")
