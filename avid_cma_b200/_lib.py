"""ctypes binding of libavid_b200.so (include/avid_b200.h).

There is deliberately no fallback: if the library cannot be loaded, every op raises.  The oracle
under oracle/ is test infrastructure and is never imported from here.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVID_B200_LIB") or os.path.join(HERE, "libavid_b200.so")      # the override is for A/B probes of a kernel build
HEADER = os.path.join(os.path.dirname(HERE), "include", "avid_b200.h")

AVID_MAX_KEYS = 8
MATH_FP32, MATH_BF16X3, MATH_BF16 = 0, 1, 2

c_void_p, c_int32, c_int64, c_uint64, c_float, c_size_t = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t


class NceKey(C.Structure):
    _fields_ = [("ctx", c_int32), ("bank", c_int32), ("pos_mode", c_int32), ("num_neg", c_int32), ("weight", c_float)]


class NceArgs(C.Structure):
    _fields_ = [
        ("emb_video", c_void_p), ("emb_audio", c_void_p), ("y", c_void_p), ("bank_video", c_void_p), ("bank_audio", c_void_p),
        ("num_rows", c_int64), ("row_begin", c_int64), ("row_end", c_int64),
        ("batch", c_int32), ("mean_batch", c_int32), ("num_neg", c_int32),
        ("neg_idx", c_void_p), ("seed", c_uint64), ("offset", c_uint64),
        ("positive_set", c_void_p), ("pos_k", c_int32), ("num_keys", c_int32),
        ("keys", NceKey * AVID_MAX_KEYS),
        ("avg_exp_score", c_void_p), ("temperature", c_float),
        ("loss_keys", c_void_p), ("loss_total", c_void_p), ("grad_video", c_void_p), ("grad_audio", c_void_p),
        ("scores", c_void_p), ("neg_idx_out", c_void_p),
        ("grad_hat_video", c_void_p), ("grad_hat_audio", c_void_p), ("loss_part", c_void_p),
        ("bad_index", c_void_p), ("group_batch", c_int32), ("in_group_stride", c_int64), ("out_group_stride", c_int64),
    ]


class BnBackwardFuse(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("z", "mean", "invstd", "gamma", "beta", "sums")]


AVID_MAX_PEERS = 16


class PeerPtrs(C.Structure):
    _fields_ = [("ptr", c_void_p * AVID_MAX_PEERS)]


class ConvShape(C.Structure):
    _fields_ = [(n, c_int32) for n in ("n", "ti", "hi", "wi", "ci", "to", "ho", "wo", "co", "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw")]


class VideoPrep(C.Structure):
    """avid_video_prep_t (include/avid_b200.h)."""
    _fields_ = [(n, c_int32) for n in ("frames", "height", "width", "crop_top", "crop_left", "crop_h", "crop_w", "out_h", "out_w", "flip", "num_ops")] + \
               [("op_kind", c_int32 * 4), ("op_factor", c_float * 4), ("hue_shift", c_int32), ("normalize", c_int32), ("mean", c_float * 3), ("std", c_float * 3)]


_P, _I, _L, _F, _Z, _U, _D = c_void_p, c_int32, c_int64, c_float, c_size_t, c_uint64, C.c_double
_SIGNATURES = {
    "avid_version": (C.c_int, []),
    "avid_last_error": (C.c_char_p, []),
    "avid_launch_count": (c_uint64, []),
    "avid_reset_launch_count": (None, []),
    "avid_nce_workspace_bytes": (c_size_t, [_I, _I, _I, _I]),
    "avid_nce_forward_backward": (C.c_int, [C.POINTER(NceArgs), _P, _Z, _P]),
    "avid_nce_finalize": (C.c_int, [C.POINTER(NceArgs), _P, _Z, _P]),
    "avid_nce_partition_mean": (C.c_int, [C.POINTER(NceArgs), _I, _P, _P, _Z, _P]),
    "avid_bank_update": (C.c_int, [_P, _P, _L, _L, _P, _P, _P, _I, _I, _L, _F, _F, _P]),
    "avid_rows_l2_normalize": (C.c_int, [_P, _L, _P]),
    "avid_bank_init": (C.c_int, [_P, _L, _L, _U, _I, _P]),
    "avid_sample_negatives": (C.c_int, [_P, _I, _I, _L, _P, _I, _U, _U, _P, _P]),
    "avid_cma_topk_workspace_bytes": (c_size_t, [_L]),
    "avid_cma_topk_begin": (C.c_int, [_L, _P, _Z, _P]),
    "avid_cma_topk_scan": (C.c_int, [_P, _P, _L, _P, _P, _L, _L, _I, _I, _P, _Z, _P]),
    "avid_cma_topk_finish": (C.c_int, [_L, _I, _P, _P, _Z, _P]),
    "avid_log_spectrogram_workspace_bytes": (c_size_t, [_I]),
    "avid_log_spectrogram": (C.c_int, [_P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _Z, _P]),
    "avid_video_prep_workspace_bytes": (c_size_t, [C.POINTER(VideoPrep)]),
    "avid_video_prep": (C.c_int, [_P, C.POINTER(VideoPrep), _P, _P, _Z, _P]),
    "avid_video_prep_batch_workspace_bytes": (c_size_t, [C.POINTER(VideoPrep), _I]),
    "avid_video_prep_batch": (C.c_int, [C.POINTER(c_void_p), C.POINTER(VideoPrep), _I, C.POINTER(c_void_p), _P, _Z, _P]),
    "avid_cma_to_half": (C.c_int, [_P, _P, _L, _P]),
    "avid_cma_topk_scan_tc": (C.c_int, [_P, _P, _L, _P, _P, _L, _L, _I, _P, _Z, _P]),
    "avid_cma_topk_rescore": (C.c_int, [_P, _P, _L, _P, _P, _L, _L, _I, _P, _Z, _P]),
    "avid_cma_topk_certify": (C.c_int, [_L, _I, _F, _P, _Z, _P, _P, _P]),
    "avid_conv_forward": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _P, _I, _P]),
    "avid_conv_dgrad": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _P, _I, _P]),
    "avid_conv_wgrad": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _I, _P]),
    "avid_split_bf16": (C.c_int, [_P, _P, _P, _L, _P]),
    "avid_conv_tc_uses_cta_pairs": (C.c_int, [C.POINTER(ConvShape), C.c_int32]),
    "avid_conv_forward_tc": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _P, _P, _P, _P, _P]),
    "avid_conv_dgrad_tc": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _P, _P, _P, C.POINTER(BnBackwardFuse), _P]),
    "avid_conv_tc_plan": (C.c_int, [C.POINTER(ConvShape), c_int32, C.POINTER(c_int64)]),
    "avid_conv_dgrad_tc_sub": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _P, _P, C.POINTER(C.c_int32), _P, C.POINTER(BnBackwardFuse), _P]),
    "avid_conv_wgrad_tc": (C.c_int, [C.POINTER(ConvShape), _P, _P, _P, _P, _P, _P]),
    "avid_stem_pack": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "avid_stem_filter_pack": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "avid_stem_forward_tc": (C.c_int, [C.POINTER(ConvShape), _P, _P, _I, _P, _P, _P, _P, _P]),
    "avid_stem_wgrad_tc": (C.c_int, [C.POINTER(ConvShape), _P, _P, _I, _P, _P, _P, _P]),
    "avid_filter_to_tapmajor": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "avid_filter_from_tapmajor": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "avid_nchw_to_nhwc": (C.c_int, [_P, _P, _I, _I, _L, _I, _P]),
    "avid_nhwc_to_nchw": (C.c_int, [_P, _P, _I, _I, _L, _P]),
    "avid_bn_stats": (C.c_int, [_P, _L, _I, _P, _P]),
    "avid_bn_finalize": (C.c_int, [_P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "avid_bn_relu_forward": (C.c_int, [_P, _P, _P, _P, _L, _I, _P]),
    "avid_bn_relu_forward_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, _I, _P]),
    "avid_bn_relu_backward_apply_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _P]),
    "avid_bn_relu_backward_reduce": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, _I, _P, _P]),
    "avid_bn_relu_backward_apply": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P]),
    "avid_bn_relu_maxpool_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "avid_bn_relu_maxpool_backward_reduce": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "avid_bn_relu_maxpool_backward_apply": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "avid_maxpool_1x3x3_forward": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "avid_maxpool_1x3x3_backward": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "avid_global_maxpool_forward": (C.c_int, [_P, _P, _P, _I, _L, _I, _P]),
    "avid_global_maxpool_backward": (C.c_int, [_P, _P, _P, _I, _L, _I, _P]),
    "avid_linear_forward": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "avid_linear_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "avid_filter_to_planes": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "avid_filter_to_planes_multi": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "avid_filter_from_tapmajor_multi": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "avid_adam_step_multi": (C.c_int, [_P, _P, _P, _P, _P, _I, _L, _F, _F, _F, _F, _F, _F, _P]),
    "avid_zero_bytes": (C.c_int, [_P, _Z, _P]),
    "avid_adam_shard_step": (C.c_int, [_P, C.POINTER(PeerPtrs), _I, _P, _P, _L, _L, _L, _D, _D, _D, _D, _D, _D, _P]),
    "avid_pull_shards": (C.c_int, [_P, C.POINTER(PeerPtrs), _I, _I, _L, _P]),
    "avid_add_inplace": (C.c_int, [_P, _P, _L, _P]),
    "avid_adam_step": (C.c_int, [_P, _P, _P, _P, _L, _L, _F, _F, _F, _F, _F, _F, _P]),
}


def declared_symbols(header=HEADER):
    """Names of every function include/avid_b200.h declares (used by the symbol-export test)."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avid_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    """The loaded library; raises RuntimeError when it has not been built (python -m avid_cma_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m avid_cma_b200.build`; "
                               "avid_cma_b200 has no CPU or PyTorch fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class AvidError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise AvidError(f"libavid_b200 error {rc}: {lib().avid_last_error().decode()}")
