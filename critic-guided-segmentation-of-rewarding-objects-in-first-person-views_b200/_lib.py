"""ctypes binding of libcgs_b200.so (include/cgs_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgs_b200.so")

SRC_PLAIN, SRC_CATUP, SRC_POOLBWD, SRC_SIGGRAD, SRC_LEAKYGRAD, SRC_U8ROLL = range(6)
EPI_LINEAR, EPI_LEAKY, EPI_RELU_POOL, EPI_SIGMOID, EPI_MUL, EPI_SPLIT_UP = range(6)

_f32p = C.c_void_p
_u8p = C.c_void_p


class Src(C.Structure):
    _fields_ = [("mode", C.c_int32), ("C", C.c_int32), ("C0", C.c_int32), ("shift", C.c_int32),
                ("a", _f32p), ("b", _f32p), ("idx", _u8p)]


class Conv3x3Args(C.Structure):
    _fields_ = [("src", Src), ("w", _f32p), ("bias", _f32p), ("transposed", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cout", C.c_int32), ("epi", C.c_int32),
                ("out", _f32p), ("out2", _f32p), ("idx_out", _u8p), ("mul", _f32p),
                ("C0", C.c_int32), ("shift2", C.c_int32), ("thresh", C.c_float), ("precision", C.c_int32)]


class Wgrad3x3Args(C.Structure):
    _fields_ = [("x", Src), ("dy", Src), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("dw", _f32p), ("db", _f32p), ("precision", C.c_int32)]


class CriticWeights(C.Structure):
    """cgs_critic_weights: the 14 NewCritic tensors in state_dict order."""
    _fields_ = [(n, _f32p) for n in ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4", "wl1", "bl1", "wl2", "bl2")]


class MaskerWeights(C.Structure):
    """cgs_masker_weights: the 14 UnetDecoder tensors in state_dict order."""
    _fields_ = [(n, _f32p) for n in ("wd0", "bd0", "wd1", "bd1", "wd2", "bd2", "wd3", "bd3", "wd4", "bd4", "wm0", "bm0", "wm2", "bm2")]


class WidePackJob(C.Structure):
    """cgs_wide_packjob"""
    _fields_ = [("w", _f32p), ("out", C.c_void_p), ("Cin", C.c_int32), ("Cout", C.c_int32), ("transposed", C.c_int32)]


class WideColJob(C.Structure):
    """cgs_wide_coljob"""
    _fields_ = [("X", _f32p), ("out", _f32p), ("n", C.c_int32), ("scale", C.c_float), ("accumulate", C.c_int32)]


class AdamArgs(C.Structure):
    """cgs_adam_args"""
    _fields_ = [("p", _f32p), ("g", _f32p), ("m", _f32p), ("v", _f32p), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("step_state", C.c_void_p), ("barrier", C.c_void_p),
                ("world", C.c_int32), ("rank", C.c_int32), ("npad", C.c_int64), ("peer_recv", C.c_void_p)]


EXPORTS = {
    "cgs_critic_fused_supported": [C.c_int32] * 5,
    "cgs_critic_train_fused": [_u8p, _f32p, C.c_int32, C.c_int32, C.c_void_p, _f32p, _f32p, _f32p,
                               C.c_float, C.c_uint64, C.c_void_p,
                               C.POINTER(CriticWeights), C.POINTER(CriticWeights), _f32p, C.POINTER(AdamArgs), C.c_float,
                               C.c_int32, _f32p, _f32p, C.c_void_p],
    "cgs_critic_train_bf16": [_u8p, _f32p, C.c_int32, C.c_int32, C.c_void_p, _f32p, _f32p, _f32p, C.c_float, C.c_uint64, C.c_void_p,
                              C.POINTER(CriticWeights), _f32p, C.POINTER(AdamArgs), C.c_float, C.c_int32, _f32p, _f32p, C.c_void_p],
    "cgs_hg_set_trace_critic": [C.c_void_p],
    "cgs_critic_fused_grid": [C.c_int32],
    "cgs_critic_fused_partial_stride": [],
    "cgs_reduce_partials": [_f32p, C.c_int64, _f32p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p],
    "cgs_adam_step_partials": [_f32p] * 4 + [C.c_int64] + [C.c_double] * 4 + [C.c_void_p, C.c_float, _f32p, C.c_int32,
                                                                             C.c_int64, C.c_int64, C.c_int64, C.c_void_p],
    "cgs_critic_loss_xgrad": [_f32p, _f32p, C.c_int32, _f32p, _f32p, _f32p, C.c_float, C.c_uint64, C.c_void_p,
                              C.POINTER(CriticWeights), C.c_float, C.c_int32, _f32p, _f32p, _f32p, C.c_void_p],
    "cgs_critic_forward_frames": [_u8p, C.c_int32, C.c_int32, C.c_void_p, _f32p, _f32p, _f32p, C.c_float, C.c_uint64, C.c_void_p,
                                  C.POINTER(CriticWeights), _f32p, C.c_void_p],
    "cgs_hg_score_bf16": [_u8p, _u8p, C.c_int32, C.c_int32, C.c_void_p, _f32p, _f32p, _f32p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p,
                          C.POINTER(CriticWeights), C.c_void_p, C.c_float, _f32p, C.c_float, C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p,
                          C.c_void_p],
    "cgs_hg_score": [_u8p, _u8p, C.c_int32, C.c_int32, C.c_void_p, _f32p, _f32p, _f32p] + [_f32p] * 6 +
                    [C.c_float, C.c_uint64, C.c_void_p, C.POINTER(CriticWeights), C.c_float, _f32p, C.c_float, C.c_float,
                     _f32p, _f32p, _f32p, _f32p, C.c_void_p],
    "cgs_hg_pack_words": [],
    "cgs_hg_tape_bytes": [],
    "cgs_hg_grid": [C.c_int32],
    "cgs_hg_partial_stride": [],
    "cgs_hg_debug_floats": [],
    "cgs_hg_status": [],
    "cgs_hg_set_trace": [C.c_void_p, C.c_void_p, C.c_void_p],
    "cgs_hg_pack": [C.POINTER(CriticWeights), C.POINTER(MaskerWeights), C.c_void_p, C.c_void_p],
    "cgs_hg_forward": [_u8p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(CriticWeights), C.POINTER(MaskerWeights), C.c_void_p,
                       C.c_int32, _f32p, _f32p, _f32p, C.c_float, C.c_uint64, C.c_void_p, C.c_float, _f32p, _f32p, _u8p,
                       C.c_void_p, C.c_void_p],
    "cgs_hg_backward": [_u8p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(MaskerWeights), C.c_void_p, C.c_void_p, _f32p, _f32p,
                        _f32p, _f32p, C.c_void_p],
    "cgs_infer_pack_floats": [],
    "cgs_infer_pack_decoder": [_f32p] * 5 + [C.c_void_p],
    "cgs_infer_fused": [_u8p, C.c_int32, C.POINTER(CriticWeights)] + [_f32p] * 9 + [C.c_void_p],
    "cgs_masker_fused": [_u8p, _f32p, C.c_int32, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, _u8p, C.c_void_p],
    "cgs_p2p_stage": [_f32p, C.c_int64, C.c_int64, _f32p, _f32p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                      C.c_void_p],
    "cgs_p2p_allreduce_adam": [_f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] +
                              [C.c_double] * 4 + [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p],
    "cgs_conv3x3": [C.POINTER(Conv3x3Args), C.c_void_p],
    "cgs_wgrad3x3": [C.POINTER(Wgrad3x3Args), C.c_void_p],
    "cgs_conv_rgb_fwd": [_u8p] + [C.c_int32] * 4 + [C.c_void_p, _f32p, _f32p, C.c_int32, _f32p, _u8p, C.c_void_p],
    "cgs_head_fwd": [_f32p] * 9 + [C.c_int32] * 3 + [_f32p] * 3 + [C.c_void_p],
    "cgs_head_bwd": [_f32p] * 11 + [C.c_int32] * 3 + [_f32p] * 7 + [C.c_void_p],
    "cgs_tail_supported": [C.c_int32] * 4,
    "cgs_tail_fwd": [_f32p] * 12 + [C.c_int32] * 4 + [_f32p, _u8p, _f32p, _f32p, _f32p, C.c_void_p],
    "cgs_tail_bwd": [_f32p] * 9 + [_u8p] + [_f32p] * 6 + [C.c_int32] * 4 + [_f32p] * 9 + [C.c_void_p],
    "cgs_dense_fwd": [_f32p] * 3 + [C.c_int32] * 3 + [_f32p, C.c_void_p],
    "cgs_dense_bwd": [_f32p] * 3 + [C.c_int32] * 3 + [_f32p] * 3 + [C.c_void_p],
    "cgs_occlude_fwd": [_f32p] * 3 + [C.c_int64, C.c_int32, _f32p, C.c_void_p],
    "cgs_occlude_bwd": [_f32p] * 4 + [C.c_int64, C.c_int32] + [_f32p] * 3 + [C.c_void_p],
    "cgs_pred_loss": [_f32p, _f32p, C.c_int32, C.c_int32, C.c_float, _f32p, _f32p, C.c_void_p],
    "cgs_mask_reg": [_f32p, _f32p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_float, _f32p, _f32p, C.c_void_p],
    "cgs_frames_to_float": [_u8p] + [C.c_int32] * 5 + [C.c_void_p, _f32p, C.c_void_p],
    "cgs_adam_step": [_f32p] * 4 + [C.c_int64] + [C.c_double] * 4 + [C.c_void_p, C.c_float, C.c_int32, C.c_void_p],
    "cgs_threshold": [_f32p, C.c_int64, C.c_float, C.c_int32, _u8p, C.c_void_p],
    "cgs_dropout_masks": [_f32p, C.c_int64, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p],
    "cgs_iou_counts": [_f32p, _u8p, C.c_int64, C.c_float, C.c_int32, C.c_void_p, C.c_void_p],
    "cgs_gather_frames": [_u8p, C.c_int64, C.c_void_p, C.c_int32, _u8p, C.c_void_p],
    "cgs_mask_images": [_f32p, _u8p, C.c_int32, _u8p, _u8p, C.c_int32, _u8p, _u8p, C.c_void_p],
    "cgs_saliency_normalize": [_f32p, _f32p, C.c_int32, C.c_int32, C.c_float, _f32p, _f32p, _u8p, _f32p, C.c_void_p],
    "cgs_tc_status": [],
    "cgs_wide_conv3x3": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p, C.c_void_p, _f32p, C.c_int32, C.c_int32, C.c_int32,
                         C.c_void_p, _f32p, _u8p, _u8p, _f32p, C.c_void_p],
    "cgs_wide_pack": [C.c_void_p, C.c_int32, C.c_void_p],
    "cgs_wide_packed_bytes": [C.c_int32, C.c_int32],
    "cgs_wide_wgrad3x3": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p, _f32p, C.c_int64,
                          C.c_void_p],
    "cgs_wide_wgrad_workspace": [C.c_int32, C.c_int32, C.c_int32, C.c_int32],
    "cgs_wide_status": [],
    "cgs_wide_conv0_fwd": [_u8p, C.c_int32, C.c_int32, C.c_void_p, _f32p, _f32p, C.c_int32, C.c_void_p, _u8p, C.c_void_p],
    "cgs_wide_conv0_wgrad": [_u8p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, _u8p, C.c_int32, _f32p, _f32p, _f32p, C.c_int64, C.c_void_p],
    "cgs_wide_gemm": [_f32p, C.c_int32, C.c_int32, _f32p, C.c_int32, C.c_int32, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p,
                      C.c_int32, C.c_int32, C.c_int32, _f32p, C.c_void_p, C.c_void_p],
    "cgs_wide_head_mid": [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_float, C.c_int32, _f32p, _f32p, _f32p, _f32p, _f32p,
                          C.c_void_p],
    "cgs_wide_colsums": [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p],
    "cgs_wide_unpool3": [_f32p, _u8p, _f32p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    "cgs_wide_set_trace": [C.c_void_p],
    "cgs_tc_set_trace": [C.c_void_p],
    "cgs_critic_fused_set_trace": [C.c_void_p],
}

_lib = None


class CgsError(RuntimeError):
    pass


def lib():
    """Load the library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CgsError(f"{LIB_PATH} is missing: run `python -m cgs_b200.build` "
                           "(or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        for name, argtypes in EXPORTS.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        L.cgs_last_error.restype = C.c_char_p
        L.cgs_last_error.argtypes = []
        L.cgs_version.restype = C.c_int
        L.cgs_version.argtypes = []
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise CgsError(f"{what} failed (rc={rc}): {lib().cgs_last_error().decode()}")
