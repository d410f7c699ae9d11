"""ctypes binding of include/llmvox_b200.h.  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libllmvox_b200.so")

LVX_OK = 0
PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_EXACT = 2
PATH_AUTO, PATH_CLUSTER, PATH_PER_OP, PATH_CLUSTER16, PATH_CLUSTER8, PATH_PER_OP_TAIL = 0, 1, 2, 3, 4, 5


class LvxConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_layer", "n_head", "n_embd", "block_size", "vocab_size", "bias",
        "text_vocab", "text_dim", "code_dim", "n_codes",
        "voc_dim", "voc_inter", "voc_layers", "voc_ada_rows", "n_fft", "hop",
        "max_sessions", "max_context", "kv_page_tokens", "kv_pages", "max_batch", "max_vocode_frames",
        "precision", "pad_token_id", "eoa_token_id", "decode_lanes")]


class LvxSampling(C.Structure):
    _fields_ = [("greedy", C.c_int32), ("top_k", C.c_int32), ("temperature", C.c_float),
                ("seed", C.c_uint64), ("d_uniform", C.c_void_p)]


# every symbol include/llmvox_b200.h declares: (name, restype, argtypes)
_I32P = C.POINTER(C.c_int32)
_VP = C.c_void_p
SYMBOLS = [
    ("lvx_last_error", C.c_char_p, []),
    ("lvx_version", C.c_int, []),
    ("lvx_config_default", C.c_int, [C.POINTER(LvxConfig)]),
    ("lvx_engine_create", C.c_int, [C.POINTER(LvxConfig), C.c_int, C.POINTER(_VP)]),
    ("lvx_engine_destroy", C.c_int, [_VP]),
    ("lvx_load_tensor", C.c_int, [_VP, C.c_char_p, _VP, C.POINTER(C.c_int64), C.c_int]),
    ("lvx_finalize_weights", C.c_int, [_VP]),
    ("lvx_session_open", C.c_int, [_VP, _I32P, C.c_int, _VP]),
    ("lvx_session_close", C.c_int, [_VP, _I32P, C.c_int, _VP]),
    ("lvx_feed_text", C.c_int, [_VP, _I32P, _I32P, _I32P, C.c_int, _VP]),
    ("lvx_feed_utf8", C.c_int, [_VP, _I32P, _I32P, C.c_char_p, C.c_int, C.c_int, _I32P, _VP]),
    ("lvx_decode_steps", C.c_int, [_VP, _I32P, C.c_int, C.c_int, C.POINTER(LvxSampling), _VP]),
    ("lvx_decode_steps_lane", C.c_int, [_VP, C.c_int, _I32P, C.c_int, C.c_int, C.POINTER(LvxSampling), _VP]),
    ("lvx_decode_steps_ex", C.c_int, [_VP, C.c_int, _I32P, C.c_int, C.c_int, C.POINTER(LvxSampling), C.c_int, _VP]),
    ("lvx_session_progress", C.c_int, [_VP, _I32P, C.c_int, _VP, _VP]),
    ("lvx_set_cluster_decode", C.c_int, [_VP, C.c_int]),
    ("lvx_cluster_capacity", C.c_int, [_VP, _I32P, _I32P]),
    ("lvx_decode_step_logits", C.c_int, [_VP, _I32P, C.c_int, C.POINTER(LvxSampling), _VP, _VP, _VP, _VP]),
    ("lvx_peek_buffer", C.c_int, [_VP, C.c_int, C.c_int, _VP, C.c_int64]),
    ("lvx_peek_trace", C.c_int, [_VP, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    ("lvx_peek_logits", C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP]),
    ("lvx_decode_step_embeds", C.c_int, [_VP, _I32P, C.c_int, _VP, _I32P, _VP, _VP]),
    ("lvx_gather_codes", C.c_int, [_VP, _I32P, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    ("lvx_gather_code_ranges", C.c_int, [_VP, _I32P, _I32P, _I32P, C.c_int, _VP, _VP]),
    ("lvx_session_length", C.c_int, [_VP, C.c_int, _I32P]),
    ("lvx_session_text", C.c_int, [_VP, C.c_int, _I32P, C.c_int, _I32P, _VP]),
    ("lvx_codes_to_features", C.c_int, [_VP, _VP, C.c_int, _VP, _VP]),
    ("lvx_text_embed", C.c_int, [_VP, _VP, C.c_int, _VP, _VP]),
    ("lvx_vocode", C.c_int, [_VP, _VP, _I32P, C.c_int, C.c_int, _VP, _VP]),
    ("lvx_vocode_features", C.c_int, [_VP, _VP, _I32P, C.c_int, C.c_int, _VP, _VP]),
    ("lvx_vocode_stage", C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    ("lvx_test_gemm", C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    ("lvx_profile_enable", C.c_int, [_VP, C.c_int]),
    ("lvx_profile_report", C.c_int, [_VP, C.c_char_p, C.c_int64]),
    ("lvx_kernel_launches", C.c_int64, [_VP]),
    ("lvx_device_bytes", C.c_int64, [_VP]),
]

_lib = None


class LvxError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"llmvox_b200 status {status}: {msg}")
        self.status = status


def load() -> C.CDLL:
    """Loads libllmvox_b200.so (built in-tree by llmvox_b200.build).  Raises if it is missing: the product
    path has no CPU or PyTorch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m llmvox_b200.build` (nvcc, sm_100a). "
            "llmvox_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != LVX_OK:
        raise LvxError(status, load().lvx_last_error().decode("utf-8", "replace"))


def i32_array(values):
    """int32 pointer argument from a sequence / array.  Goes through numpy (a C loop): the ctypes array constructor cost
    1.5 ms for the 12,800 text ids of a 64-stream batch, 40 % of its first-chunk latency.  The returned pointer keeps the
    array alive (numpy: ctypes.data_as)."""
    import numpy as np
    arr = np.ascontiguousarray(values, dtype=np.int32)
    if arr.size == 0:
        arr = np.zeros((1,), dtype=np.int32)
    return arr.ctypes.data_as(_I32P)
