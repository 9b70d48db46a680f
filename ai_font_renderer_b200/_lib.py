"""ctypes binding of libafr_sm100.so (include/afr_sm100.h).

There is no CPU or PyTorch fallback: if the library cannot be loaded the import of the product
path fails loudly (AfrLibraryError).
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build


class AfrLibraryError(RuntimeError):
    pass


class AfrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libafr_sm100 error {code}: {message}")
        self.code = code
        self.message = message


AFR_OK = 0
AFR_ERR_INVALID = -1
AFR_ERR_CUDA = -2
AFR_ERR_UNSUPPORTED = -3
AFR_ERR_STATE = -4
AFR_ERR_TOKEN_RANGE = -5

OUT_SHEET_F32, OUT_SHEET_U8, OUT_LOGITS_F32 = 0, 1, 2
TARGET_U8, TARGET_F32 = 0, 1


class AfrConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "device", "vocab", "max_length", "embed_dim", "num_heads", "hidden",
        "sheet_h", "sheet_w", "max_batch", "training")]


TENSOR_FIELDS = (
    "positional_encoding", "embedding_weight", "in_proj_weight", "in_proj_bias",
    "out_proj_weight", "out_proj_bias", "layer_norm_weight", "layer_norm_bias",
    "fc1_weight", "fc1_bias", "fc_output_weight", "fc_output_bias")

# state_dict keys of the reference module in the same order
STATE_DICT_KEYS = (
    "positional_encoding", "embedding.weight", "attention.in_proj_weight",
    "attention.in_proj_bias", "attention.out_proj.weight", "attention.out_proj.bias",
    "layer_norm.weight", "layer_norm.bias", "fc1.weight", "fc1.bias",
    "fc_output.weight", "fc_output.bias")


class AfrTensors(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in TENSOR_FIELDS]


class AfrDropout(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("seed", C.c_uint64), ("step", C.c_uint64),
        ("sample_offset", C.c_int64),
        ("mask_embed", C.c_void_p), ("mask_attn", C.c_void_p), ("mask_fc1", C.c_void_p),
        ("p_embed", C.c_double), ("p_attn", C.c_double), ("p_fc1", C.c_double)]


# every symbol include/afr_sm100.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "afr_abi_version": (C.c_int, []),
    "afr_create": (C.c_int, [C.POINTER(AfrConfig), C.POINTER(_P)]),
    "afr_destroy": (C.c_int, [_P]),
    "afr_last_error": (C.c_char_p, [_P]),
    "afr_bind_params": (C.c_int, [_P, C.POINTER(AfrTensors)]),
    "afr_bind_grads": (C.c_int, [_P, C.POINTER(AfrTensors)]),
    "afr_bind_adam_state": (C.c_int, [_P, C.POINTER(AfrTensors), C.POINTER(AfrTensors)]),
    "afr_sync_shadow": (C.c_int, [_P, _P]),
    "afr_bind_shadow": (C.c_int, [_P, _P, _P]),
    "afr_bind_font_embedding": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "afr_set_font_ids": (C.c_int, [_P, _P]),
    "afr_set_sm_limit": (C.c_int, [_P, C.c_int]),
    "afr_set_smem_reserve": (C.c_int, [_P, C.c_int]),
    "afr_shadow_index": (C.c_int, [_P]),
    "afr_shadow_commit": (C.c_int, [_P]),
    "afr_forward_eval": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, C.c_int, _P]),
    "afr_train_frontend": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(AfrDropout), _P]),
    "afr_train_loss": (C.c_int, [_P, _P, C.c_int, C.c_double, _P, _P]),
    "afr_train_forward_loss": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, C.c_int,
                                         C.POINTER(AfrDropout), C.c_double, _P, _P]),
    "afr_train_wgrad": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "afr_train_dgrad": (C.c_int, [_P, _P]),
    "afr_train_dgrad_gemm": (C.c_int, [_P, _P]),
    "afr_train_frontend_backward": (C.c_int, [_P, _P]),
    "afr_set_coresident": (C.c_int, [_P, C.c_int]),
    "afr_train_step": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, C.c_int,
                                 C.POINTER(AfrDropout), C.c_double, _P, _P]),
    "afr_forward_train": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(AfrDropout),
                                    _P, _P]),
    "afr_backward": (C.c_int, [_P, _P, _P]),
    "afr_adamw_step": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_int64, _P]),
    "afr_adamw_rows": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_int64, C.c_int, C.c_int, _P]),
    "afr_adamw_rows_gather": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_int64, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P),
                                        C.c_int, C.c_int, _P]),
    "afr_adamw_rows_gather_nvls": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_int64, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "afr_adamw_rows_bg": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_int64, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P]),
    "afr_train_wgrad_to": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "afr_train_wgrad_to_bf16": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "afr_adamw_rows_gather_bf16": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_int64, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P),
                                             C.c_int, C.c_int, _P]),
    "afr_adamw_rows_gather_nvls_bf16": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                                  C.c_int64, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "afr_adamw_small": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_int64, _P]),
    "afr_train_wgrad_adamw": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_int64, C.c_int, C.c_int, _P]),
    "afr_train_bgrad": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "afr_debug_div_sqrt": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "afr_check_tokens": (C.c_int, [_P, _P]),
    "afr_workspace_ptr": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "afr_workspace_copy": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P]),
    "afr_gemm_tiles": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "afr_launch_count": (C.c_int64, [_P]),
    "afr_debug_phase_cycles": (C.c_int, [_P, C.c_int]),
    "afr_debug_frontend_forward": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int,
                                             C.POINTER(AfrDropout), _P, _P]),
    "afr_debug_frontend_backward": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int,
                                              C.POINTER(AfrDropout), _P, _P]),
    "afr_gemm_bf16": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int, _P, C.c_int64, C.c_int, _P,
                                C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                _P]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building first if the sources are newer and nvcc exists). Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not _build.is_current():
        try:
            _build.build(verbose=False)
        except Exception as exc:
            if not os.path.exists(path):
                raise AfrLibraryError(
                    f"{path} is missing and could not be built ({exc}). The B200 path has no "
                    "fallback: run `python -m ai_font_renderer_b200.build`.") from exc
            # a library that does not match the sources is only used when the caller says so
            # (AFR_ALLOW_STALE_LIB=1, e.g. a box without nvcc that received a prebuilt .so)
            if os.environ.get("AFR_ALLOW_STALE_LIB", "0") != "1":
                raise AfrLibraryError(
                    f"{path} does not match the sources and could not be rebuilt ({exc}); set "
                    "AFR_ALLOW_STALE_LIB=1 to load it anyway") from exc
    try:
        lib = C.CDLL(path)
    except OSError as exc:
        raise AfrLibraryError(f"cannot load {path}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise AfrLibraryError(f"{path} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    if lib.afr_abi_version() != 1:
        raise AfrLibraryError("libafr_sm100.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, ctx=None) -> None:
    if rc == AFR_OK:
        return
    msg = load().afr_last_error(ctx)
    text = msg.decode("utf-8", "replace") if msg else ""
    if rc == AFR_ERR_TOKEN_RANGE:
        raise IndexError("index out of range in self (token id outside the embedding table)")
    raise AfrError(rc, text)
