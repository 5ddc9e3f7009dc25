"""ctypes binding of liblafs_b200.so (the C ABI declared in include/lafs_b200.h).

There is NO CPU fallback: if the shared library is missing or a tensor is not on a CUDA
device the wrappers raise.  Build the library with `python -m lafs_cvpr2024_b200.build`
(or __graft_entry__.build()).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblafs_b200.so")

F32, BF16, F16, U8 = 0, 1, 2, 3
LAYOUT_MOSAIC, LAYOUT_TOKENS = 0, 1
COORD_DIV, COORD_RECIP = 0, 1
EMA_CHUNK = 16384
TOK_LD = 208          # row pitch of the saved-token tensor (lafs_gather_embed_fwd_save)

_p, _i, _f, _z, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64

# name -> (restype, argtypes); mirrors include/lafs_b200.h one to one
SIGNATURES = {
    "lafs_version": (_i, []),
    "lafs_last_error_string": (C.c_char_p, []),
    "lafs_device_ok": (_i, []),
    "lafs_ema_multi": (_i, [_p, _i, _f, _f, _i, _p]),
    "lafs_dino_workspace_bytes": (_z, [_i, _i, _i]),
    "lafs_dino_fwd": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _z, _p, _f, _f, _p]),
    "lafs_dino_fused_workspace_bytes": (_z, [_i, _i, _i]),
    "lafs_dino_fwd_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _p, _z, _p, _f, _f, _p]),
    "lafs_dino_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _p, _p]),
    "lafs_center_ema": (_i, [_p, _p, _f, _f, _f, _i, _p, _p]),
    "lafs_colsum": (_i, [_p, _i, _i, _i, _p, _p, _z, _p]),
    "lafs_landmark_post": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "lafs_landmark_post_bwd": (_i, [_p, _p, _p, _i, _i, _f, _p]),
    "lafs_gather_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "lafs_gather_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "lafs_embed_weight_prep": (_i, [_p, _p, _i, _p, _p, _p]),
    "lafs_gather_embed_fwd": (_i, [_p, _i, _f, _f, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "lafs_gather_embed_fwd_save": (_i, [_p, _i, _f, _f, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "lafs_gather_embed_seq_fwd": (_i, [_p, _i, _f, _f, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _f,
                                       C.c_uint, _p]),
    "lafs_embed_bwd_weight_perm": (_i, [_p, _p, _i, _i, _p, _p, _i, _p, _z, _p]),
    "lafs_normalize_rows": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "lafs_head_workspace_bytes": (_z, [_i, _i, _i]),
    "lafs_head_fwd": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _f, _f, _i, _p, _p, _z, _p]),
    "lafs_head_merge": (_i, [_p, _i, _i, _p, _p]),
    "lafs_head_loss": (_i, [_p, _p, _p, _f, _i, _p, _p, _p]),
    "lafs_head_logits": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _f, _f, _i, _p, C.c_longlong, _p]),
    "lafs_head_grad_logits": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _f, _f, _i, _p, _p, _f, _p, C.c_longlong, _p]),
    "lafs_head_bwd_workspace_bytes": (_z, [_i, _i, _i]),
    "lafs_head_bwd_embed": (_i, [_p, C.c_longlong, _p, _i, _i, _i, _p, _p, _z, _p]),
    "lafs_head_bwd_weight": (_i, [_p, C.c_longlong, _p, _p, _p, _i, _i, _i, _p, _p]),
    "lafs_normalize_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "lafs_embed_bwd_workspace_bytes": (_z, [_i, _i]),
    "lafs_embed_bwd_weight": (_i, [_p, _p, _i, _i, _p, _p, _z, _p]),
    "lafs_embed_bwd_tokens": (_i, [_p, _p, _i, _i, _p, _p]),
    "lafs_head_grad_logits_t": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _f, _f, _i, _p, _p, _f, _p, C.c_longlong, _p,
                                     C.c_longlong, _p]),
    "lafs_head_bwd_weight_t": (_i, [_p, C.c_longlong, _p, _p, _p, _p, _i, C.c_longlong, _i, _i, _i, _p, _p]),
    "lafs_gemm_tn": (_i, [_p, C.c_longlong, _p, C.c_longlong, _i, _i, _i, _p, C.c_longlong, _p]),
    "lafs_dh_extra_cols": (_i, []),
    "lafs_dh_prep_rows": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "lafs_dh_xsum": (_i, [_p, _i, _i, _i, _p, _p]),
    "lafs_dh_prep_weight": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "lafs_dh_lse2": (_i, [_p, _i, _p, _p]),
    "lafs_dh_loss": (_i, [_p, _p, _p, _i, _i, _i, _f, _p, _p, _p]),
    "lafs_dh_bwd_rows": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _p, _p]),
    "lafs_dh_wn_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _f, _p, _p, _p]),
    "lafs_optim_workspace_bytes": (_z, [_i, _i]),
    "lafs_adamw_ema_multi": (_i, [_p, _i, _p, _p, _i, _p, _p, _p, _p, _z, _p]),
    "lafs_xchg_bytes": (_z, [_i, _i, _i, _p]),
    "lafs_xchg_stats": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "lafs_xchg_allreduce": (_i, [_p, _i, _i, _i, _i, _p]),
}

_lib = None


def lib():
    """Loads (once) and returns the ctypes handle.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m lafs_cvpr2024_b200.build`.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def call(name, *args):
    """Calls an int-returning entry point and raises RuntimeError on a non-zero status."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        msg = lib().lafs_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed ({rc}): {msg}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"unsupported dtype {t.dtype} (expected float32, bfloat16 or float16)")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lafs_cvpr2024_b200 runs on CUDA (sm_100a) tensors only; got a "
                               f"{t.device} tensor and there is no CPU fallback")


def ptr(t):
    return None if t is None else t.data_ptr()
