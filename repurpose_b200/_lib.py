"""ctypes binding of librepurpose_b200.so (the C ABI declared in include/repurpose_b200.h).

There is deliberately no fallback: if the shared library is missing the import fails, and if no
sm_100 device is present every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os

# RP_LIB_PATH: load another build of the library (A/B experiments); default = the in-tree build
LIB_PATH = Path(os.environ.get("RP_LIB_PATH", _PKG / "librepurpose_b200.so"))

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class RpModelCfg(C.Structure):
    _fields_ = [(n, c_i32) for n in ("vis_dim", "aud_dim", "text_dim", "d_model", "num_layers",
                                     "num_heads", "d_ff", "head_hidden", "max_len")]


class RpDropout(C.Structure):
    _fields_ = [("key_a", C.c_uint32), ("key_b", C.c_uint32), ("p", c_f32)]


class RpDecodeCfg(C.Structure):
    _fields_ = [("pre_nms_topk", c_i32), ("pre_nms_thresh", c_f32), ("duration_thresh", c_f32),
                ("duration_thresh_max", c_f32), ("nms_sigma", c_f32), ("min_score", c_f32)]


# name -> (restype, argtypes); must list every symbol include/repurpose_b200.h declares
SIGNATURES = {
    "rp_abi_version": (c_i32, []),
    "rp_last_error": (C.c_char_p, []),
    "rp_launch_count": (c_i64, []),
    "rp_create": (c_i32, [C.POINTER(RpModelCfg), C.POINTER(c_vp)]),
    "rp_destroy": (None, [c_vp]),
    "rp_load_weight": (c_i32, [c_vp, C.c_char_p, c_vp, c_i64, c_vp]),
    "rp_weights_complete": (c_i32, [c_vp]),
    "rp_workspace_bytes": (c_i64, [c_vp, c_i32, c_i32]),
    "rp_forward": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp,
                           c_i64, c_vp]),
    "rp_forward_ragged": (c_i32, [c_vp] * 8 + [c_i32, c_i32] + [c_vp] * 4 + [c_i64, c_vp]),
    "rp_forward_ragged_bf16": (c_i32, [c_vp] * 8 + [c_i32, c_i32] + [c_vp] * 4 + [c_i64, c_vp]),
    "rp_focal_loss_sum": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp, c_vp]),
    "rp_focal_loss_grad": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_vp, c_vp]),
    "rp_layernorm512_bwd_scratch_bytes": (c_i64, []),
    "rp_layernorm512_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "rp_adam_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp, c_vp]),
    "rp_train_scratch_bytes": (c_i64, []),
    "rp_cast_scaled": (c_i32, [c_vp, c_i64, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "rp_layernorm512_bwd_acc": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                                        c_vp]),
    "rp_gemm_bwd": (c_i32, [c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "rp_splitk_reduce": (c_i32, [c_vp, c_i32, c_i64, c_vp, c_vp]),
    "rp_colsum_bf16": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "rp_relu_bwd": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp]),
    "rp_relu_bwd_colsum": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "rp_head_out_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "rp_fmha_train": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "rp_fmha_bwd": (c_i32, [c_vp] * 10 + [c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "rp_dropout_mask_u8": (c_i32, [C.POINTER(RpDropout), c_i64, c_vp, c_vp]),
    "rp_attn_dropout_bits": (c_i32, [C.POINTER(RpDropout), c_i64, c_vp, c_vp]),
    "rp_gemm_bf16_dropout": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_i32,
                                     c_i32, c_i32, C.POINTER(RpDropout), c_vp]),
    "rp_layernorm512_dropout": (c_i32, [c_i32, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, C.POINTER(RpDropout), c_vp]),
    "rp_layernorm512_bwd_acc_dropout": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                                c_i64, C.POINTER(RpDropout), c_vp]),
    "rp_relu_bwd_scaled": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp]),
    "rp_relu_bwd_colsum_scaled": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp, c_vp, c_i64, c_vp]),
    "rp_head_out_bwd_scaled": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "rp_fmha_train_dropout": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64,
                                      c_f32, c_vp]),
    "rp_fmha_bwd_dropout": (c_i32, [c_vp] * 10 + [c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_f32, c_vp]),
    "rp_gemm_head_dot": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp]),
    "rp_gemm_resid_ln": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_f32, c_vp, c_i64,
                                 c_i32, c_i32, c_vp]),
    "rp_set_skip_padding": (c_i32, [c_vp, c_i32]),
    "rp_set_attn_bwd_deterministic": (c_i32, [c_i32]),
    "rp_profile_begin": (c_i32, [c_vp]),
    "rp_profile_end": (c_i32, [c_vp, c_vp, c_vp]),
    "rp_decode_nms": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, C.POINTER(RpDecodeCfg), c_i32,
                              c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "rp_soft_nms": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_f32, c_i32, c_vp, c_vp,
                            c_vp, c_vp]),
    "rp_atiou": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "rp_gemm_bf16": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_i32,
                             c_i32, c_i32, c_vp]),
    "rp_fmha": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                        c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_i64, c_i64, c_vp]),
    "rp_concat_cast": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp]),
    "rp_mask_lens": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "rp_cast_bf16": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "rp_layernorm512": (c_i32, [c_i32, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_vp, c_vp]),
    "rp_head_out": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
}


PROFILE_TAGS = ("cast", "gemm_in", "layernorm", "gemm_qkv", "fmha", "gemm_out", "gemm_ff1", "gemm_ff2",
                "gemm_fmap", "gemm_head", "head_out")


class RepurposeError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the library (built by `python -m repurpose_b200.build`) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RepurposeError(
            f"{LIB_PATH} not found: build it with `python -m repurpose_b200.build` "
            "(there is no CPU/PyTorch fallback for this path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rp_abi_version() != 1:
        raise RepurposeError("ABI version mismatch between _lib.py and librepurpose_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rp_last_error()
        raise RepurposeError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().rp_launch_count())


def ptr(t) -> int:
    """device/host pointer of a torch tensor (None -> NULL)"""
    return 0 if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
