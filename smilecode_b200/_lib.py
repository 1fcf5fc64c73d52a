"""ctypes binding of libsmilecode_b200.so (the C ABI declared in include/smilecode_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, the
caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsmilecode_b200.so")

P = c_void_p
# name -> argtypes, mirroring include/smilecode_b200.h one to one (tests/test_abi.py checks this)
SIGNATURES = {
    "smile_modet_attn_fwd": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_warp3d_fwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_upsample2x_fwd": [P, P, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_flow_compose_fwd": [P, P, P, c_int, c_int, c_int, c_int, c_float, P],
    "smile_modet_fused_fwd": [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, P],
    "smile_proj_ln_fwd": [P, P, P, P, P, P, c_int, c_int, c_int, c_longlong, c_float, P],
    "smile_warp_proj_ln_fwd": [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_conv3d_fwd": [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_conv3d_bf16_fwd": [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_instnorm_lrelu_pool_fwd": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_cwm_fuse_fwd": [P, P, P, c_int, c_int, c_longlong, P],
    "smile_modet_qkrpb_fwd": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "smile_modet_qkrpb_bwd": [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "smile_ncc_vxm_fwd": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_grad3d_l2_fwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_warp3d_bwd": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_upsample2x_bwd": [P, P, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_modet_attn_bwd": [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_proj_ln_bwd": [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_longlong, c_float, P],
    "smile_cwm_fuse_bwd": [P, P, P, P, P, c_int, c_int, c_longlong, P],
    "smile_conv3d_flip_weights": [P, P, c_int, c_int, P],
    "smile_conv3d_wgrad": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "smile_conv3d_wgrad_bf16": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "smile_in_lrelu_bwd": [P, P, P, P, P, c_int, c_int, c_longlong, c_float, c_int, P],
    "smile_avgpool2_bwd_add": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_ncc_vxm_bwd": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_grad3d_l2_bwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_adam_amsgrad_step": [P, P, P, P, P, c_longlong, c_float, c_float, c_float, c_float, c_int, P],
    "smile_conv3d_tc_prep": [P, P, c_int, c_int, P],
    "smile_conv3d_prepped_fwd": [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smile_warp3d_nearest_fwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smile_dice_counts_fwd": [P, P, P, c_int, P, c_longlong, P],
    "smile_jacdet_fwd": [P, P, P, c_int, c_int, c_int, P],
}

_lib = None


class SmileError(RuntimeError):
    """Raised when a C-ABI entry point returns a non-zero status."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmileError(
                f"{LIB_PATH} is missing: build it with `python -m smilecode_b200.build` "
                "(nvcc, sm_100a).  smilecode_b200 has no CPU / PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        handle.smile_version.restype = c_int
        handle.smile_version.argtypes = []
        handle.smile_last_error.restype = c_char_p
        handle.smile_last_error.argtypes = []
        handle.smile_ncc_vxm_work_bytes.restype = c_longlong
        handle.smile_ncc_vxm_work_bytes.argtypes = [c_int, c_int, c_int, c_int]
        handle.smile_conv3d_tc_prep_floats.restype = c_longlong
        handle.smile_conv3d_tc_prep_floats.argtypes = [c_int, c_int]
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = c_int
            fn.argtypes = argtypes
        _lib = handle
    return _lib


LAUNCHES = 0          # number of C-ABI kernel launches issued by this process (bench.py reports it)
_profile = None       # optional {label: [cuda event pairs]} filled when profiling is on


def profile_start() -> None:
    global _profile
    _profile = {}


def profile_stop():
    """Returns {label: (calls, total_ms)} measured with CUDA events around every C-ABI call."""
    global _profile
    import torch
    torch.cuda.synchronize()
    out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (_profile or {}).items()}
    _profile = None
    return out


def call(name: str, *args, label: str = "") -> None:
    global LAUNCHES
    handle = lib()
    LAUNCHES += 1
    if _profile is not None:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(handle, name)(*args)
        b.record()
        _profile.setdefault(f"{name}{label}", []).append((a, b))
    else:
        rc = getattr(handle, name)(*args)
    if rc != 0:
        msg = handle.smile_last_error()
        raise SmileError(f"{name} failed (code {rc}): {msg.decode() if msg else '?'}")
