"""ctypes binding of `libb200enc.so` (C ABI declared in include/b200enc.h).

The product path has NO fallback: if the shared library is missing or an entry point
fails, the call raises.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200enc.so")
CSRC = os.path.join(_HERE, "csrc")

EPI_STORE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_ADD, EPI_ATOMIC, EPI_BIAS_RES32 = range(8)
DT_F16, DT_F32 = 0, 1

_p, _i, _f, _sz, _ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong

# name -> argtypes; every function returns int (0 = ok) unless listed in _RESTYPE
_PROTOS = {
    "b200_gemm_f16": [_p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _p, _i, _p, _i, _p],
    "b200_attn_fwd": [_p, _i, _i, _p, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p],
    "b200_attn_probs": [_p, _i, _i, _p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p],
    "b200_mask_to_bias": [_p, _i, _p, _p, _i, _i, _p],
    "b200_layernorm_fwd": [_p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _f, _p],
    "b200_layernorm_bwd": [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "b200_embed_ln_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p],
    "b200_embed_ln_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p],
    "b200_cls_head_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "b200_ce_stats": [_p, _p, _p, _p, _i, _i, _p],
    "b200_cls_head_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "b200_colsum": [_p, _i, _p, _p, _i, _i, _p],
    "b200_cast_f32_to_f16": [_p, _p, _sz, _p],
    "b200_cast_f16_to_f32": [_p, _p, _sz, _p],
    "b200_scale_cast_grad": [_p, _p, _sz, _f, _p, _p, _p],
    "b200_unscale_cast_grad": [_p, _p, _sz, _p, _p],
    "b200_attn_bwd_workspace": [_i, _i, _i],
    "b200_ponet_workspace": [_i, _i, _i, _i, _i],
    "b200_ponet_mix_fwd": [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b200_ponet_bwd_workspace": [_i, _i, _i, _i, _i],
    "b200_ponet_mix_bwd": [_p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "b200_attn_bwd": [_p, _i, _i, _p, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "b200_grad_sumsq": [_p, _sz, _p, _p],
    "b200_clip_coef": [_p, _f, _f, _p, _p],
    "b200_adamw_step": [_p, _p, _p, _p, _p, _sz, _f, _f, _f, _f, _f, _f, _f, _p, _p],
    "b200_gemm_f16_drop": [_p, _i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _p, C.c_uint, _f, _p],
    "b200_layernorm_bwd_drop": [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, C.c_uint, _f, _p],
    "b200_embed_ln_fwd_drop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, C.c_uint, _f, _p],
    "b200_embed_ln_bwd_drop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, C.c_uint, _f, _p],
    "b200_attn_fwd_drop": [_p, _i, _i, _p, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_attn_bwd_drop": [_p, _i, _i, _p, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_cls_head_fwd_drop": [_p, _p, _p, _p, _p, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_cls_head_bwd_drop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_set_hyper": [_p, _f, _f, _f, _f, _f, _f, _f, _p],
    "b200_adamw_step_dev": [_p, _p, _p, _p, _p, _sz, _p, _p, _p],
    "b200_adamw_step_dev_zero": [_p, _p, _p, _p, _p, _sz, _p, _p, _p],
}
_RESTYPE = {"b200_set_gemm_impl": None, "b200_set_gemm_debug": None, "b200_set_sm_limit": None, "b200_last_error": C.c_char_p, "b200_launch_count": _ll, "b200_version": _i, "b200_attn_bwd_workspace": _sz, "b200_ponet_workspace": _sz, "b200_ponet_bwd_workspace": _sz}

_lock = threading.Lock()
_lib = None


def build(verbose: bool = False) -> str:
    """Compile libb200enc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libb200enc.so failed (see output above)")
    return LIB_PATH


def load() -> C.CDLL:
    """Load the library (never builds implicitly on the GPU box: the .so travels with the tree)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(the B200 path has no CPU/PyTorch fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, argt in _PROTOS.items():
                fn = getattr(lib, name)      # AttributeError here == header/library mismatch: fail loudly
                fn.argtypes = argt
                fn.restype = _RESTYPE.get(name, _i)
            for name, rt in _RESTYPE.items():
                getattr(lib, name).restype = rt
            lib.b200_set_gemm_impl.argtypes = [_i]
            lib.b200_set_gemm_debug.argtypes = [_i]
            lib.b200_set_sm_limit.argtypes = [_i]
            _lib = lib
    return _lib


def exported_symbols():
    return sorted(set(_PROTOS) | set(_RESTYPE))


class B200Error(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200_last_error()
        raise B200Error(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().b200_launch_count())
