"""ctypes binding of `libb200enc.so` (C ABI declared in include/b200enc.h).

The product path has NO fallback: if the shared library is missing or an entry point
fails, the call raises.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200enc.so")
CSRC = os.path.join(_HERE, "csrc")

EPI_STORE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_ADD, EPI_ATOMIC, EPI_BIAS_RES32, EPI_RESADD, EPI_STORE_DELTA = range(10)
DT_F16, DT_F32, DT_BF16 = 0, 1, 2

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
    "b200_embed_ln_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _ll, _p, _p, _p],
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
    "b200_clip_coef_scaled": [_p, _f, _f, _p, _p, _p, _i, _f, _f, _f, _f, _p],
    "b200_adamw_step": [_p, _p, _p, _p, _p, _sz, _f, _f, _f, _f, _f, _f, _f, _p, _p],
    "b200_gemm_f16_drop": [_p, _i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _p, C.c_uint, _f, _p],
    "b200_layernorm_bwd_drop": [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, C.c_uint, _f, _p],
    "b200_embed_ln_fwd_drop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, C.c_uint, _f, _p],
    "b200_embed_ln_bwd_drop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _ll, _p, _p, _p, C.c_uint, _f, _p],
    "b200_attn_fwd_drop": [_p, _i, _i, _p, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_attn_bwd_drop": [_p, _i, _i, _p, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_cls_head_fwd_drop": [_p, _p, _p, _p, _p, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_cls_head_bwd_drop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_set_hyper": [_p, _f, _f, _f, _f, _f, _f, _f, _p],
    "b200_adamw_step_dev": [_p, _p, _p, _p, _p, _sz, _p, _p, _p],
    "b200_adamw_step_dev_zero": [_p, _p, _p, _p, _p, _sz, _p, _p, _p],
    # bf16 operands (config 4)
    "b200_gemm_bf16": [_p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _i, _p],
    "b200_cast_f32_to_bf16": [_p, _p, _sz, _p],
    # packed rows
    "b200_gather_i64": [_p, _p, _i, _p, _p],
    "b200_unpack_rows": [_p, _p, _i, _i, _p, _p],
    "b200_attn_fwd_varlen": [_p, _i, _i, _i, _i, _p, _ll, _p, _i, _p, _i, _i, _i, _p, C.c_uint, _f, _p],
    "b200_attn_bwd_workspace_varlen": [_i, _i, _i, _ll],
    "b200_attn_bwd_varlen": [_p, _i, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _ll, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, C.c_uint, _f, _p],
    # loss heads on the labelled rows (csrc/heads.cuh)
    "b200_heads_compact": [_p, _ll, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "b200_heads_gather_keys": [_p, _p, _i, _p, _p],
    "b200_heads_topic_ids": [_p, _p, _i, _p, _p],
    "b200_heads_gather_rows": [_p, _p, _i, _i, _p, _p],
    "b200_heads_scatter_rows": [_p, _p, _i, _i, _f, _p, _p],
    "b200_heads_segmax_fwd": [_p, _p, _p, _p, _i, _i, _i, _p, _p],
    "b200_heads_segmax_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _p],
    "b200_heads_normalize": [_p, _i, _i, _f, _p, _p, _p],
    "b200_heads_normalize_bwd": [_p, _p, _p, _i, _i, _f, _p, _p, _p],
    "b200_heads_pair_cos_fwd": [_p, _p, _p, _p, _p, _i, _i, _f, _p, _p, _i, _p],
    "b200_heads_pair_cos_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _f, _p, _p],
    "b200_heads_bce_fwd": [_p, _p, _i, _p, _p],
    "b200_heads_bce_bwd": [_p, _p, _i, _f, _p, _p, _p],
    "b200_heads_cssl_matrix_fwd": [_p, _p, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p],
    "b200_heads_cssl_matrix_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _f, _p, _p],
    "b200_heads_cssl_list_fwd": [_p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _p, _p],
    "b200_heads_cssl_list_bwd": [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p],
    "b200_heads_rows_ce_fwd": [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p],
    "b200_heads_rows_ce_bwd": [_p, _p, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p],
    "b200_heads_cls_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "b200_heads_focal_stats": [_p, _p, _p, _f, _i, _i, _p, _p],
    "b200_heads_cls_bwd": [_p, _p, _p, _p, _p, _p, _f, _f, _p, _i, _i, _i, _p, _p, _p, _p],
    # fused epilogue variants
    "b200_gemm_f16_resadd": [_p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _p, C.c_uint, _f, _p],
    "b200_gemm_f16_dgelu_colsum": [_p, _i, _p, _i, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p],
    "b200_gemm_f16_dgrad_delta": [_p, _i, _p, _i, _i, _i, _i, _p, _i, _p, _i, _p, _i, _i, _p],
    "b200_attn_bwd_delta_ptr": [_p],
    "b200_attn_bwd_ext": [_p, _i, _i, _p, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, C.c_uint, _f, _i, _p],
}
_RESTYPE = {"b200_set_sm_limit": None, "b200_last_error": C.c_char_p, "b200_launch_count": _ll, "b200_version": _i, "b200_source_hash": C.c_char_p, "b200_attn_bwd_workspace": _sz, "b200_attn_bwd_workspace_varlen": _sz, "b200_ponet_workspace": _sz, "b200_ponet_bwd_workspace": _sz, "b200_attn_bwd_delta_ptr": _p}

_lock = threading.Lock()
_lib = None


def source_hash() -> str:
    """sha1 over everything libb200enc.so is compiled from (csrc/*.cu, *.cuh, *.inc, the Makefile, include/b200enc.h)."""
    import hashlib
    h = hashlib.sha1()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc")) or f == "Makefile")
    for path in [os.path.join(CSRC, f) for f in files] + [os.path.join(os.path.dirname(_HERE), "include", "b200enc.h")]:
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def built_hash() -> Optional[str]:
    """Source hash recorded inside the library file on disk (read without loading it); None if there is no library,
    'unknown' for a library from a hand-run `make`."""
    if not os.path.exists(LIB_PATH):
        return None
    with open(LIB_PATH, "rb") as fh:
        blob = fh.read()
    i = blob.find(b"B200SRC:")
    if i < 0:
        return "unknown"
    return blob[i + 8:blob.index(b"\0", i)].decode(errors="replace")


def is_stale() -> bool:
    b = built_hash()
    if b is None:
        return True
    if b == "unknown" or not os.path.isdir(CSRC):      # hand-run make / a binary-only deployment: nothing to compare against
        return False
    return b != source_hash()


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile libb200enc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).  The source hash is recorded
    in the library; nothing is rebuilt when the library on disk already carries the hash of this tree.  Concurrent callers
    (one process per GPU) serialise on a lock file."""
    import fcntl
    want = source_hash()
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and built_hash() == want:
            return LIB_PATH
        r = subprocess.run(["make", "-B", "-C", CSRC, f"SRC_HASH={want}"], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-4000:])
            print(r.stderr[-4000:])
        if r.returncode != 0:
            raise RuntimeError("building libb200enc.so failed (see output above)")
    return LIB_PATH


def load() -> C.CDLL:
    """Load the library.  There is no fallback: a missing library, or one built from other sources than this tree's, is
    rebuilt when nvcc is available (the GPU box has the same toolchain) and otherwise refused — never half-used."""
    global _lib
    with _lock:
        if _lib is None:
            if is_stale():
                import shutil
                if shutil.which("nvcc") is None or shutil.which("make") is None:
                    raise RuntimeError(
                        f"{LIB_PATH} is missing or was built from different sources and nvcc is not on PATH: run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` (the B200 path has no CPU/PyTorch fallback)")
                build()
            lib = C.CDLL(LIB_PATH)
            for name, argt in _PROTOS.items():
                fn = getattr(lib, name)      # AttributeError here == header/library mismatch: fail loudly
                fn.argtypes = argt
                fn.restype = _RESTYPE.get(name, _i)
            for name, rt in _RESTYPE.items():
                getattr(lib, name).restype = rt
            lib.b200_set_sm_limit.argtypes = [_i]
            _lib = lib
    return _lib


def is_loaded() -> bool:
    return _lib is not None


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
