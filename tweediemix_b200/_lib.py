"""ctypes binding of libtmx.so (the C ABI in include/tmx.h).

The library is built in-tree by ``python -m tweediemix_b200.build`` (or ``__graft_entry__.build()``).
There is no fallback of any kind: if the shared library is missing or a call fails, a
``RuntimeError`` is raised with ``tmx_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# $TMX_LIB_PATH selects another in-tree build of the same ABI (tuning variants compiled with other -D flags; see build.py)
LIB_PATH = os.environ.get("TMX_LIB_PATH") or os.path.join(_HERE, "lib", "libtmx.so")

F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_SILU = 0, 1
NCHW, NHWC = 0, 1
ROUND_FP32, ROUND_REF = 0, 1
EPI_NONE, EPI_GEGLU = 0, 1

_vp, _i, _f, _sz, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64

# name -> (restype, argtypes); kept in sync with include/tmx.h (tests/test_abi.py checks the header)
SIGNATURES = {
    "tmx_version": (_i, []),
    "tmx_last_error": (C.c_char_p, []),
    "tmx_init": (_i, [_i]),
    "tmx_tweedie_blend_ddim_fwd": (_i, [_vp, _vp, _vp, C.POINTER(_f), _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _i, _i, _vp]),
    "tmx_blend_partial_fwd": (_i, [_vp, _vp, C.POINTER(_f), C.POINTER(_i), _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "tmx_blend_finish_fwd": (_i, [_vp, _vp, _vp, C.POINTER(_f), _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _vp]),
    "tmx_groupnorm_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "tmx_groupnorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _i, _vp]),
    "tmx_groupnorm_set_variant": (_i, [_i]),
    "tmx_groupnorm_cat_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "tmx_groupnorm_launches": (_i, [_i, _i, _i, _i, _i]),
    "tmx_resadd_fwd": (_i, [_vp, _vp, _vp, _sz, _f, _i, _vp]),
    "tmx_cat_channels_fwd": (_i, [_vp, _vp, _vp, _sz, _i, _i, _i, _vp]),
    "tmx_upsample_nearest2x_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "tmx_bias_resadd_fwd": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _f, _i, _vp]),
    "tmx_resadd_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _f, _i, _vp]),
    "tmx_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _f, _i, _vp]),
    "tmx_geglu_fwd": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "tmx_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _f, _i, _vp]),
    "tmx_attn_set_variant": (_i, [_i]),
    "tmx_routed_linear_fwd": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "tmx_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _vp, C.POINTER(_vp), _i, _i, _vp, _i, _vp]),
    "tmx_linear_workspace_bytes": (_sz, []),
    "tmx_linear_set_variant": (_i, [_i]),
    "tmx_lora_t_fwd": (_i, [_vp, C.POINTER(_vp), _vp, _i, _i, _i, _i64, _i, _i, _vp]),
    "tmx_vpred_cfg_ddim_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _i, _i, _vp]),
    "tmx_frame_inject_fwd": (_i, [_vp, _i, _i, _sz, _f, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None
_inited = set()


def load() -> C.CDLL:
    """dlopen libtmx.so and bind every symbol of the ABI; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"libtmx.so not found at {LIB_PATH}: build it with `python -m tweediemix_b200.build` "
                    "(there is no CPU / PyTorch fallback for the tmx kernels)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)            # AttributeError if the symbol is not exported
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def last_error() -> str:
    return load().tmx_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def init(device: int) -> None:
    if device not in _inited:
        check(load().tmx_init(int(device)), "tmx_init")
        _inited.add(device)
