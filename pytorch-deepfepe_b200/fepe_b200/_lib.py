"""ctypes binding of the C ABI declared in include/fepe_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) into
``pytorch-deepfepe_b200/lib/libfepe_b200.so``.  There is NO fallback: if the library is missing
or a call returns a non-zero status, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEPE_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libfepe_b200.so")

SAVED_DOUBLES = 64          # FEPE_SAVED_DOUBLES
POSE_OUT_FLOATS = 32        # FEPE_POSE_OUT_FLOATS
RECOVER_OUT_FLOATS = 24     # FEPE_RECOVER_OUT_FLOATS
GT_FLOATS = 32              # FEPE_GT_FLOATS

DISPATCH_FIT, DISPATCH_GRAM_TEAM, DISPATCH_MLP_GEMM, DISPATCH_MLP_FUSE, DISPATCH_WGRAD = 0, 1, 2, 3, 4   # FEPE_DISPATCH_*
DISPATCH_NN_DIST = 5

_lib = None

_c_f = ctypes.c_float
_c_i = ctypes.c_int
_c_p = ctypes.c_void_p

_SIGNATURES = {
    "fepe_version": (ctypes.c_char_p, []),
    "fepe_max_correspondences": (_c_i, []),
    "fepe_set_dispatch": (_c_i, [_c_i, _c_i]),
    "fepe_debug_trace": (_c_i, [_c_p, _c_i]),
    "fepe_debug_trace_count": (_c_i, []),
    "fepe_fit_fwd": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_f, _c_f, _c_f, _c_f, _c_f,
                            _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fepe_fit_bwd": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_f, _c_f, _c_f, _c_f, _c_f,
                            _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fepe_fit_bwd_coords": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_f, _c_f, _c_f, _c_f, _c_f,
                                   _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fepe_pose_fwd": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_f, _c_f, _c_f, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p,
                             _c_i, _c_f, _c_p, _c_p]),
    "fepe_fit_pose_fwd": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_f, _c_f, _c_f, _c_f, _c_f, _c_p, _c_p, _c_p, _c_p,
                                 _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_f, _c_p, _c_p]),
    "fepe_pose_bwd": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_f, _c_f, _c_f, _c_f, _c_p, _c_p, _c_p, _c_p, _c_i, _c_f,
                             _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fepe_recover_pose": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_f, _c_p, _c_p, _c_p, _c_p]),
    "fepe_gt_virt": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fepe_nn_match_workspace_bytes": (ctypes.c_size_t, [_c_i, _c_i, _c_i]),
    "fepe_nn_match": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "fepe_mlp_first": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_gemm": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_norm": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_f, _c_f, _c_p]),
    "fepe_mlp_last": (_c_i, [_c_p, _c_p, _c_f, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_scale_shift": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_f, _c_i, _c_p]),
    "fepe_mlp_last_norm": (_c_i, [_c_p, _c_p, _c_f, _c_p, _c_f, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_gemm_norm": (_c_i, [_c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_wgrad": (_c_i, [_c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_normbwd": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_f, _c_f, _c_p]),
    "fepe_mlp_last_bwd": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp_first_bwd": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_prepare_weights": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "fepe_mlp32_first": (_c_i, [_c_p, _c_f, _c_f, _c_f, _c_f, _c_p, _c_i, _c_p, _c_i, _c_p, _c_i, _c_p, _c_i,
                                _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_scale_shift": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_f, _c_i, _c_p]),
    "fepe_mlp32_gemm": (_c_i, [_c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i,
                               _c_p]),
    "fepe_mlp32_last": (_c_i, [_c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_last_bwd": (_c_i, [_c_p, _c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_normbwd": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_wgrad": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_f, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_first_bwd": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "fepe_mlp32_affine_grads": (_c_i, [_c_p, _c_i, _c_i, _c_p, _c_p, _c_p, _c_p]),
}

_ERRORS = {-1: "FEPE_E_BADARG (null pointer, misaligned buffer or non-positive size)",
           -2: "FEPE_E_TOOLARGE (N does not fit one shared-memory stage)",
           -3: "FEPE_E_NODEVICE (no sm_100 CUDA device)"}


def exported_symbols():
    """Names include/fepe_b200.h declares; tests check the library exports every one."""
    return list(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                "g.build()').  fepe_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            if not hasattr(handle, name):
                continue            # reported by tests/test_abi.py; calling it raises below
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


_DISPATCH_KEYS = {"fit": (DISPATCH_FIT, {"auto": 0, "small": 1, "ring": 2, "split": 3}),
                  "gram_team": (DISPATCH_GRAM_TEAM, {"auto": 0, "1": 1, "2": 2, "4": 3}),
                  "mlp_gemm": (DISPATCH_MLP_GEMM, {"auto": 0, "persist": 0, "tile": 1, "persist128": 2}),
                  "mlp_fuse": (DISPATCH_MLP_FUSE, {"auto": 0, "1": 0, "2": 2}),
                  "wgrad": (DISPATCH_WGRAD, {"auto": 0, "128": 1, "256": 2}),
                  "nn_dist": (DISPATCH_NN_DIST, {"auto": 0, "simt": 1, "tc": 2})}


def set_dispatch(key: str, value) -> int:
    """Force a kernel variant by name (tests, development scripts): set_dispatch("fit", "split"), ("fit", "auto"), ...
    -- a thin wrapper of fepe_set_dispatch (include/fepe_b200.h).  Returns the previous raw value."""
    which, values = _DISPATCH_KEYS[key]
    st = lib().fepe_set_dispatch(which, values[str(value)])
    if st < 0:
        raise RuntimeError(f"fepe_set_dispatch({key}, {value}) failed: {st}")
    return st


def check(status: int, what: str) -> None:
    if status == 0:
        return
    if status < 0:
        raise RuntimeError(f"{what}: {_ERRORS.get(status, status)}")
    raise RuntimeError(f"{what}: CUDA error {status}")
