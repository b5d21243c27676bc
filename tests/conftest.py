"""pytest configuration: path setup, the `gpu` marker, shared fixtures."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference on seeded inputs (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    return dict(np.load(path, allow_pickle=False))


CSRC = os.path.join(ROOT, "pytorch-deepfepe_b200", "csrc")
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="session")
def shim():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "host_shim.so")
    src = os.path.join(ROOT, "tests", "host_shim.cpp")
    hdrs = [os.path.join(CSRC, h) for h in ("fepe_math.cuh", "fepe_fit_adjoint.cuh", "fepe_recover.cuh", "fepe_virt.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-x", "c++", "-I", CSRC, src, "-o", so])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.shim_eig9.argtypes = [dp, dp, dp]
    lib.shim_eig9.restype = ctypes.c_int
    lib.shim_eig9_multishift.argtypes = [dp, dp, dp]
    lib.shim_eig9_multishift.restype = ctypes.c_int
    lib.shim_eig9_multishift128.argtypes = [dp, dp, dp]
    lib.shim_eig9_multishift128.restype = ctypes.c_int
    lib.shim_pinv.argtypes = [dp, dp, ctypes.c_double, dp, dp]
    lib.shim_refine_step.argtypes = [dp, dp, ctypes.c_double, dp, dp]
    lib.shim_refine_step.restype = None
    lib.shim_svd3.argtypes = [dp, dp, dp, dp]
    lib.shim_svd3_direct.argtypes = [dp, dp, dp, dp]
    lib.shim_rank2.argtypes = [dp, dp]
    lib.shim_g36_index.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.shim_g36_index.restype = ctypes.c_int
    lib.shim_essential.argtypes = [dp] * 6
    lib.shim_quat.argtypes = [dp, dp]
    lib.shim_pose_adjoint.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, dp]
    lib.shim_rank2_adjoint.argtypes = [dp, dp, dp]
    lib.shim_fit_pair_fwd_bwd.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double] + [dp] * 8
    ip, fp, bp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_ubyte)
    lib.shim_recover_pose.argtypes = [dp, dp, fp, ctypes.c_int, ctypes.c_double, dp, dp, dp, ip, ip, dp, bp]
    lib.shim_tridiag9.argtypes = [dp, dp, dp, dp]
    for name in ("shim_eig9_tri32", "shim_eig9_tri64", "shim_eig9_tri_serial"):
        getattr(lib, name).argtypes = [dp, dp, dp]
        getattr(lib, name).restype = ctypes.c_int
    lib.shim_correct_matches.argtypes = [dp, fp, fp, ctypes.c_int, fp, fp]
    lib.shim_correct_matches.restype = ctypes.c_int
    lib.shim_solve_poly6.argtypes = [dp, dp]
    lib.shim_solve_poly6.restype = ctypes.c_int
    lib.shim_gt_from_motion.argtypes = [fp, fp, dp]
    return lib


@pytest.fixture
def dispatch():
    """Force a kernel variant through the C ABI's explicit hook (fepe_set_dispatch, include/fepe_b200.h); everything is
    back on automatic when the test ends.  Usage: dispatch("fit", "split"), dispatch("mlp_gemm", "tile"), ..."""
    from fepe_b200 import _lib
    which = {"fit": _lib.DISPATCH_FIT, "gram_team": _lib.DISPATCH_GRAM_TEAM, "mlp_gemm": _lib.DISPATCH_MLP_GEMM,
             "mlp_fuse": _lib.DISPATCH_MLP_FUSE, "wgrad": _lib.DISPATCH_WGRAD, "nn_dist": _lib.DISPATCH_NN_DIST}
    values = {"fit": {"auto": 0, "small": 1, "ring": 2, "split": 3},
              "gram_team": {"auto": 0, "1": 1, "2": 2, "4": 3},
              "mlp_gemm": {"auto": 0, "persist": 0, "tile": 1, "persist128": 2},
              "mlp_fuse": {"auto": 0, "1": 0, "2": 2},
              "wgrad": {"auto": 0, "128": 1, "256": 2},
              "nn_dist": {"auto": 0, "simt": 1, "tc": 2}}
    touched = set()

    def force(key, value):
        st = _lib.lib().fepe_set_dispatch(which[key], values[key][str(value)])
        assert st >= 0, st
        touched.add(key)

    yield force
    for key in touched:
        _lib.lib().fepe_set_dispatch(which[key], 0)
