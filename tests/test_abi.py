"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol the
header declares, and the host wrappers refuse to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from fepe_b200 import _lib
    return _lib


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "fepe_b200.h")).read()
    declared = set(re.findall(r"\b(fepe_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = built.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in include/fepe_b200.h but not exported: {missing}"
    assert set(built.exported_symbols()) == declared
    assert lib.fepe_version().decode().startswith("fepe_b200")


def _header_prototypes():
    """name -> (return kind, [parameter kinds]) parsed from include/fepe_b200.h; kind in {'p', 'i', 'f', 'z', 's'}
    (pointer, int, float, size_t, const char*)."""
    hdr = open(os.path.join(ROOT, "include", "fepe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(fepe_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()

        def kind(decl):
            decl = decl.strip()
            if "*" in decl:
                return "s" if re.match(r"const\s+char\s*\*", decl) and decl.count("*") == 1 and "(" not in decl and \
                    decl.split("*")[1].strip() == "" else "p"
            base = re.sub(r"\b(const|unsigned|signed)\b", "", decl).split()
            return {"int": "i", "float": "f", "size_t": "z"}[base[0]]
        plist = [] if params in ("", "void") else [kind(x) for x in params.split(",")]
        protos[name] = (kind(ret + " x") if "*" not in ret else "s", plist)
    return protos


def test_ctypes_signatures_match_the_header(built):
    """Every entry of fepe_b200/_lib.py::_SIGNATURES has the parameter count and the parameter classes (pointer / int /
    float / size_t) of its prototype in include/fepe_b200.h -- a mismatch would corrupt arguments silently."""
    import ctypes
    protos = _header_prototypes()
    assert set(protos) == set(built._SIGNATURES)
    cls = {ctypes.c_void_p: "p", ctypes.c_int: "i", ctypes.c_float: "f", ctypes.c_size_t: "z", ctypes.c_char_p: "s"}
    for name, (res, args) in built._SIGNATURES.items():
        want_res, want_args = protos[name]
        assert cls[res] == want_res, f"{name}: return type {res} vs header {want_res}"
        got = [cls[a] for a in args]
        assert got == want_args, f"{name}: ctypes {''.join(got)} vs header {''.join(want_args)}"


_INFO_ONLY = ("fepe_version", "fepe_max_correspondences", "fepe_nn_match_workspace_bytes",
              "fepe_debug_trace", "fepe_debug_trace_count")      # diagnostics: NULL is a valid argument (stop tracing)


def test_entry_points_reject_null_pointers_before_touching_cuda(built):
    """Error behaviour of the boundary, checked without a GPU: every compute entry point validates its arguments before
    its first CUDA call -- null buffers with positive sizes return FEPE_E_BADARG (-1) instead of a CUDA error or a fault
    -- and the solver / pose / matcher entry points treat an empty batch as a successful no-op."""
    import ctypes
    lib = built.lib()
    for name, (res, args) in built._SIGNATURES.items():
        if name in _INFO_ONLY:
            continue
        fn = getattr(lib, name)
        nulls = [0.0 if a is ctypes.c_float else (0 if a is ctypes.c_void_p else 128) for a in args]
        assert fn(*nulls) == -1, f"{name}: null pointers must give FEPE_E_BADARG"
        zeros = [0.0 if a is ctypes.c_float else 0 for a in args]
        want = -1 if name.startswith("fepe_mlp") else 0           # the MLP entry points require B > 0
        assert fn(*zeros) == want, f"{name}: empty batch"
    assert lib.fepe_nn_match_workspace_bytes(10, 10, 10) > 0
    assert lib.fepe_max_correspondences() in (-3,) or lib.fepe_max_correspondences() > 1000   # -3: no sm_100 device here


def test_mlp_entry_points_reject_bad_shapes(built):
    """Shape / alignment contract of the tensor-core MLP entry points (include/fepe_b200.h), again without a GPU."""
    lib = built.lib()
    p = 0x1000                                   # a non-null, 16-byte aligned dummy: validation fails before any access
    assert lib.fepe_mlp_gemm(p, p, 0, p, p, 2, 250, 200, 128, 256, 0) == -1        # Npad not a multiple of 128
    assert lib.fepe_mlp_gemm(p, p, 0, p, p, 2, 256, 300, 128, 256, 0) == -1        # more valid rows than padded rows
    assert lib.fepe_mlp_gemm(p, p, 0, p, p, 2, 256, 200, 100, 256, 0) == -1        # K not a multiple of 64
    assert lib.fepe_mlp_gemm(p, p, 0, p, p, 2, 256, 200, 128, 100, 0) == -1        # Co not a multiple of 64
    assert lib.fepe_mlp_gemm(p, p, p + 4, p, p, 2, 256, 200, 128, 256, 0) == -1    # bias not 16-byte aligned
    assert lib.fepe_mlp_gemm_norm(p, p, 0.01, p, 0, p, p, 2, 256, 200, 128, 192, 0) == -1    # fused variant: Co % 128
    assert lib.fepe_mlp_gemm_norm(p, p + 4, 0.01, p, 0, p, p, 2, 256, 200, 128, 256, 0) == -1  # ss not 16-byte aligned
    assert lib.fepe_mlp_gemm_norm(p, p, 1.5, p, 0, p, p, 2, 256, 200, 128, 256, 0) == -1     # slope outside (0, 1)
    assert lib.fepe_mlp_scale_shift(p, p, p, p, 2, 63, 200, 1e-5, 0, 0) == -1      # odd channel count
    assert lib.fepe_mlp_last_norm(p, p, 0.0, p, 0.0, p, p, 2, 200, 256, 256, 0) == -1        # slope outside (0, 1)
    assert lib.fepe_mlp_first(p, p, p, p, p, 2, 200, 256, 9, 64, 0) == -1          # more than 8 input channels
    assert lib.fepe_mlp_wgrad(p, p, p, 100, 128, 64, 0) == -1                      # M not a multiple of 64
    # the fp32-parity path (fepe_mlp32_*)
    assert lib.fepe_mlp32_gemm(p, p, 0.01, 0, p, p, p, 0, p, p, 2, 250, 200, 128, 256, 0) == -1     # Npad % 128
    assert lib.fepe_mlp32_gemm(p, p, 0.01, 0, p, p, p, 0, p, p, 2, 256, 200, 100, 256, 0) == -1     # K % 64
    assert lib.fepe_mlp32_gemm(p, p, 0.01, 0, p, p, p, 0, p, p, 2, 256, 200, 128, 100, 0) == -1     # Co % 64
    assert lib.fepe_mlp32_gemm(p, p + 4, 0.01, 0, p, p, p, 0, p, p, 2, 256, 200, 128, 256, 0) == -1  # ss alignment
    assert lib.fepe_mlp32_gemm(p, p, 1.5, 0, p, p, p, 0, p, p, 2, 256, 200, 128, 256, 0) == -1      # slope outside (0, 1]
    assert lib.fepe_mlp32_first(p, 1.0, 0.0, 1.0, 0.0, p, 13, 0, 0, 0, 0, 0, 0, p, p, p, p, 0, 2, 200, 256, 64, 0) == -1   # 17 channels
    assert lib.fepe_mlp32_first(0, 1.0, 0.0, 1.0, 0.0, 0, 0, 0, 0, 0, 0, 0, 0, p, p, p, p, 0, 2, 200, 256, 64, 0) == -1    # no channels
    assert lib.fepe_mlp32_last(p, p, 0.01, p, p, p, 0, 2, 200, 256, 256, 3, 0) == -1             # Co must be 1 or 4
    assert lib.fepe_mlp32_last(p, p, 0.01, p, p, p, p, 2, 200, 256, 256, 4, 0) == -1             # softmax only for Co = 1
    assert lib.fepe_mlp32_normbwd(p, p, p, p, p, 0.01, p, p, p, 2, 256, 200, 192, 0) == -1       # C / 4 = 48: not a power of two
    assert lib.fepe_mlp32_wgrad(p, p, p, p, 0.01, p, 512, 256, 192, 64, 0) == -1                 # Co % 128
    assert lib.fepe_mlp32_wgrad(p, p, p, p, 0.01, p, 500, 256, 128, 64, 0) == -1                 # M not a multiple of Npad
    assert lib.fepe_mlp32_last_bwd(p, p, p, 0.01, p, p, p, p, 2, 200, 256, 256, 2, 0) == -1      # Co must be 1 or 4


def test_no_cpu_fallback(built):
    from fepe_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.fit_forward(torch.zeros(2, 8, 4), torch.zeros(2, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch-deepfepe_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{fn} mentions the oracle"
