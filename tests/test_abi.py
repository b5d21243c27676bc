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
