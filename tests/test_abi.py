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
