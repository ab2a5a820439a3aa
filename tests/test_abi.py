"""CPU-only: the C-ABI library loads, exports every symbol include/zkb200.h declares, and refuses to run
without a CUDA device (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

import zkt_plonk_b200 as z
from zkt_plonk_b200 import _lib


def test_library_present_and_exports_all_declared_symbols():
    assert os.path.exists(z.LIB_PATH), "build libzkb200.so first (__graft_entry__.build())"
    names = z.declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(z.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    exported = subprocess.check_output(["nm", "-D", "--defined-only", z.LIB_PATH], text=True)
    undeclared = [l.split()[-1] for l in exported.splitlines() if " T zkb_" in l and l.split()[-1] not in names]
    assert not undeclared, undeclared


def test_version_and_null_handling():
    lib = _lib.lib()
    assert b"sm_100a" in lib.zkb_version()
    assert lib.zkb_ctx_create(0, None) == _lib.ZKB_ERR_INVALID
    assert lib.zkb_ctx_sync(None) == _lib.ZKB_ERR_INVALID
    assert lib.zkb_last_error(None) == b"null context"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(z.ZkbError):
        z.Context(0)


def test_product_never_imports_oracle():
    """The shipped package must not reference oracle/ (ROUND SPEC (3))."""
    pkg = os.path.dirname(z.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "zkb_oracle" not in src, f


def test_cpp_mirror_header_is_self_contained():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = '#include "zkb200.hpp"\nint main() { return sizeof(zkb::G1Affine) == 64 ? 0 : 1; }\n'
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(root, "include"), "-x", "c++", "-"],
                         input=src, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_domain_metadata_matches_arkworks_rules():
    from zkt_plonk_b200 import field
    D = z.GpuEvaluationDomain
    ctx = object()                                    # metadata needs no device
    d = D.new(5, ctx)
    assert d.size() == 8 and d.log_size() == 3
    assert D.new(1, ctx).size() == 1 and D.new(0, ctx).size() == 1
    assert D.new((1 << 28) + 1, ctx) is None          # -> Error::InvalidEvalDomainSize
    assert D.new(1 << 28, ctx).size() == 1 << 28
    w = d.group_gen()
    assert pow(w, 8, field.R_MOD) == 1 and pow(w, 4, field.R_MOD) == field.R_MOD - 1
    assert list(d.elements())[3] == d.element(3)
    assert d.evaluate_vanishing_polynomial(d.element(5)) == 0
