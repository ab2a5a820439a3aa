"""GPU: the C++ host-side mirror (include/zkb200.hpp: GpuDomain, GpuKZG10, compute_z1_poly, compute_z2_poly,
extend_prover_key, quotient_compute) driven from a C++ program and checked bit for bit against the oracle."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_binary(ctx):
    exe = os.path.join(ROOT, "tests", "cpp", "test_mirror")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all C++ mirror checks passed" in res.stdout and "FAIL" not in res.stdout


def test_cpp_header_compiles_standalone():
    """zkb200.hpp must be self-contained C++17 (no CUDA headers needed by a host integrator)."""
    src = '#include "zkb200.hpp"\nint main() { return sizeof(zkb::Fr) == 32 ? 0 : 1; }\n'
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                         input=src, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
