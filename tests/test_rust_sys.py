"""CPU-only: rust/zkb200-sys/src/lib.rs (the `extern "C"` half of the Rust FFI crate) is generated from
include/zkb200.h; the committed file must be current and must declare every symbol the shared library exports."""
import importlib.util
import os
import re

from zkt_plonk_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_generated_bindings_are_current_and_complete():
    gen = _gen()
    with open(gen.OUT) as f:
        committed = f.read()
    assert committed == gen.generate(), "stale: run `python tools/gen_rust_sys.py`"
    declared = set(re.findall(r"pub fn (zkb_\w+)\(", committed))
    assert declared == set(_lib.declared_symbols())


def test_c_to_rust_type_mapping():
    rt = _gen().rust_type
    assert rt("zkb_ctx *ctx") == ("ctx", "*mut zkb_ctx")
    assert rt("const zkb_plonk_pk *pk") == ("pk", "*const zkb_plonk_pk")
    assert rt("zkb_plonk_pk **out") == ("out", "*mut *mut zkb_plonk_pk")
    assert rt("const uint64_t *const wit[9]") == ("wit", "*const *const u64")
    assert rt("uint64_t *const *ptrs_host") == ("ptrs_host", "*const *mut u64")
    assert rt("const uint64_t *const *polys_dev") == ("polys_dev", "*const *const u64")
    assert rt("uint64_t out_xy[8]") == ("out_xy", "*mut u64")
    assert rt("const uint64_t z[4]") == ("z", "*const u64")
    assert rt("uint8_t proof_out[802]") == ("proof_out", "*mut u8")
    assert rt("const char *path") == ("path", "*const c_char")
    assert rt("void *cuda_stream") == ("cuda_stream", "*mut c_void")
    assert rt("size_t n") == ("n", "usize") and rt("unsigned log_n") == ("log_n", "c_uint")
    assert rt("uint64_t *const coeffs_mont_out[10]") == ("coeffs_mont_out", "*const *mut u64")
