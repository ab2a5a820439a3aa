"""Loader for libzkb200.so (the C ABI of include/zkb200.h) and its per-curve builds.

One shared object per curve, same entry points (the reference's generics monomorphise per pairing engine, plonk.rs:226-254):
`lib()` / `lib("bn254")` is libzkb200.so (everything), `lib("bls12_381")` and `lib("bls12_377")` hold everything but
EthereumTranscript (bound to Bn254 upstream: ZKB_ERR_UNSUPPORTED there).

There is NO CPU fallback: if the shared library is missing the import fails loudly, and if no CUDA device
is present `Context()` raises.  Nothing under oracle/ is ever imported from here.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZKB200_LIB selects another build of the same sources (e.g. the experimental libzkb200_sqr.so); never a fallback.
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(_HERE, "libzkb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "zkb200.h")

CURVE_LIBS = {"bn254": LIB_PATH, "bls12_381": os.path.join(_HERE, "libzkb200_bls12_381.so"),
              "bls12_377": os.path.join(_HERE, "libzkb200_bls12_377.so")}
CURVE_IDS = {"bn254": 0, "bls12_381": 1, "bls12_377": 2}

ZKB_OK, ZKB_ERR_INVALID, ZKB_ERR_DOMAIN, ZKB_ERR_CUDA, ZKB_ERR_NO_SRS, ZKB_ERR_OOM, ZKB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


class ZkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"zkb200 error {code}: {msg}")
        self.code = code


def declared_symbols():
    """Every function name declared in include/zkb200.h."""
    with open(HEADER_PATH) as f:
        return sorted(set(re.findall(r"^ZKB_API [\w \*]*?\b(zkb_\w+)\(", f.read(), flags=re.M)))


def load(curve="bn254"):
    path = CURVE_LIBS[curve]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  zkt_plonk_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, sz, u, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int
    sig = {
        "zkb_ctx_create": (i, [i, ctypes.POINTER(vp)]),
        "zkb_ctx_destroy": (None, [vp]),
        "zkb_ctx_set_stream": (i, [vp, vp]),
        "zkb_ctx_sync": (i, [vp]),
        "zkb_last_error": (ctypes.c_char_p, [vp]),
        "zkb_version": (ctypes.c_char_p, []),
        "zkb_curve_info": (i, [ctypes.POINTER(i)] * 5),
        "zkb_g1_generator": (i, [vp]),
        "zkb_dev_alloc": (i, [vp, sz, ctypes.POINTER(vp)]),
        "zkb_dev_free": (i, [vp, vp]),
        "zkb_h2d": (i, [vp, vp, vp, sz]),
        "zkb_d2h": (i, [vp, vp, vp, sz]),
        "zkb_ntt": (i, [vp, vp, sz, u, i, i]),
        "zkb_ntt_dev": (i, [vp, vp, sz, u, i, i]),
        "zkb_ntt_batch_dev": (i, [vp, ctypes.POINTER(vp), sz, sz, u, i, i]),
        "zkb_ntt_set_direct_tables": (i, [vp, i]),
        "zkb_ntt_set_kernel": (i, [vp, i]),
        "zkb_srs_load_g1": (i, [vp, vp, sz]),
        "zkb_srs_load_g1_dev": (i, [vp, vp, sz]),
        "zkb_srs_size": (sz, [vp]),
        "zkb_srs_precompute": (i, [vp, i]),
        "zkb_msm_g1": (i, [vp, vp, sz, sz, vp, ctypes.POINTER(i)]),
        "zkb_msm_g1_dev": (i, [vp, vp, sz, sz, vp, ctypes.POINTER(i)]),
        "zkb_msm_g1_dev_partial": (i, [vp, vp, sz, sz, vp]),
        "zkb_msm_g1_sharded_dev": (i, [vp, vp, sz, sz, vp, ctypes.POINTER(i)]),
        "zkb_msm_g1_sharded": (i, [vp, vp, sz, sz, vp, ctypes.POINTER(i)]),
        "zkb_g1_sum_partials": (i, [vp, sz, vp, ctypes.POINTER(i)]),
        "zkb_msm_g1_bases": (i, [vp, vp, vp, sz, vp, ctypes.POINTER(i)]),
        "zkb_msm_g1_points_dev": (i, [vp, vp, vp, sz, i, vp, ctypes.POINTER(i)]),
        "zkb_ipa_round_lr_dev": (i, [vp, vp, vp, vp, sz, vp, vp, ctypes.POINTER(i), vp, ctypes.POINTER(i), vp, vp]),
        "zkb_ipa_final_key_dev": (i, [vp, vp, sz, vp, vp, ctypes.POINTER(i)]),
        "zkb_test_glv_split": (i, [vp, vp, ctypes.POINTER(i), vp, ctypes.POINTER(i)]),
        "zkb_ipa_round_fold_dev": (i, [vp, vp, vp, vp, sz, vp, vp]),
        "zkb_commit_batch_dev": (i, [vp, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz), sz, vp, ctypes.POINTER(i)]),
        "zkb_commit_dev": (i, [vp, vp, sz, sz, vp, ctypes.POINTER(i)]),
        "zkb_commit_push": (i, [vp, vp, sz, sz]),
        "zkb_commit_finish": (i, [vp, vp, ctypes.POINTER(i)]),
        "zkb_g1_fixed_base_mul_dev": (i, [vp, vp, vp, sz, vp]),
        "zkb_msm_set_window": (i, [vp, i]),
        "zkb_msm_set_mode": (i, [vp, i]),
        "zkb_msm_set_parts": (i, [vp, i, i, i]),
        "zkb_test_fp_binop": (i, [vp, i, i, vp, vp, vp, sz]),
        "zkb_z1_evals_dev": (i, [vp, u, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "zkb_z2_evals_dev": (i, [vp, u, vp, vp, vp, vp, vp, vp, vp]),
        "zkb_grand_product_failed": (i, [vp]),
        "zkb_quotient_evals_dev": (i, [vp, u, vp, ctypes.POINTER(vp), ctypes.POINTER(vp), vp]),
        "zkb_l1_coset_dev": (i, [vp, u, vp]),
        "zkb_poly_eval_dev": (i, [vp, vp, sz, vp, vp]),
        "zkb_poly_eval_many_dev": (i, [vp, sz, vp, vp, vp, vp]),
        "zkb_poly_lincomb_dev": (i, [vp, sz, ctypes.POINTER(vp), ctypes.POINTER(sz), vp, vp, sz]),
        "zkb_poly_divide_linear_dev": (i, [vp, vp, sz, vp, vp, vp]),
        "zkb_poly_add_blinders_dev": (i, [vp, vp, sz, vp, sz]),
        "zkb_poly_effective_len_dev": (i, [vp, vp, sz, ctypes.POINTER(sz)]),
        "zkb_plonk_setup": (i, [vp, u, ctypes.POINTER(vp), ctypes.POINTER(vp), sz, ctypes.POINTER(sz), sz, ctypes.POINTER(vp)]),
        "zkb_plonk_pk_destroy": (None, [vp, vp]),
        "zkb_plonk_pk_set_transcript": (i, [vp, i]),
        "zkb_plonk_proof_bytes": (sz, []),
        "zkb_plonk_pk_set_lookup_mode": (i, [vp, i]),
        "zkb_lookup_multisets_dev": (i, [vp, u, vp, sz, vp, vp, vp, vp, vp, vp, ctypes.POINTER(i)]),
        "zkb_test_transcript": (i, [i, vp, sz, vp, vp]),
        "zkb_plonk_vk_commitments": (i, [vp, vp, ctypes.POINTER(i)]),
        "zkb_plonk_prove": (i, [vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, ctypes.POINTER(ctypes.c_float)]),
        "zkb_plonk_pk_set_wiring": (i, [vp, vp, vp, vp, vp]),
        "zkb_plonk_prove_vars": (i, [vp, vp, vp, sz, vp, sz, vp, vp, vp, ctypes.POINTER(ctypes.c_float)]),
        "zkb_ck_file_info": (i, [ctypes.c_char_p, ctypes.POINTER(sz), ctypes.POINTER(sz)]),
        "zkb_ck_file_read": (i, [ctypes.c_char_p, sz, sz, vp]),
        "zkb_ck_file_write": (i, [ctypes.c_char_p, vp, sz, vp, sz, sz]),
        "zkb_cvk_file_read": (i, [ctypes.c_char_p, vp, vp, vp, vp]),
        "zkb_srs_load_ck_file": (i, [vp, ctypes.c_char_p, sz]),
        "zkb_pk_file_info": (i, [ctypes.c_char_p, ctypes.POINTER(sz)]),
        "zkb_pk_file_read": (i, [ctypes.c_char_p, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz)]),
        "zkb_pk_file_write": (i, [ctypes.c_char_p, ctypes.POINTER(vp), ctypes.POINTER(sz)]),
        "zkb_vk_file_read": (i, [ctypes.c_char_p, ctypes.POINTER(sz), vp, sz, ctypes.POINTER(sz), vp, ctypes.POINTER(i)]),
        "zkb_vk_file_write": (i, [ctypes.c_char_p, sz, vp, sz, vp, ctypes.POINTER(i)]),
        "zkb_plonk_pk_from_polys": (i, [vp, u, ctypes.POINTER(vp), ctypes.POINTER(sz), sz, ctypes.POINTER(sz), sz, vp,
                                    ctypes.POINTER(i), ctypes.POINTER(vp)]),
        "zkb_plonk_load_keys": (i, [vp, ctypes.c_char_p, ctypes.c_char_p, sz, ctypes.POINTER(vp)]),
        "zkb_plonk_save_keys": (i, [vp, vp, ctypes.c_char_p, ctypes.c_char_p]),
        "zkb_plonk_verify": (i, [sz, vp, sz, vp, ctypes.POINTER(i), vp, vp, vp, vp, i]),
        "zkb_pairing": (i, [vp, vp, vp]),
        "zkb_pairing_product_is_one": (i, [vp, vp, sz, ctypes.POINTER(i)]),
        "zkb_g2_mul": (i, [vp, vp, vp]),
        "zkb_launch_count": (ctypes.c_uint64, [vp]),
        "zkb_msm_last_timing": (i, [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint64)]),
        "zkb_msm_last_pair_rounds": (i, [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i)]),
        "zkb_bench_int": (i, [vp, i, ctypes.POINTER(ctypes.c_double)]),
        "zkb_comm_unique_id": (i, [vp]),
        "zkb_comm_init": (i, [vp, vp, i, i]),
        "zkb_comm_destroy": (i, [vp]),
        "zkb_comm_rank": (i, [vp]),
        "zkb_comm_world": (i, [vp]),
        "zkb_comm_allgather_host": (i, [vp, vp, sz, vp]),
        "zkb_srs_set_range": (i, [vp, sz, sz]),
        "zkb_srs_set_replicated": (i, [vp, i]),
        "zkb_commit_expect": (i, [vp, sz]),
        "zkb_commit_finish_partials": (i, [vp, vp]),
        "zkb_test_set_rank_world": (i, [vp, i, i]),
        "zkb_test_replicated_share": (i, [i, i, i, sz, sz, sz, sz, ctypes.POINTER(sz)]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(lib, name):
            continue          # optional newer symbols are bound lazily by their wrappers
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_LIBS = {}


def lib(curve="bn254"):
    if curve not in _LIBS:
        _LIBS[curve] = load(curve)
    return _LIBS[curve]


def curve_info(curve="bn254"):
    """(curve_id, fr_words, fq_words, fr_bits, has_prover) of a curve's library -- needs no GPU."""
    v = [ctypes.c_int(0) for _ in range(5)]
    rc = lib(curve).zkb_curve_info(*[ctypes.byref(x) for x in v])
    if rc != 0:
        raise ZkbError(rc, "zkb_curve_info failed")
    return tuple(x.value for x in v)
