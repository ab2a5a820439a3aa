"""ctypes loader for oracle/libzkb_oracle.so (TEST INFRASTRUCTURE ONLY -- see zkb_oracle.c header).

Arrays are numpy uint64 of shape (n, 4) (field elements, LE limbs) or (n, 8) (affine G1 x||y).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzkb_oracle.so")
FR, FQ = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "zkb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def to_mont(field, a):
    out = np.empty_like(a)
    lib().zko_to_mont(field, _p(out), _p(a), ctypes.c_size_t(a.size // 4))
    return out


def from_mont(field, a):
    out = np.empty_like(a)
    lib().zko_from_mont(field, _p(out), _p(a), ctypes.c_size_t(a.size // 4))
    return out


def normalize(field, a):
    lib().zko_normalize(field, _p(a), ctypes.c_size_t(a.size // 4))
    return a


def rand_fe(field, n, seed):
    """n pseudo-random canonical field elements in [0, p)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    return normalize(field, np.ascontiguousarray(a))


def binop(field, op, a, b=None):
    b = a if b is None else b
    out = np.empty_like(a)
    lib().zko_fp_binop(field, op, _p(out), _p(a), _p(b), ctypes.c_size_t(a.size // 4))
    return out


def ntt(data, log_n, inverse=False, coset=False, threads=0):
    out = np.ascontiguousarray(data.copy())
    assert out.shape == (1 << log_n, 4)
    rc = lib().zko_ntt(_p(out), log_n, int(inverse), int(coset), threads)
    assert rc == 0
    return out


def msm_g1(points, scalars, threads=0):
    n = min(points.shape[0], scalars.shape[0])
    out = np.zeros(8, dtype=np.uint64)
    inf = ctypes.c_int(0)
    lib().zko_msm_g1(_p(points), _p(scalars), ctypes.c_size_t(n), _p(out), ctypes.byref(inf), threads)
    return out, bool(inf.value)


def g1_mul(base_xy, scalars):
    n = scalars.shape[0]
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().zko_g1_mul(_p(base_xy), _p(scalars), ctypes.c_size_t(n), _p(out))
    return out


def g1_walk(start_xy, step_xy, n):
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().zko_g1_walk(_p(start_xy), _p(step_xy), ctypes.c_size_t(n), _p(out))
    return out


def g1_sum(points):
    out = np.zeros(8, dtype=np.uint64)
    lib().zko_g1_sum(_p(points), ctypes.c_size_t(points.shape[0]), _p(out))
    return out


def g1_on_curve(xy):
    return bool(lib().zko_g1_on_curve(_p(np.ascontiguousarray(xy))))


def z1_evals(log_n, beta, gamma, a, b, c, s1, s2, s3):
    out = np.empty((1 << log_n, 4), dtype=np.uint64)
    lib().zko_z1_evals(log_n, _p(beta), _p(gamma), _p(a), _p(b), _p(c), _p(s1), _p(s2), _p(s3), _p(out))
    return out


def z2_evals(log_n, delta, eps, f, t, h1, h2):
    out = np.empty((1 << log_n, 4), dtype=np.uint64)
    lib().zko_z2_evals(log_n, _p(delta), _p(eps), _p(f), _p(t), _p(h1), _p(h2), _p(out))
    return out


def epk_free_tables(log_n):
    n4 = 4 << log_n
    x, zh, l1 = (np.empty((n4, 4), dtype=np.uint64) for _ in range(3))
    lib().zko_epk_free_tables(log_n, _p(x), _p(zh), _p(l1))
    return x, zh, l1


WIT_ORDER = ("z1", "z2", "a", "b", "c", "pi", "t", "h1", "h2")
EPK_ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_lookup", "q_table", "sigma1", "sigma2", "sigma3", "x", "l1", "zh")


def quotient_evals(log_n, ch, wit, epk):
    """ch: (5,4) alpha,beta,gamma,delta,epsilon; wit/epk: dict name -> (4n,4)."""
    n4 = 4 << log_n
    out = np.empty((n4, 4), dtype=np.uint64)
    W = (ctypes.c_void_p * 9)(*[wit[k].ctypes.data for k in WIT_ORDER])
    E = (ctypes.c_void_p * 13)(*[epk[k].ctypes.data for k in EPK_ORDER])
    lib().zko_quotient_evals(log_n, _p(ch), W, E, _p(out))
    return out


def poly_eval(coeffs, z_limbs):
    """DensePolynomial::evaluate: coeffs (n, 4) Montgomery, z (4,) Montgomery -> (4,) Montgomery."""
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    z = np.ascontiguousarray(z_limbs, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().zko_poly_eval(_p(c), ctypes.c_size_t(c.shape[0]), _p(z), _p(out))
    return out


def poly_lincomb(polys, scalars, out_len):
    """sum_j scalars[j] * polys[j] (zero-extended) -> (out_len, 4); polys: list of (len_j, 4), scalars (k, 4), Montgomery."""
    keep = [np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4) for p in polys]
    ptrs = (ctypes.c_void_p * len(keep))(*[p.ctypes.data if p.shape[0] else None for p in keep])
    lens = (ctypes.c_size_t * len(keep))(*[p.shape[0] for p in keep])
    sc = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((out_len, 4), dtype=np.uint64)
    lib().zko_poly_lincomb(ctypes.c_size_t(len(keep)), ptrs, lens, _p(sc), _p(out), ctypes.c_size_t(out_len))
    return out


def poly_divide_linear(coeffs, z_limbs):
    """((p(X) - p(z)) / (X - z) as (n - 1, 4), p(z) as (4,)), Montgomery."""
    c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    z = np.ascontiguousarray(z_limbs, dtype=np.uint64)
    quot = np.zeros((max(c.shape[0] - 1, 1), 4), dtype=np.uint64)
    ev = np.zeros(4, dtype=np.uint64)
    lib().zko_poly_divide_linear(_p(c), ctypes.c_size_t(c.shape[0]), _p(z), _p(quot), _p(ev))
    return quot[: max(c.shape[0] - 1, 0)], ev


def num_threads():
    return lib().zko_num_threads()


# ---- int <-> limb helpers
def ints_to_limbs(vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            a[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return a


def limbs_to_ints(a):
    a = a.reshape(-1, 4)
    return [sum(int(a[i, j]) << (64 * j) for j in range(4)) for i in range(a.shape[0])]
