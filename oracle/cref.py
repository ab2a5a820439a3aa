"""ctypes loader for oracle/libzkb_oracle*.so (TEST INFRASTRUCTURE ONLY -- see zkb_oracle.c header).

One library per curve (zkb_oracle.c compiled with -DZKO_CURVE=0/1/2): `Oracle("bn254")`, `Oracle("bls12_381")`,
`Oracle("bls12_377")`.  The module-level functions are the BN254 oracle's (what every BN254 test uses).
Arrays are numpy uint64 of shape (n, 4) (Fr elements, LE limbs), (n, nq) (Fq elements: nq = 4 on BN254, 6 on the BLS12
curves) or (n, 2 nq) (affine G1 x||y).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CURVES = {"bn254": ("libzkb_oracle.so", 0), "bls12_381": ("libzkb_oracle_bls12_381.so", 1), "bls12_377": ("libzkb_oracle_bls12_377.so", 2)}
_SO = os.path.join(_HERE, "libzkb_oracle.so")
FR, FQ = 0, 1


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("zkb_oracle.c", "zko_field.inc", "zko_curve_params.h")]
    newest = max(os.path.getmtime(f) for f in srcs)
    sos = [os.path.join(_HERE, so) for so, _ in CURVES.values()]
    if force or any(not os.path.exists(so) or os.path.getmtime(so) < newest for so in sos):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """The restated CPU algorithms for one curve."""

    def __init__(self, curve="bn254"):
        build()
        so, cid = CURVES[curve]
        self.curve = curve
        self.lib = ctypes.CDLL(os.path.join(_HERE, so))
        got, nq, bits = ctypes.c_int(-1), ctypes.c_int(0), ctypes.c_int(0)
        self.lib.zko_curve_info(ctypes.byref(got), ctypes.byref(nq), ctypes.byref(bits))
        assert got.value == cid, (curve, got.value)
        self.nq, self.fr_bits = nq.value, bits.value

    def words(self, field):
        return self.nq if field == FQ else 4

    def to_mont(self, field, a):
        out = np.empty_like(a)
        self.lib.zko_to_mont(field, _p(out), _p(a), ctypes.c_size_t(a.size // self.words(field)))
        return out


    def from_mont(self, field, a):
        out = np.empty_like(a)
        self.lib.zko_from_mont(field, _p(out), _p(a), ctypes.c_size_t(a.size // self.words(field)))
        return out


    def normalize(self, field, a):
        self.lib.zko_normalize(field, _p(a), ctypes.c_size_t(a.size // self.words(field)))
        return a


    def rand_fe(self, field, n, seed):
        """n pseudo-random canonical field elements in [0, p)."""
        rng = np.random.default_rng(seed)
        a = rng.integers(0, 2**64, size=(n, self.words(field)), dtype=np.uint64)
        return self.normalize(field, np.ascontiguousarray(a))


    def binop(self, field, op, a, b=None):
        b = a if b is None else b
        out = np.empty_like(a)
        self.lib.zko_fp_binop(field, op, _p(out), _p(a), _p(b), ctypes.c_size_t(a.size // self.words(field)))
        return out


    def ntt(self, data, log_n, inverse=False, coset=False, threads=0):
        out = np.ascontiguousarray(data.copy())
        assert out.shape == (1 << log_n, 4)
        rc = self.lib.zko_ntt(_p(out), log_n, int(inverse), int(coset), threads)
        assert rc == 0
        return out


    def msm_g1(self, points, scalars, threads=0):
        n = min(points.shape[0], scalars.shape[0])
        out = np.zeros(2 * self.nq, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self.lib.zko_msm_g1(_p(points), _p(scalars), ctypes.c_size_t(n), _p(out), ctypes.byref(inf), threads)
        return out, bool(inf.value)


    def g1_mul(self, base_xy, scalars):
        n = scalars.shape[0]
        out = np.zeros((n, 2 * self.nq), dtype=np.uint64)
        self.lib.zko_g1_mul(_p(base_xy), _p(scalars), ctypes.c_size_t(n), _p(out))
        return out


    def g1_walk(self, start_xy, step_xy, n):
        out = np.zeros((n, 2 * self.nq), dtype=np.uint64)
        self.lib.zko_g1_walk(_p(start_xy), _p(step_xy), ctypes.c_size_t(n), _p(out))
        return out


    def g1_sum(self, points):
        out = np.zeros(2 * self.nq, dtype=np.uint64)
        self.lib.zko_g1_sum(_p(points), ctypes.c_size_t(points.shape[0]), _p(out))
        return out


    def g1_on_curve(self, xy):
        return bool(self.lib.zko_g1_on_curve(_p(np.ascontiguousarray(xy))))


    def z1_evals(self, log_n, beta, gamma, a, b, c, s1, s2, s3):
        out = np.empty((1 << log_n, 4), dtype=np.uint64)
        self.lib.zko_z1_evals(log_n, _p(beta), _p(gamma), _p(a), _p(b), _p(c), _p(s1), _p(s2), _p(s3), _p(out))
        return out


    def z2_evals(self, log_n, delta, eps, f, t, h1, h2):
        out = np.empty((1 << log_n, 4), dtype=np.uint64)
        self.lib.zko_z2_evals(log_n, _p(delta), _p(eps), _p(f), _p(t), _p(h1), _p(h2), _p(out))
        return out


    def epk_free_tables(self, log_n):
        n4 = 4 << log_n
        x, zh, l1 = (np.empty((n4, 4), dtype=np.uint64) for _ in range(3))
        self.lib.zko_epk_free_tables(log_n, _p(x), _p(zh), _p(l1))
        return x, zh, l1


    WIT_ORDER = ("z1", "z2", "a", "b", "c", "pi", "t", "h1", "h2")
    EPK_ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_lookup", "q_table", "sigma1", "sigma2", "sigma3", "x", "l1", "zh")


    def quotient_evals(self, log_n, ch, wit, epk):
        """ch: (5,4) alpha,beta,gamma,delta,epsilon; wit/epk: dict name -> (4n,4)."""
        n4 = 4 << log_n
        out = np.empty((n4, 4), dtype=np.uint64)
        W = (ctypes.c_void_p * 9)(*[wit[k].ctypes.data for k in self.WIT_ORDER])
        E = (ctypes.c_void_p * 13)(*[epk[k].ctypes.data for k in self.EPK_ORDER])
        self.lib.zko_quotient_evals(log_n, _p(ch), W, E, _p(out))
        return out


    def poly_eval(self, coeffs, z_limbs):
        """DensePolynomial::evaluate: coeffs (n, 4) Montgomery, z (4,) Montgomery -> (4,) Montgomery."""
        c = np.ascontiguousarray(coeffs, dtype=np.uint64)
        z = np.ascontiguousarray(z_limbs, dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.zko_poly_eval(_p(c), ctypes.c_size_t(c.shape[0]), _p(z), _p(out))
        return out


    def poly_lincomb(self, polys, scalars, out_len):
        """sum_j scalars[j] * polys[j] (zero-extended) -> (out_len, 4); polys: list of (len_j, 4), scalars (k, 4), Montgomery."""
        keep = [np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4) for p in polys]
        ptrs = (ctypes.c_void_p * len(keep))(*[p.ctypes.data if p.shape[0] else None for p in keep])
        lens = (ctypes.c_size_t * len(keep))(*[p.shape[0] for p in keep])
        sc = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros((out_len, 4), dtype=np.uint64)
        self.lib.zko_poly_lincomb(ctypes.c_size_t(len(keep)), ptrs, lens, _p(sc), _p(out), ctypes.c_size_t(out_len))
        return out


    def poly_divide_linear(self, coeffs, z_limbs):
        """((p(X) - p(z)) / (X - z) as (n - 1, 4), p(z) as (4,)), Montgomery."""
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        z = np.ascontiguousarray(z_limbs, dtype=np.uint64)
        quot = np.zeros((max(c.shape[0] - 1, 1), 4), dtype=np.uint64)
        ev = np.zeros(4, dtype=np.uint64)
        self.lib.zko_poly_divide_linear(_p(c), ctypes.c_size_t(c.shape[0]), _p(z), _p(quot), _p(ev))
        return quot[: max(c.shape[0] - 1, 0)], ev


    def num_threads(self):
        return self.lib.zko_num_threads()


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


_ORACLES = {}


def oracle(curve="bn254"):
    if curve not in _ORACLES:
        _ORACLES[curve] = Oracle(curve)
    return _ORACLES[curve]


def lib():
    return oracle("bn254").lib


def _bn254(name):
    def call(*args, **kwargs):
        return getattr(oracle("bn254"), name)(*args, **kwargs)
    call.__name__ = name
    return call


for _name, _val in list(vars(Oracle).items()):
    if _name.startswith("_") or _name == "words":
        continue
    globals()[_name] = _bn254(_name) if callable(_val) else _val
