"""The reference CLI's key files (`zkt compile` output: ck / pk / vk, bin/src/main.rs:96-113) through the C ABI.

Thin wrappers over csrc/keyfile.cu (host code, no GPU needed for the *_file_* calls): files hold ark-serialize 0.3
`serialize_unchecked` bytes (bin/src/parser.rs:5-29), memory holds this library's forms -- (k, 4) uint64 arrays of
Montgomery limbs for field elements, (k, 8) for affine points with (0, 0) as the identity ((k, 12) on the BLS12 curves: the
files are the same `derive(CanonicalSerialize)` layouts with 48-byte base-field elements; the curve is `field.use_curve`'s).
"""
import ctypes
import os

import numpy as np

from . import _lib, field

PK_ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "sigma1", "sigma2", "sigma3", "q_lookup", "q_table")   # keys/mod.rs:29-40
VK_ORDER = PK_ORDER                                                                                   # keys/mod.rs:264-274


def _check(rc, what):
    if rc != 0:
        raise _lib.ZkbError(rc, what)


def _path(p):
    return os.fsencode(p)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _aw():
    return 2 * field.FQ_WORDS             # words of an affine point


def ck_info(path):
    """(number of powers_of_g, max_degree) of a sonic_pc::CommitterKey file."""
    n, md = ctypes.c_size_t(), ctypes.c_size_t()
    _check(_lib.lib(field.CURVE).zkb_ck_file_info(_path(path), ctypes.byref(n), ctypes.byref(md)), f"{path}: not a CommitterKey file")
    return n.value, md.value


def ck_read(path, first=0, count=None):
    """powers_of_g[first : first + count] as (count, 8) Montgomery affine points."""
    if count is None:
        count = ck_info(path)[0] - first
    out = np.zeros((count, _aw()), dtype=np.uint64)
    _check(_lib.lib(field.CURVE).zkb_ck_file_read(_path(path), first, count, _vp(out)), f"{path}: cannot read {count} powers at {first}")
    return out


def ck_write(path, powers_xy, gamma_xy=None, max_degree=None):
    powers_xy = np.ascontiguousarray(powers_xy, dtype=np.uint64).reshape(-1, _aw())
    gamma_xy = np.zeros((0, _aw()), dtype=np.uint64) if gamma_xy is None else np.ascontiguousarray(gamma_xy, dtype=np.uint64).reshape(-1, _aw())
    md = powers_xy.shape[0] - 1 if max_degree is None else max_degree
    _check(_lib.lib(field.CURVE).zkb_ck_file_write(_path(path), _vp(powers_xy), powers_xy.shape[0], _vp(gamma_xy), gamma_xy.shape[0], md),
           f"{path}: cannot write")


def pk_read(path):
    """{name: (len, 4) Montgomery coefficients} in PK_ORDER."""
    lib = _lib.lib(field.CURVE)
    lens = (ctypes.c_size_t * 10)()
    _check(lib.zkb_pk_file_info(_path(path), lens), f"{path}: not a ProverKey file")
    bufs = [np.zeros((max(lens[k], 1), 4), dtype=np.uint64) for k in range(10)]
    ptrs = (ctypes.c_void_p * 10)(*[b.ctypes.data for b in bufs])
    caps = (ctypes.c_size_t * 10)(*[lens[k] for k in range(10)])
    _check(lib.zkb_pk_file_read(_path(path), ptrs, caps, lens), f"{path}: non-canonical coefficient")
    return {name: bufs[k][: lens[k]] for k, name in enumerate(PK_ORDER)}


def pk_write(path, polys):
    """polys: {name: (len, 4) Montgomery coefficients}; trailing zero coefficients are dropped as DensePolynomial does."""
    bufs = [np.ascontiguousarray(polys[name], dtype=np.uint64).reshape(-1, 4) for name in PK_ORDER]
    keep = [b if b.shape[0] else np.zeros((1, 4), dtype=np.uint64) for b in bufs]
    ptrs = (ctypes.c_void_p * 10)(*[b.ctypes.data for b in keep])
    lens = (ctypes.c_size_t * 10)(*[b.shape[0] for b in bufs])
    _check(_lib.lib(field.CURVE).zkb_pk_file_write(_path(path), ptrs, lens), f"{path}: cannot write")


def vk_read(path):
    """(n, pi_roots (k, 4) Montgomery, commitments (10, 8) Montgomery affine in VK_ORDER, is_inf list)."""
    lib = _lib.lib(field.CURVE)
    n, nr = ctypes.c_size_t(), ctypes.c_size_t()
    xy = np.zeros((10, _aw()), dtype=np.uint64)
    inf = (ctypes.c_int * 10)()
    _check(lib.zkb_vk_file_read(_path(path), ctypes.byref(n), None, 0, ctypes.byref(nr), _vp(xy), inf), f"{path}: not a VerifierKey file")
    roots = np.zeros((max(nr.value, 1), 4), dtype=np.uint64)
    _check(lib.zkb_vk_file_read(_path(path), ctypes.byref(n), _vp(roots), nr.value, ctypes.byref(nr), _vp(xy), inf), f"{path}: re-read failed")
    return n.value, roots[: nr.value], xy, [bool(x) for x in inf]


def vk_write(path, n, pi_roots, commits_xy, is_inf=None):
    roots = np.ascontiguousarray(pi_roots, dtype=np.uint64).reshape(-1, 4)
    keep = roots if roots.shape[0] else np.zeros((1, 4), dtype=np.uint64)
    xy = np.ascontiguousarray(commits_xy, dtype=np.uint64).reshape(10, _aw())
    inf = (ctypes.c_int * 10)(*[int(bool(x)) for x in (is_inf or [0] * 10)])
    _check(_lib.lib(field.CURVE).zkb_vk_file_write(_path(path), n, _vp(keep), roots.shape[0], _vp(xy), inf), f"{path}: cannot write")


def cvk_read(path):
    """(g, gamma_g, h, beta_h) of a sonic_pc::VerifierKey file: (8,), (8,), (16,), (16,) uint64 Montgomery arrays."""
    g, gg = np.zeros(_aw(), dtype=np.uint64), np.zeros(_aw(), dtype=np.uint64)
    h, bh = np.zeros(2 * _aw(), dtype=np.uint64), np.zeros(2 * _aw(), dtype=np.uint64)
    _check(_lib.lib(field.CURVE).zkb_cvk_file_read(_path(path), _vp(g), _vp(gg), _vp(h), _vp(bh)), f"{path}: not a sonic_pc::VerifierKey file")
    return g, gg, h, bh
