"""Context: one per GPU / proving thread.  Thin, exception-raising wrapper over the C ABI."""
import ctypes

import numpy as np

from . import _lib
from ._lib import ZkbError


def _host_ptr(a):
    if not isinstance(a, np.ndarray) or a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("expected a C-contiguous numpy uint64 array")
    return ctypes.c_void_p(a.ctypes.data)


def _dev_ptr(t):
    """Device pointer of a CUDA tensor (any 8-byte dtype) or a raw int address."""
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    if not t.is_cuda or not t.is_contiguous():
        raise TypeError("expected a contiguous CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


class Context:
    def __init__(self, device=0, stream=None):
        self._lib = _lib.lib()
        h = ctypes.c_void_p()
        rc = self._lib.zkb_ctx_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise ZkbError(rc, "zkb_ctx_create failed (no CUDA device? zkt_plonk_b200 has no CPU fallback)")
        self._h = h
        self.device = int(device)
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise ZkbError(rc, self._lib.zkb_last_error(self._h).decode())

    def set_stream(self, stream):
        """stream: a torch.cuda.Stream, a raw cudaStream_t integer, or None for the default stream."""
        ptr = 0 if stream is None else (stream if isinstance(stream, int) else stream.cuda_stream)
        self._check(self._lib.zkb_ctx_set_stream(self._h, ctypes.c_void_p(ptr)))

    def sync(self):
        self._check(self._lib.zkb_ctx_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.zkb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- NTT
    def ntt_host(self, data, log_n, inverse=False, coset=False, length=None):
        """In place on a host (2^log_n, 4) uint64 array; the first `length` rows are input."""
        n = 1 << log_n
        if data.shape != (n, 4):
            raise ValueError("data must have shape (2^log_n, 4)")
        self._check(self._lib.zkb_ntt(self._h, _host_ptr(data), n if length is None else length, log_n,
                                      int(inverse), int(coset)))
        return data

    def ntt_dev(self, t, log_n, inverse=False, coset=False, length=None):
        n = 1 << log_n
        self._check(self._lib.zkb_ntt_dev(self._h, _dev_ptr(t), n if length is None else length, log_n,
                                          int(inverse), int(coset)))
        return t

    def ntt_batch_dev(self, tensors, log_n, inverse=False, coset=False, length=None):
        n = 1 << log_n
        arr = (ctypes.c_void_p * len(tensors))(*[_dev_ptr(t).value for t in tensors])
        self._check(self._lib.zkb_ntt_batch_dev(self._h, arr, len(tensors), n if length is None else length, log_n,
                                                int(inverse), int(coset)))
        return tensors

    # -- SRS / MSM
    def srs_load(self, points):
        """points: host (n, 8) uint64 array or CUDA tensor with n*8 8-byte words (affine x||y, Montgomery)."""
        if isinstance(points, np.ndarray):
            self._check(self._lib.zkb_srs_load_g1(self._h, _host_ptr(points), points.shape[0]))
        else:
            self._check(self._lib.zkb_srs_load_g1_dev(self._h, _dev_ptr(points), points.numel() // 8))

    def srs_size(self):
        return int(self._lib.zkb_srs_size(self._h))

    def msm(self, scalars, offset=0, n=None):
        """scalars: canonical, host (n,4) uint64 array or CUDA tensor.  Returns ((8,) uint64 affine, is_inf)."""
        out = np.zeros(8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        if isinstance(scalars, np.ndarray):
            n = scalars.shape[0] if n is None else n
            rc = self._lib.zkb_msm_g1(self._h, _host_ptr(scalars), offset, n, _host_ptr(out), ctypes.byref(inf))
        else:
            n = scalars.numel() // 4 if n is None else n
            rc = self._lib.zkb_msm_g1_dev(self._h, _dev_ptr(scalars), offset, n, _host_ptr(out), ctypes.byref(inf))
        self._check(rc)
        return out, bool(inf.value)

    def msm_partial(self, scalars_dev, offset, n):
        out = np.zeros(16, dtype=np.uint64)
        self._check(self._lib.zkb_msm_g1_dev_partial(self._h, _dev_ptr(scalars_dev), offset, n, _host_ptr(out)))
        return out

    def msm_bases(self, points, scalars):
        n = min(points.shape[0], scalars.shape[0])
        out = np.zeros(8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.zkb_msm_g1_bases(self._h, _host_ptr(points), _host_ptr(scalars), n, _host_ptr(out),
                                               ctypes.byref(inf)))
        return out, bool(inf.value)

    def commit_dev(self, coeffs_mont_dev, offset, n):
        out = np.zeros(8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.zkb_commit_dev(self._h, _dev_ptr(coeffs_mont_dev), offset, n, _host_ptr(out),
                                             ctypes.byref(inf)))
        return out, bool(inf.value)

    def set_msm_window(self, c):
        self._check(self._lib.zkb_msm_set_window(self._h, int(c)))

    def g1_fixed_base_mul_dev(self, base_xy, scalars_dev, n, out_dev):
        self._check(self._lib.zkb_g1_fixed_base_mul_dev(self._h, _host_ptr(base_xy), _dev_ptr(scalars_dev), n,
                                                        _dev_ptr(out_dev)))
        return out_dev

    def launch_count(self):
        return int(self._lib.zkb_launch_count(self._h))

    def msm_last_timing(self):
        """dict of device-side phase times (ms) of the last MSM plus its plan."""
        ms = (ctypes.c_float * 5)()
        info = (ctypes.c_uint64 * 3)()
        self._check(self._lib.zkb_msm_last_timing(self._h, ms, info))
        return {"sort_ms": ms[0], "accumulate_ms": ms[1], "heavy_ms": ms[2], "reduce_ms": ms[3], "total_ms": ms[4],
                "entries": int(info[0]), "c": int(info[1]), "windows": int(info[2])}

    def bench_int(self, mode):
        v = ctypes.c_double(0)
        self._check(self._lib.zkb_bench_int(self._h, int(mode), ctypes.byref(v)))
        return v.value

    # -- test hook
    def fp_binop(self, field, op, a, b=None):
        b = a if b is None else b
        out = np.empty_like(a)
        self._check(self._lib.zkb_test_fp_binop(self._h, field, op, _host_ptr(out), _host_ptr(a), _host_ptr(b),
                                                a.shape[0]))
        return out


def sum_partials(parts):
    """parts: (k, 16) uint64 XYZZ shard results -> ((8,) affine, is_inf).  Host-only (<= 8 group additions)."""
    parts = np.ascontiguousarray(parts, dtype=np.uint64).reshape(-1, 16)
    out = np.zeros(8, dtype=np.uint64)
    inf = ctypes.c_int(0)
    rc = _lib.lib().zkb_g1_sum_partials(_host_ptr(parts), parts.shape[0], _host_ptr(out), ctypes.byref(inf))
    if rc != 0:
        raise ZkbError(rc, "zkb_g1_sum_partials failed")
    return out, bool(inf.value)
