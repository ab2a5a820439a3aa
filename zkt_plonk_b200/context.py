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
    def __init__(self, device=0, stream=None, curve="bn254"):
        """curve: "bn254" (default; the whole library), "bls12_381" or "bls12_377" (field / NTT / MSM / polynomial kernels)."""
        self._lib = _lib.lib(curve)
        self.curve = curve
        _, self.fr_words, self.fq_words, self.fr_bits, self.has_prover = _lib.curve_info(curve)
        self.aff_words, self.xyzz_words = 2 * self.fq_words, 4 * self.fq_words      # an affine point / an XYZZ partial sum
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

    def ntt_set_kernel(self, kind):
        """0: radix-4 pass kernel on 2048-element tiles (default), 1: generic kernel everywhere, 2: radix-4 wherever it exists."""
        self._check(self._lib.zkb_ntt_set_kernel(self._h, int(kind)))

    def ntt_set_direct_tables(self, enable):
        self._check(self._lib.zkb_ntt_set_direct_tables(self._h, int(bool(enable))))

    # -- SRS / MSM
    def srs_load(self, points):
        """points: host (n, 8) uint64 array or CUDA tensor with n*8 8-byte words (affine x||y, Montgomery)."""
        if isinstance(points, np.ndarray):
            self._check(self._lib.zkb_srs_load_g1(self._h, _host_ptr(points), points.shape[0]))
        else:
            self._check(self._lib.zkb_srs_load_g1_dev(self._h, _dev_ptr(points), points.numel() // self.aff_words))

    def srs_load_ck_file(self, path, max_points=0):
        """powers_of_g of the reference CLI's committer-key file (`ck`, bin/src/main.rs:274) become the resident SRS."""
        import os
        self._check(self._lib.zkb_srs_load_ck_file(self._h, os.fsencode(path), int(max_points)))

    def set_msm_mode(self, mode):
        """Batched-affine pair rounds in front of the XYZZ bucket accumulation (csrc/msm_pairs.cuh): 0 = none, 1..6 = that
        many rounds, -1 = chosen per MSM.  Results are identical for every mode."""
        self._check(self._lib.zkb_msm_set_mode(self._h, int(mode)))

    def srs_precompute(self, c=0):
        """Build (c >= 0) or drop (c < 0) the fixed-base window tables of the resident SRS."""
        self._check(self._lib.zkb_srs_precompute(self._h, int(c)))

    def srs_size(self):
        return int(self._lib.zkb_srs_size(self._h))

    def msm(self, scalars, offset=0, n=None):
        """scalars: canonical, host (n,4) uint64 array or CUDA tensor.  Returns ((8,) uint64 affine, is_inf)."""
        out = np.zeros(self.aff_words, dtype=np.uint64)
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
        out = np.zeros(self.xyzz_words, dtype=np.uint64)
        self._check(self._lib.zkb_msm_g1_dev_partial(self._h, _dev_ptr(scalars_dev), offset, n, _host_ptr(out)))
        return out

    def msm_sharded(self, scalars_dev, offset=0, n=None):
        """Collective (every rank of the context's communicator): this rank's scalars against its resident SRS range; the
        partial sums are exchanged inside the library (NCCL) and every rank returns the same ((8,) affine, is_inf)."""
        out = np.zeros(self.aff_words, dtype=np.uint64)
        inf = ctypes.c_int(0)
        if isinstance(scalars_dev, np.ndarray):                    # host scalars: upload overlapped with the accumulation
            n = scalars_dev.shape[0] if n is None else n
            self._check(self._lib.zkb_msm_g1_sharded(self._h, _host_ptr(scalars_dev), offset, n, _host_ptr(out), ctypes.byref(inf)))
            return out, bool(inf.value)
        n = scalars_dev.numel() // 4 if n is None else n
        self._check(self._lib.zkb_msm_g1_sharded_dev(self._h, _dev_ptr(scalars_dev), offset, n, _host_ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_bases(self, points, scalars):
        n = min(points.shape[0], scalars.shape[0])
        out = np.zeros(self.aff_words, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.zkb_msm_g1_bases(self._h, _host_ptr(points), _host_ptr(scalars), n, _host_ptr(out),
                                               ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_points_dev(self, points_dev, scalars_dev, n, scalars_mont=True):
        """MSM over caller-held device points (n affine points) and device scalars (Montgomery coefficients by default)."""
        out = np.zeros(self.aff_words, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.zkb_msm_g1_points_dev(self._h, _dev_ptr(points_dev), _dev_ptr(scalars_dev), n, int(bool(scalars_mont)),
                                                    _host_ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    # -- inner-product-argument rounds (csrc/ipa.cu)
    def ipa_round_lr_dev(self, coeffs_dev, z_dev, key_dev, n, h_prime=None):
        """((L affine, is_inf), (R affine, is_inf), <c_r, z_l>, <c_l, z_r>) of one opening round over vectors of length n;
        h_prime: (aff_words,) uint64 affine Montgomery point, or None for the two MSMs without the h' terms."""
        l, r = np.zeros(self.aff_words, dtype=np.uint64), np.zeros(self.aff_words, dtype=np.uint64)
        li, ri = ctypes.c_int(0), ctypes.c_int(0)
        ipl, ipr = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        self._check(self._lib.zkb_ipa_round_lr_dev(self._h, _dev_ptr(coeffs_dev), _dev_ptr(z_dev), _dev_ptr(key_dev), n,
                                                   _host_ptr(h_prime) if h_prime is not None else None, _host_ptr(l),
                                                   ctypes.byref(li), _host_ptr(r), ctypes.byref(ri), _host_ptr(ipl), _host_ptr(ipr)))
        return (l, bool(li.value)), (r, bool(ri.value)), ipl, ipr

    def ipa_final_key_dev(self, key_dev, n, challenges_mont):
        """<h, G> for the check polynomial of the round challenges ((log2 n, 4) uint64 Montgomery, proof order): (affine, is_inf)."""
        out = np.zeros(self.aff_words, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.zkb_ipa_final_key_dev(self._h, _dev_ptr(key_dev), n, _host_ptr(challenges_mont), _host_ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def ipa_round_fold_dev(self, coeffs_dev, z_dev, key_dev, n, x_mont, x_inv_mont):
        """c_l += x^-1 c_r, z_l += x z_r, G_l += x G_r in place (x, x_inv: (4,) uint64 Montgomery)."""
        self._check(self._lib.zkb_ipa_round_fold_dev(self._h, _dev_ptr(coeffs_dev), _dev_ptr(z_dev), _dev_ptr(key_dev), n,
                                                     _host_ptr(x_mont), _host_ptr(x_inv_mont)))

    def commit_dev(self, coeffs_mont_dev, offset, n):
        out = np.zeros(self.aff_words, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.zkb_commit_dev(self._h, _dev_ptr(coeffs_mont_dev), offset, n, _host_ptr(out),
                                             ctypes.byref(inf)))
        return out, bool(inf.value)

    def commit_batch_dev(self, coeffs_list, lens, offsets=None):
        """Pipelined kzg10::commit of several HBM-resident polynomials.  Returns [((8,) affine, is_inf), ...]."""
        k = len(coeffs_list)
        P = (ctypes.c_void_p * k)(*[_dev_ptr(t).value for t in coeffs_list])
        L = (ctypes.c_size_t * k)(*lens)
        O = (ctypes.c_size_t * k)(*(offsets if offsets is not None else [0] * k))
        out = np.zeros((k, self.aff_words), dtype=np.uint64)
        inf = (ctypes.c_int * k)()
        self._check(self._lib.zkb_commit_batch_dev(self._h, P, O, L, k, _host_ptr(out), inf))
        return [(out[j].copy(), bool(inf[j])) for j in range(k)]

    # -- multi-GPU (one process per GPU; commitments sharded by point range)
    def comm_init(self, group=None):
        """Collective: attach an NCCL communicator spanning `group` (default: the world) to this context.  Rank 0
        creates the NCCL id, torch.distributed (any backend) broadcasts its 128 bytes."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        idb = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            rc = self._lib.zkb_comm_unique_id(ctypes.c_void_p(idb.ctypes.data))
            if rc != 0:
                raise ZkbError(rc, "zkb_comm_unique_id failed (libnccl.so.2 not loadable?)")
        t = torch.from_numpy(idb)
        if dist.get_backend(group) == "nccl":
            t = t.to(f"cuda:{self.device}")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idb = np.ascontiguousarray(t.cpu().numpy())
        self._check(self._lib.zkb_comm_init(self._h, ctypes.c_void_p(idb.ctypes.data), rank, world))
        return rank, world

    def comm_destroy(self):
        self._check(self._lib.zkb_comm_destroy(self._h))

    def comm_rank_world(self):
        return int(self._lib.zkb_comm_rank(self._h)), int(self._lib.zkb_comm_world(self._h))

    def comm_allgather(self, arr):
        """arr: host numpy array (same shape on every rank) -> (world, *arr.shape)."""
        arr = np.ascontiguousarray(arr)
        _, world = self.comm_rank_world()
        out = np.empty((world,) + arr.shape, dtype=arr.dtype)
        self._check(self._lib.zkb_comm_allgather_host(self._h, ctypes.c_void_p(arr.ctypes.data), arr.nbytes,
                                                      ctypes.c_void_p(out.ctypes.data)))
        return out

    def srs_set_range(self, global_lo, global_n):
        """The resident SRS is [global_lo, global_lo + resident) of a key of global_n powers (commitments shard)."""
        self._check(self._lib.zkb_srs_set_range(self._h, int(global_lo), int(global_n)))

    def srs_set_replicated(self, fanout=-1):
        """Every rank holds the whole committer key; the library splits each batch of commitments among the ranks (fanout:
        -1 cost model, 0 always shard every commitment over all ranks, 1 always one group of ranks per commitment)."""
        self._check(self._lib.zkb_srs_set_replicated(self._h, int(fanout)))

    def commit_expect(self, count):
        """Size of the next commit_push batch (lets a replicated key fan the batch out over the ranks)."""
        self._check(self._lib.zkb_commit_expect(self._h, int(count)))

    def commit_push(self, coeffs_dev, length, offset=0):
        """Enqueue the MSM of one more HBM-resident polynomial of the open batch (returns at once)."""
        self._check(self._lib.zkb_commit_push(self._h, _dev_ptr(coeffs_dev), offset, length))

    def commit_finish(self, count):
        """Wait for the open batch of `count` pushed polynomials; [((8,) affine, is_inf), ...] in push order."""
        out = np.zeros((max(count, 1), self.aff_words), dtype=np.uint64)
        inf = (ctypes.c_int * max(count, 1))()
        self._check(self._lib.zkb_commit_finish(self._h, _host_ptr(out), inf))
        return [(out[j].copy(), bool(inf[j])) for j in range(count)]

    def commit_finish_partials(self, count):
        """Close the open batch without the exchange between ranks: (count, 16) XYZZ partial sums of this rank."""
        out = np.zeros((max(count, 1), self.xyzz_words), dtype=np.uint64)
        self._check(self._lib.zkb_commit_finish_partials(self._h, _host_ptr(out)))
        return out[:count]

    def set_msm_parts(self, dev_parts=1, host_parts=4, min_log=19):
        """A large single MSM as point ranges through shared buckets (zkb_msm_set_parts); 1 = the whole MSM at once."""
        self._check(self._lib.zkb_msm_set_parts(self._h, int(dev_parts), int(host_parts), int(min_log)))

    def set_msm_window(self, c):
        self._check(self._lib.zkb_msm_set_window(self._h, int(c)))

    def g1_fixed_base_mul_dev(self, base_xy, scalars_dev, n, out_dev):
        self._check(self._lib.zkb_g1_fixed_base_mul_dev(self._h, _host_ptr(base_xy), _dev_ptr(scalars_dev), n,
                                                        _dev_ptr(out_dev)))
        return out_dev

    # -- grand products / quotient / polynomial utilities (device-resident)
    def z1_evals_dev(self, log_n, beta, gamma, a, b, c, s1, s2, s3, out):
        self._check(self._lib.zkb_z1_evals_dev(self._h, log_n, _host_ptr(beta), _host_ptr(gamma), _dev_ptr(a), _dev_ptr(b),
                                               _dev_ptr(c), _dev_ptr(s1), _dev_ptr(s2), _dev_ptr(s3), _dev_ptr(out)))
        return out

    def z2_evals_dev(self, log_n, delta, epsilon, f, t, h1, h2, out):
        self._check(self._lib.zkb_z2_evals_dev(self._h, log_n, _host_ptr(delta), _host_ptr(epsilon), _dev_ptr(f), _dev_ptr(t),
                                               _dev_ptr(h1), _dev_ptr(h2), _dev_ptr(out)))
        return out

    def grand_product_failed(self):
        return bool(self._lib.zkb_grand_product_failed(self._h))

    def quotient_evals_dev(self, log_n, challenges, wit, epk, out):
        """challenges: host (5,4) alpha,beta,gamma,delta,epsilon; wit: 9 device tensors (z1,z2,a,b,c,pi,t,h1,h2);
        epk: 11 device tensors (q_m,q_l,q_r,q_o,q_c,q_lookup,q_table,sigma1,sigma2,sigma3,l1); all 4n elements."""
        W = (ctypes.c_void_p * 9)(*[_dev_ptr(t).value for t in wit])
        E = (ctypes.c_void_p * 11)(*[_dev_ptr(t).value for t in epk])
        self._check(self._lib.zkb_quotient_evals_dev(self._h, log_n, _host_ptr(challenges), W, E, _dev_ptr(out)))
        return out

    def l1_coset_dev(self, log_n, out):
        self._check(self._lib.zkb_l1_coset_dev(self._h, log_n, _dev_ptr(out)))
        return out

    def lookup_multisets_dev(self, log_n, table, q_lookup_evals, c_evals, t, f, h1, h2):
        """Round 2's witness plumbing on the device (csrc/lookup.cu): t = table || zeros, f = q_lookup * c and
        (h1, h2) = combine_split(t, f), n = 2^log_n elements each.  table: host (table_len, 4) Montgomery array; the rest are
        device tensors.  Returns the status word (0 ok, bit 0 ElementNotIndexedInTable, bit 1 halves not n long) after a sync."""
        table = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 4)
        status = ctypes.c_int(0)
        self._check(self._lib.zkb_lookup_multisets_dev(self._h, log_n, _host_ptr(table) if table.shape[0] else None, table.shape[0],
                                                       _dev_ptr(q_lookup_evals), _dev_ptr(c_evals), _dev_ptr(t), _dev_ptr(f),
                                                       _dev_ptr(h1), _dev_ptr(h2), ctypes.byref(status)))
        self.sync()
        return int(status.value)

    def poly_eval_dev(self, coeffs, n, z):
        out = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.zkb_poly_eval_dev(self._h, _dev_ptr(coeffs), n, _host_ptr(z), _host_ptr(out)))
        return out

    def poly_eval_many_dev(self, polys, lens, points):
        """polys[j] (device, lens[j] coefficients) evaluated at points[j] (k x 4 words, Montgomery form); returns k x 4 words."""
        k = len(polys)
        P = (ctypes.c_void_p * k)(*[_dev_ptr(t).value for t in polys])
        L = (ctypes.c_size_t * k)(*lens)
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(k, 4)
        out = np.zeros((k, 4), dtype=np.uint64)
        self._check(self._lib.zkb_poly_eval_many_dev(self._h, k, P, L, _host_ptr(pts), _host_ptr(out)))
        return out

    def poly_lincomb_dev(self, polys, lens, scalars, out, out_len):
        k = len(polys)
        P = (ctypes.c_void_p * k)(*[_dev_ptr(t).value for t in polys])
        L = (ctypes.c_size_t * k)(*lens)
        self._check(self._lib.zkb_poly_lincomb_dev(self._h, k, P, L, _host_ptr(scalars), _dev_ptr(out), out_len))
        return out

    def poly_divide_linear_dev(self, coeffs, n, z, quot):
        ev = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.zkb_poly_divide_linear_dev(self._h, _dev_ptr(coeffs), n, _host_ptr(z), _dev_ptr(quot),
                                                         _host_ptr(ev)))
        return ev

    def poly_add_blinders_dev(self, coeffs, length, blinders):
        self._check(self._lib.zkb_poly_add_blinders_dev(self._h, _dev_ptr(coeffs), length, _host_ptr(blinders),
                                                        blinders.shape[0]))
        return coeffs

    def poly_effective_len_dev(self, coeffs, n):
        out = ctypes.c_size_t(0)
        self._check(self._lib.zkb_poly_effective_len_dev(self._h, _dev_ptr(coeffs), n, ctypes.byref(out)))
        return int(out.value)

    def launch_count(self):
        return int(self._lib.zkb_launch_count(self._h))

    def msm_last_timing(self):
        """dict of device-side phase times (ms) of the last MSM plus its plan."""
        ms = (ctypes.c_float * 5)()
        info = (ctypes.c_uint64 * 3)()
        self._check(self._lib.zkb_msm_last_timing(self._h, ms, info))
        pm, pr = ctypes.c_float(0), ctypes.c_int(0)
        self._check(self._lib.zkb_msm_last_pair_rounds(self._h, ctypes.byref(pm), ctypes.byref(pr)))
        return {"sort_ms": ms[0] - pm.value, "pair_rounds_ms": pm.value, "pair_rounds": int(pr.value), "accumulate_ms": ms[1],
                "heavy_ms": ms[2], "reduce_ms": ms[3], "total_ms": ms[4],
                "entries": int(info[0]), "c": int(info[1]), "windows": int(info[2])}

    def bench_int(self, mode):
        v = ctypes.c_double(0)
        self._check(self._lib.zkb_bench_int(self._h, int(mode), ctypes.byref(v)))
        return v.value

    # -- test hook
    def proof_bytes(self):
        """bytes of a serialised Proof on this context's curve: 802 (BN254) or 1010 (BLS12-381 / BLS12-377)"""
        return int(self._lib.zkb_plonk_proof_bytes())

    def g1_generator(self):
        out = np.zeros(self.aff_words, dtype=np.uint64)
        self._check(self._lib.zkb_g1_generator(_host_ptr(out)))
        return out

    def fp_binop(self, field, op, a, b=None):
        """field 0: Fr ((n, 4) words), field 1: Fq ((n, fq_words) words)."""
        b = a if b is None else b
        out = np.empty_like(a)
        self._check(self._lib.zkb_test_fp_binop(self._h, field, op, _host_ptr(out), _host_ptr(a), _host_ptr(b),
                                                a.shape[0]))
        return out


def sum_partials(parts, curve="bn254"):
    """parts: (k, 16) uint64 XYZZ shard results -> ((8,) affine, is_inf) (24 / 12 words on the BLS12 curves).  Host-only
    (<= 8 group additions)."""
    fq_words = _lib.curve_info(curve)[2]
    parts = np.ascontiguousarray(parts, dtype=np.uint64).reshape(-1, 4 * fq_words)
    out = np.zeros(2 * fq_words, dtype=np.uint64)
    inf = ctypes.c_int(0)
    rc = _lib.lib(curve).zkb_g1_sum_partials(_host_ptr(parts), parts.shape[0], _host_ptr(out), ctypes.byref(inf))
    if rc != 0:
        raise ZkbError(rc, "zkb_g1_sum_partials failed")
    return out, bool(inf.value)
