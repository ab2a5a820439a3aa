"""GPU: NTT / INTT / coset variants through the C ABI, bit-exact against the oracle and the golden vectors."""
import os

import numpy as np
import pytest

from oracle import cref
from tests.util import rand_fr_mont, to_dev, to_host

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "vectors.npz"))
MODES = [(False, False), (True, False), (False, True), (True, True)]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_golden_vectors(ctx, log_n):
    x = cref.to_mont(cref.FR, np.ascontiguousarray(GOLD[f"ntt_in_{log_n}"]))
    for key, inv, cos in (("fwd", 0, 0), ("inv", 1, 0), ("cfwd", 0, 1), ("cinv", 1, 1)):
        got = cref.from_mont(cref.FR, ctx.ntt_host(x.copy(), log_n, inv, cos))
        assert np.array_equal(got, GOLD[f"ntt_{key}_{log_n}"]), (log_n, key)


# 1 pass (<= 2^11), 2 passes (<= 2^22), 3 passes (> 2^22); odd and even splits
@pytest.mark.parametrize("log_n", [4, 9, 11, 12, 13, 16, 18, 20])
def test_against_oracle_all_modes(ctx, log_n):
    x = rand_fr_mont(1 << log_n, 1000 + log_n)
    for inv, cos in MODES:
        got = ctx.ntt_host(x.copy(), log_n, inv, cos)
        assert np.array_equal(got, cref.ntt(x, log_n, inv, cos)), (log_n, inv, cos)


@pytest.mark.parametrize("log_n", [22, 23, 24])
def test_large_against_oracle(ctx, log_n):
    x = rand_fr_mont(1 << log_n, 77)
    d = to_dev(x)
    ctx.ntt_dev(d, log_n, False, True)
    assert np.array_equal(to_host(d), cref.ntt(x, log_n, False, True))
    ctx.ntt_dev(d, log_n, True, True)
    assert np.array_equal(to_host(d), x)


@pytest.mark.parametrize("kind", [1, 2])
def test_both_pass_kernels_give_the_same_results(ctx, kind):
    """zkb_ntt_set_kernel: the generic pass kernel everywhere (1) and the radix-4 kernel on every tile of 256..2048 elements
    (2; the default uses it on 2048-element tiles only) against the oracle: 2^16 = 8+8, 2^17 = 9+8, 2^19 = 10+9, 2^21 = 11+10,
    2^22 = 11+11 bits, all four modes, and a zero-padded batch."""
    ctx.ntt_set_kernel(kind)
    try:
        for log_n in (16, 17, 19, 21, 22):
            x = rand_fr_mont(1 << log_n, 2000 + log_n)
            for inv, cos in (MODES if log_n <= 19 else [(False, True), (True, True)]):
                d = to_dev(x)
                ctx.ntt_dev(d, log_n, inv, cos)
                assert np.array_equal(to_host(d), cref.ntt(x, log_n, inv, cos)), (log_n, inv, cos)
        log_n, length = 18, (1 << 16) + 3
        xs = [rand_fr_mont(1 << log_n, 2100 + k) for k in range(3)]
        ds = [to_dev(x) for x in xs]
        ctx.ntt_batch_dev(ds, log_n, False, True, length=length)
        for x, d in zip(xs, ds):
            ref_in = np.zeros_like(x)
            ref_in[:length] = x[:length]
            assert np.array_equal(to_host(d), cref.ntt(ref_in, log_n, False, True))
    finally:
        ctx.ntt_set_kernel(0)


def test_two_level_tables_give_the_same_results(ctx):
    """zkb_ntt_set_direct_tables(0): the small two-level twiddle path (used above 2^26 or when HBM is scarce)."""
    log_n = 14
    x = rand_fr_mont(1 << log_n, 9)
    ctx.ntt_set_direct_tables(False)
    try:
        for inv, cos in MODES:
            assert np.array_equal(ctx.ntt_host(x.copy(), log_n, inv, cos), cref.ntt(x, log_n, inv, cos))
    finally:
        ctx.ntt_set_direct_tables(True)


def test_zero_padding_like_fft_in_place_resize(ctx):
    """coset_evals_from_poly on the 4n domain: a degree < n+3 polynomial in a 4n buffer (quotient_poly.rs:52-96)."""
    log_n, length = 14, (1 << 12) + 3
    x = np.zeros((1 << log_n, 4), dtype=np.uint64)
    x[:length] = rand_fr_mont(length, 5)
    junk = x.copy()
    junk[length:] = rand_fr_mont((1 << log_n) - length, 6)      # must be ignored beyond `length`
    for cos in (False, True):
        got = ctx.ntt_host(junk.copy(), log_n, False, cos, length=length)
        assert np.array_equal(got, cref.ntt(x, log_n, False, cos))
    assert np.array_equal(cref.from_mont(cref.FR, ctx.ntt_host(
        np.concatenate([cref.to_mont(cref.FR, np.ascontiguousarray(GOLD["ntt_short_in"])), np.zeros((11, 4), np.uint64)]),
        4, length=5)), GOLD["ntt_short_fwd_4"])


def test_full_size_properties_2_24(ctx):
    """BASELINE config 3 upper size: round trip and linearity at 2^24 (size-independent properties)."""
    import torch
    log_n = 24
    n = 1 << log_n
    x = rand_fr_mont(n, 31)
    d = to_dev(x)
    ctx.ntt_dev(d, log_n, False, True)
    ctx.ntt_dev(d, log_n, True, True)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(d), x)
    # spot check: output k of the forward transform equals the DFT sum restricted to a sparse input
    sp = np.zeros((n, 4), dtype=np.uint64)
    idx = [0, 1, 12345, n - 1]
    sp[idx] = x[idx]
    d = to_dev(sp)
    ctx.ntt_dev(d, log_n, False, False)
    got = to_host(d[:4].contiguous())
    # oracle on the same sparse input (cheap enough once)
    assert np.array_equal(got, cref.ntt(sp, log_n)[:4])


def test_batch_and_domain_mirror(ctx):
    import zkt_plonk_b200 as z
    log_n = 12
    xs = [rand_fr_mont(1 << log_n, 50 + k) for k in range(9)]
    ds = [to_dev(x) for x in xs]
    ctx.ntt_batch_dev(ds, log_n, False, True)
    for x, d in zip(xs, ds):
        assert np.array_equal(to_host(d), cref.ntt(x, log_n, False, True))
    dom = z.GpuEvaluationDomain.new(3000, ctx)
    assert dom.size() == 4096
    short = xs[0][:3000]
    padded = np.zeros((4096, 4), dtype=np.uint64)
    padded[:3000] = short
    assert np.array_equal(dom.fft(short), cref.ntt(padded, 12))
    assert np.array_equal(dom.coset_fft(short), cref.ntt(padded, 12, False, True))
    assert np.array_equal(dom.ifft(xs[1]), cref.ntt(xs[1], 12, True))
    assert np.array_equal(dom.coset_ifft(xs[1]), cref.ntt(xs[1], 12, True, True))
    buf = xs[2].copy()
    dom.coset_fft_in_place(buf)
    dom.coset_ifft_in_place(buf)
    assert np.array_equal(buf, xs[2])


@pytest.mark.parametrize("log_n,inverse,length", [(18, False, None), (20, False, (1 << 18) + 3), (20, True, None), (22, False, (1 << 20) + 3)])
def test_batch_of_nine_against_oracle(ctx, log_n, inverse, length):
    """zkb_ntt_batch_dev at the prover's sizes: nine transforms in one launch per pass (round 4: nine zero-padded coset NTTs
    of 4n points from polynomials of <= n + 3 coefficients, quotient_poly.rs:52-96; rounds 1-3: batches of size-n iNTTs),
    each compared with the oracle's transform of the same (zero-padded) input."""
    n = 1 << log_n
    xs = [rand_fr_mont(n, 300 + 10 * log_n + k) for k in range(9)]
    ds = [to_dev(x) for x in xs]
    ctx.ntt_batch_dev(ds, log_n, inverse, True, length=length)
    for x, d in zip(xs, ds):
        ref_in = x
        if length is not None:
            ref_in = np.zeros_like(x)
            ref_in[:length] = x[:length]
        assert np.array_equal(to_host(d), cref.ntt(ref_in, log_n, inverse, True))


def test_error_codes(ctx):
    import zkt_plonk_b200 as z
    with pytest.raises(z.ZkbError) as e:
        ctx._check(ctx._lib.zkb_ntt_dev(ctx._h, 0, 1, 29, 0, 0))
    assert e.value.code in (-1, -2)
    d = to_dev(rand_fr_mont(8, 1))
    with pytest.raises(z.ZkbError) as e:
        ctx._check(ctx._lib.zkb_ntt_dev(ctx._h, d.data_ptr(), 8, 29, 0, 0))
    assert e.value.code == -2                       # Error::InvalidEvalDomainSize
    with pytest.raises(z.ZkbError):
        ctx._check(ctx._lib.zkb_ntt_dev(ctx._h, d.data_ptr(), 9, 3, 0, 0))
