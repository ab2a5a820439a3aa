"""GPU: Pippenger MSM through the C ABI, bit-exact (affine x, y) against the oracle's VariableBaseMSM restatement."""
import os

import numpy as np
import pytest

from oracle import cref
from tests.util import gen_xy, gpu_points, rand_fr_mont, skewed_scalars, to_dev, to_host

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "vectors.npz"))


def test_fixed_base_mul_matches_oracle(ctx):
    n = 300
    k = cref.rand_fe(cref.FR, n, 3)
    k[0] = 0
    k[1] = [1, 0, 0, 0]
    import torch
    out = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(gen_xy(), to_dev(k), n, out)
    assert np.array_equal(to_host(out), cref.g1_mul(gen_xy(), k))


def test_golden_msm(ctx):
    P = cref.to_mont(cref.FQ, np.ascontiguousarray(GOLD["msm_points"])).reshape(-1, 8)
    s = np.ascontiguousarray(GOLD["msm_scalars"])
    for c in (0, 3, 7, 13, 16):
        ctx.set_msm_window(c)
        out, inf = ctx.msm_bases(P, s)
        assert not inf
        assert np.array_equal(cref.from_mont(cref.FQ, out.reshape(2, 4)), GOLD["msm_result"]), c
    ctx.set_msm_window(0)
    Pc = cref.to_mont(cref.FQ, np.ascontiguousarray(GOLD["msm_cancel_points"])).reshape(-1, 8)
    out, inf = ctx.msm_bases(Pc, np.ascontiguousarray(GOLD["msm_cancel_scalars"]))
    assert inf and not out.any()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 1000, 1 << 14])
def test_uniform_scalars_vs_oracle(ctx, n):
    _, P = gpu_points(ctx, n, 40 + n)
    s = cref.rand_fe(cref.FR, n, 41 + n)
    exp, einf = cref.msm_g1(P, s)
    got, inf = ctx.msm_bases(P, s)
    assert inf == einf and np.array_equal(got, exp)


def test_edge_cases_vs_oracle(ctx):
    n = 2048
    _, P = gpu_points(ctx, n, 7)
    s = cref.rand_fe(cref.FR, n, 8)
    s[0] = 0
    s[1] = [1, 0, 0, 0]
    s[2] = [0, 0, 0, 1 << 61]          # a single high bit (top window only)
    rm1 = cref.ints_to_limbs([0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000000])[0]
    s[3] = rm1                       # r - 1: every signed digit path incl. top window
    s[4] = [0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0, 0]   # long carry run through the windows
    P[10] = P[11]                    # repeated base (doubling inside a bucket when digits collide)
    s[10] = s[11]
    P[20] = 0                        # point at infinity in the bases
    exp, einf = cref.msm_g1(P, s)
    for c in (0, 4, 11, 16):
        ctx.set_msm_window(c)
        got, inf = ctx.msm_bases(P, s)
        assert inf == einf and np.array_equal(got, exp), c
    ctx.set_msm_window(0)
    # empty input and all-zero scalars give the identity
    got, inf = ctx.msm_bases(P[:0], s[:0])
    assert inf and not got.any()
    got, inf = ctx.msm_bases(P[:64], np.zeros((64, 4), dtype=np.uint64))
    assert inf and not got.any()
    # all-ones scalars: plain sum of the points
    ones = np.zeros((n, 4), dtype=np.uint64)
    ones[:, 0] = 1
    got, inf = ctx.msm_bases(P, ones)
    assert np.array_equal(got, cref.msm_g1(P, ones)[0])


def test_skewed_scalars_vs_oracle(ctx):
    """Witness-like distribution: huge buckets for digit 1 / small digits exercise the task splitting."""
    n = 1 << 15
    _, P = gpu_points(ctx, n, 17)
    s = skewed_scalars(n, 18)
    exp, einf = cref.msm_g1(P, s)
    got, inf = ctx.msm_bases(P, s)
    assert inf == einf and np.array_equal(got, exp)


def test_resident_srs_offsets_partials_and_commit(ctx):
    import zkt_plonk_b200 as z
    n = 5000
    dP, P = gpu_points(ctx, n, 23)
    ctx.srs_load(dP)
    assert ctx.srs_size() == n
    s = cref.rand_fe(cref.FR, n, 24)
    exp, _ = cref.msm_g1(P, s)
    got, _ = ctx.msm(s)                                   # host scalars
    assert np.array_equal(got, exp)
    got, _ = ctx.msm(to_dev(s))                           # device scalars
    assert np.array_equal(got, exp)
    # offset window (kzg10::commit skipping leading zero coefficients)
    exp_off, _ = cref.msm_g1(P[100:], s[:n - 100])
    got, _ = ctx.msm(s[:n - 100], offset=100)
    assert np.array_equal(got, exp_off)
    # point-range shards combine to the full result (the multi-GPU path run on one device)
    ds = to_dev(s)
    cuts = [0, 1250, 2500, 3750, n]
    parts = np.stack([ctx.msm_partial(ds[a:b].contiguous(), a, b - a) for a, b in zip(cuts[:-1], cuts[1:])])
    got, inf = z.sum_partials(parts)
    assert not inf and np.array_equal(got, exp)
    # commit of a Montgomery-form polynomial with leading/trailing zeros
    coeffs = cref.to_mont(cref.FR, s)
    coeffs[:7] = 0
    coeffs[n - 5:] = 0
    kzg = z.GpuKZG10(ctx)
    got, inf = kzg.commit_one(coeffs)
    canon = s.copy()
    canon[:7] = 0
    canon[n - 5:] = 0
    assert np.array_equal(got, cref.msm_g1(P, canon)[0])
    with pytest.raises(z.ZkbError) as e:
        ctx.msm(s, offset=1)
    assert e.value.code == -4


def test_pipelined_batch_commit_vs_oracle(ctx):
    """zkb_commit_batch_dev: several commitments in flight over two workspaces, plain and fixed-base bases."""
    n = 4096
    dP, P = gpu_points(ctx, n, 71)
    ctx.srs_load(dP)
    lens = [n, n - 3, 1, 17, n, 0, 2500]
    polys = [cref.to_mont(cref.FR, cref.rand_fe(cref.FR, max(l, 1), 80 + i)) for i, l in enumerate(lens)]
    exp = [cref.msm_g1(P[:l], cref.from_mont(cref.FR, p[:l])) if l else (np.zeros(8, dtype=np.uint64), True) for p, l in zip(polys, lens)]
    for fixed in (False, True):
        ctx.srs_precompute(0 if fixed else -1)
        got = ctx.commit_batch_dev([to_dev(p) for p in polys], lens)
        for (gxy, ginf), (exy, einf) in zip(got, exp):
            assert ginf == einf and np.array_equal(gxy, exy)
        # offsets (leading zero coefficients skipped by the caller)
        got = ctx.commit_batch_dev([to_dev(polys[0][5:]), to_dev(polys[1])], [n - 5, n - 3], offsets=[5, 0])
        sc = cref.from_mont(cref.FR, polys[0][5:])
        assert np.array_equal(got[0][0], cref.msm_g1(P[5:], sc)[0]) and np.array_equal(got[1][0], exp[1][0])
    ctx.srs_precompute(-1)


def test_fixed_base_tables_vs_oracle(ctx):
    """zkb_srs_precompute: shared-bucket MSM over the tables 2^(c*w) * P_i gives the same affine point."""
    n = 6000
    dP, P = gpu_points(ctx, n, 61)
    ctx.srs_load(dP)
    s_uni, s_skew = cref.rand_fe(cref.FR, n, 62), skewed_scalars(n, 63)
    rm1 = cref.ints_to_limbs([0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000000])[0]
    s_uni[5] = rm1
    s_uni[6] = [0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0]
    exp_uni, exp_skew = cref.msm_g1(P, s_uni)[0], cref.msm_g1(P, s_skew)[0]
    exp_off = cref.msm_g1(P[777:], s_uni[: n - 777])[0]
    for c in (0, 5, 9, 13, 16):
        ctx.srs_precompute(c)
        assert np.array_equal(ctx.msm(s_uni)[0], exp_uni), c
        assert np.array_equal(ctx.msm(to_dev(s_skew))[0], exp_skew), c
        assert np.array_equal(ctx.msm(s_uni[: n - 777], offset=777)[0], exp_off), c
        got, inf = ctx.msm(np.zeros((16, 4), dtype=np.uint64))
        assert inf and not got.any()
    # a forced window bypasses the tables; dropping them restores the plain path
    ctx.set_msm_window(11)
    assert np.array_equal(ctx.msm(s_uni)[0], exp_uni)
    ctx.set_msm_window(0)
    ctx.srs_precompute(-1)
    assert np.array_equal(ctx.msm(s_uni)[0], exp_uni)


def test_msm_2_20_vs_oracle(ctx):
    """BASELINE config 2: 2^20 random points and scalars, bit-exact against the VariableBaseMSM restatement."""
    n = 1 << 20
    dP, P = gpu_points(ctx, n, 101)
    ctx.srs_load(dP)
    cases = [(s, cref.msm_g1(P, s)) for s in (cref.rand_fe(cref.FR, n, 102), skewed_scalars(n, 103))]
    for s, (exp, einf) in cases:
        got, inf = ctx.msm(s)
        assert inf == einf and np.array_equal(got, exp)
    ctx.srs_precompute(0)                      # fixed-base tables, c from the cost model (20)
    for s, (exp, einf) in cases:
        got, inf = ctx.msm(s)
        assert inf == einf and np.array_equal(got, exp)
    ctx.srs_precompute(-1)


def test_msm_2_22_vs_oracle(ctx):
    """The north star's target size: a 2^22-point MSM, uniform scalars, bit-exact against the VariableBaseMSM restatement --
    plain bases (cost-model window) and fixed-base tables."""
    n = 1 << 22
    dP, P = gpu_points(ctx, n, 201)
    ctx.srs_load(dP)
    s = cref.rand_fe(cref.FR, n, 202)
    exp, einf = cref.msm_g1(P, s)
    got, inf = ctx.msm(s)
    assert inf == einf and np.array_equal(got, exp)
    ctx.srs_precompute(0)
    got, inf = ctx.msm(s)
    assert inf == einf and np.array_equal(got, exp)
    ctx.srs_precompute(-1)
    ctx.srs_load(dP[:8].contiguous())          # give the 256 MiB key back


def test_commit_push_finish_matches_batch(ctx):
    """zkb_commit_push / zkb_commit_finish (the incremental form of zkb_commit_batch_dev the round driver uses to
    overlap uploads with commitments): same commitments as the batch call and as the oracle, for more pushes than
    there are workspaces, an empty polynomial, an offset, with other work enqueued between pushes; an open batch
    blocks the batch entry point until it is finished."""
    import torch
    from zkt_plonk_b200._lib import ZkbError
    n = 3000
    d_pts, h_pts = gpu_points(ctx, n, 61)
    ctx.srs_load(d_pts)
    ctx.srs_precompute(0)
    polys = [rand_fr_mont(m, 70 + k) for k, m in enumerate((n, 1, 777, n - 5, 2048))]
    devs = [to_dev(p) for p in polys]
    lens = [p.shape[0] for p in polys]
    offs = [0, 5, 100, 0, 952]
    want = ctx.commit_batch_dev(devs, lens, offs)
    for k in range(len(polys)):
        ctx.commit_push(devs[k], lens[k], offs[k])
        if k == 1:
            ctx.commit_push(devs[0], 0, 0)                      # empty polynomial: the identity
        junk = torch.zeros(1 << 16, dtype=torch.int64, device="cuda") + k   # unrelated work between pushes
    with pytest.raises(ZkbError):
        ctx.commit_batch_dev(devs[:1], lens[:1])                # a push batch is open
    sc = cref.from_mont(cref.FR, polys[0])
    for call in (lambda: ctx.msm(to_dev(sc)), lambda: ctx.msm(sc), lambda: ctx.commit_dev(devs[0], 0, lens[0]),
                 lambda: ctx.msm_bases(h_pts[:16], sc[:16]), lambda: ctx.msm_partial(to_dev(sc), 0, n)):
        with pytest.raises(ZkbError):                           # single MSMs would reuse the batch's staging and result slot
            call()
    got = ctx.commit_finish(len(polys) + 1)
    empty = got.pop(2)
    assert empty[1] and not empty[0].any()
    for k in range(len(polys)):
        assert got[k][1] == want[k][1] and np.array_equal(got[k][0], want[k][0])
        exp, einf = cref.msm_g1(h_pts[offs[k]:offs[k] + lens[k]], cref.from_mont(cref.FR, polys[k]))
        assert got[k][1] == einf and np.array_equal(got[k][0], exp)
    assert ctx.commit_finish(0) == []
    again = ctx.commit_batch_dev(devs, lens, offs)              # and the batch entry point works again
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(again, want))
    ctx.srs_precompute(-1)


@pytest.mark.parametrize("tables", [False, True])
def test_host_scalar_msm_two_halves(ctx, tables):
    """zkb_msm_g1 with host scalars splits large MSMs into two point-range halves (the second half uploads while the
    first is accumulated): same point as the device-scalar entry and as the oracle, with an offset into the SRS."""
    n, off = (1 << 18) + 3, 5
    d_pts, h_pts = gpu_points(ctx, n + off, 81)
    ctx.srs_load(d_pts)
    if tables:
        ctx.srs_precompute(0)
    s = cref.rand_fe(cref.FR, n, 82)
    s[:1000] = 0
    got, inf = ctx.msm(s, offset=off)
    want, winf = ctx.msm(to_dev(s), offset=off, n=n)
    assert inf == winf and np.array_equal(got, want)
    exp, einf = cref.msm_g1(h_pts[off:], s)
    assert inf == einf and np.array_equal(got, exp)
    ctx.srs_precompute(-1)


@pytest.mark.parametrize("parts", [(2, 3), (3, 7), (7, 2)])
@pytest.mark.parametrize("tables", [False, True])
def test_msm_as_point_ranges_through_shared_buckets(ctx, parts, tables):
    """zkb_msm_set_parts: one MSM cut into point ranges that accumulate into ONE bucket array (sort of range k + 1 and the
    upload of host scalars under the accumulation of range k, one reduction): device and host scalars, uniform and witness-like
    scalars (oversized buckets: the combine kernels add to the bucket's earlier sum too), an offset into the key, sizes that do
    not divide by the number of ranges -- against the oracle and against the undivided MSM."""
    n, off = (1 << 13) + 5, 3
    d_pts, h_pts = gpu_points(ctx, n + off, 91)
    ctx.srs_load(d_pts)
    if tables:
        ctx.srs_precompute(0)
    try:
        for seed, skew in ((1, False), (2, True)):
            s = skewed_scalars(n, 92 + seed) if skew else cref.rand_fe(cref.FR, n, 92 + seed)
            exp, einf = cref.msm_g1(h_pts[off:], s)
            ctx.set_msm_parts(1, 1, 19)
            whole, winf = ctx.msm(to_dev(s), offset=off, n=n)
            ctx.set_msm_parts(parts[0], parts[1], 8)
            for _ in range(2):                                     # twice: the second call reuses every buffer and event
                got, inf = ctx.msm(to_dev(s), offset=off, n=n)
                assert inf == einf == winf and np.array_equal(got, exp) and np.array_equal(got, whole), (parts, tables, skew, "dev")
                got, inf = ctx.msm(s, offset=off)
                assert inf == einf and np.array_equal(got, exp), (parts, tables, skew, "host")
            t = ctx.msm_last_timing()
            assert t["entries"] == n * t["windows"] and t["accumulate_ms"] > 0
        z = np.zeros((n, 4), dtype=np.uint64)
        got, inf = ctx.msm(to_dev(z), offset=off, n=n)
        assert inf and not got.any()
    finally:
        ctx.set_msm_parts(1, 4, 19)
        ctx.srs_precompute(-1)


@pytest.mark.parametrize("rounds", [1, 3, 6])
def test_pair_rounds_vs_oracle(ctx, rounds):
    """zkb_msm_set_mode(rounds): batched-affine pair rounds (csrc/msm_pairs.cuh) in front of the XYZZ accumulation give the
    same affine point as the oracle -- edge cases that make SPECIAL pairs (a repeated base with equal digits: P + P; P and
    -P in one bucket; the point at infinity among the bases; all-ones scalars: one giant bucket reduced over every round),
    plain bases and fixed-base tables, uniform and witness-like scalars, windows small enough that buckets hold many entries."""
    ctx.set_msm_mode(rounds)
    try:
        n = 2048
        _, P = gpu_points(ctx, n, 7)
        s = cref.rand_fe(cref.FR, n, 8)
        rm1 = cref.ints_to_limbs([0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000000])[0]
        s[0] = 0
        s[1] = [1, 0, 0, 0]
        s[3] = rm1
        P[10] = P[11]
        s[10] = s[11]                # equal digits on equal points: doubling inside a pair
        P[12] = P[13]
        s[12] = [5, 0, 0, 0]
        s[13] = cref.ints_to_limbs([0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001 - 5])[0]   # -5: P and -P cancel
        P[20] = 0                    # point at infinity in the bases
        exp, einf = cref.msm_g1(P, s)
        for c in (0, 4, 7, 11):
            ctx.set_msm_window(c)
            got, inf = ctx.msm_bases(P, s)
            assert inf == einf and np.array_equal(got, exp), c
        ctx.set_msm_window(0)
        ones = np.zeros((n, 4), dtype=np.uint64)
        ones[:, 0] = 1
        got, inf = ctx.msm_bases(P, ones)
        assert np.array_equal(got, cref.msm_g1(P, ones)[0])
        got, inf = ctx.msm_bases(P[:64], np.zeros((64, 4), dtype=np.uint64))
        assert inf and not got.any()
        # resident SRS, fixed-base tables, 2^16 points: uniform and witness-like scalars, an offset into the key
        n = 1 << 16
        dP, P = gpu_points(ctx, n + 3, 170 + rounds)
        ctx.srs_load(dP)
        for tables in (False, True):
            if tables:
                ctx.srs_precompute(0)
            for sc in (cref.rand_fe(cref.FR, n, 171), skewed_scalars(n, 172)):
                got, inf = ctx.msm(to_dev(sc), offset=3)
                exp, einf = cref.msm_g1(P[3:], sc)
                assert inf == einf and np.array_equal(got, exp), (tables,)
                tm = ctx.msm_last_timing()
                assert tm["pair_rounds"] == rounds
        ctx.srs_precompute(-1)
    finally:
        ctx.set_msm_mode(0)
        ctx.set_msm_window(0)


def test_pair_rounds_2_20_against_the_xyzz_path(ctx):
    """North-star size: the 2^20-point fixed-base MSM with 3 pair rounds and in automatic mode equals the XYZZ-only result
    (which test_msm_2_20_vs_oracle pins to the oracle) on uniform and witness-like scalars."""
    n = 1 << 20
    dP, _ = gpu_points(ctx, n, 333)
    ctx.srs_load(dP)
    ctx.srs_precompute(0)
    try:
        for sc in (cref.rand_fe(cref.FR, n, 334), skewed_scalars(n, 335)):
            d = to_dev(sc)
            ctx.set_msm_mode(0)
            want = ctx.msm(d)
            for mode in (3, -1):
                ctx.set_msm_mode(mode)
                got = ctx.msm(d)
                assert got[1] == want[1] and np.array_equal(got[0], want[0]), mode
    finally:
        ctx.set_msm_mode(0)
        ctx.srs_precompute(-1)
