"""GPU: the BLS12-381 / BLS12-377 builds of the library (libzkb200_bls12_381.so / libzkb200_bls12_377.so) through the C ABI,
bit-exact against the per-curve oracle: field operations over Fr (8 limbs) and Fq (12 limbs), NTT in all four modes, G1 MSM
(arbitrary bases, resident key with and without fixed-base tables, skewed scalars, 2^20 points against the closed form),
KZG commitments and the grand-product / quotient / polynomial kernels.  The reference runs its own full test on these two
curves (plonk-core/src/plonk.rs:226-254); the BN254 tests of the same kernels live in test_gpu_{ntt,msm,poly}.py.
"""
import ctypes

import numpy as np
import pytest

from oracle import cref
from tests.test_curves import CURVES, ints, limbs
from tests.util import to_dev, to_host

pytestmark = pytest.mark.gpu
BLS = ["bls12_381", "bls12_377"]
MODES = [(False, False), (True, False), (False, True), (True, True)]


@pytest.fixture(scope="module", params=BLS)
def cv(request):
    """(context of the curve's library, its oracle, curve constants)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import zkt_plonk_b200 as z
    c = z.Context(0, curve=request.param)
    c.set_stream(torch.cuda.current_stream())
    assert (c.fq_words, c.has_prover) == (6, 1)
    yield c, cref.oracle(request.param), CURVES[request.param]
    c.close()


def fr_mont(o, n, seed):
    return o.to_mont(cref.FR, o.rand_fe(cref.FR, n, seed))


def gpu_points(ctx, o, n, seed):
    """k_i * G built on the device: (device tensor, host (n, 12) array, the k_i as canonical limbs)"""
    import torch
    k = o.rand_fe(cref.FR, n, seed)
    out = torch.empty((n, ctx.aff_words), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), to_dev(k), n, out)
    torch.cuda.synchronize()
    return out, to_host(out), k


def test_field_ops(cv):
    ctx, o, c = cv
    for field, p, nw in ((cref.FR, c["r"], 4), (cref.FQ, c["q"], 6)):
        a, b = o.rand_fe(field, 3000, 11), o.rand_fe(field, 3000, 12)
        a[0], b[0] = 0, 0
        a[1] = limbs([p - 1], nw)[0]
        b[1] = a[1]
        a[2] = limbs([1], nw)[0]
        b[3] = limbs([p - 1], nw)[0]
        am, bm = o.to_mont(field, a), o.to_mont(field, b)
        for op in (0, 1, 2, 3):
            assert np.array_equal(ctx.fp_binop(field, op, am, bm), o.binop(field, op, am, bm)), (field, op)
        assert np.array_equal(ctx.fp_binop(field, 4, am[1:65]), o.binop(field, 4, am[1:65])), field
        assert np.array_equal(ctx.fp_binop(field, 5, a), am) and np.array_equal(ctx.fp_binop(field, 6, am), a), field


@pytest.mark.parametrize("log_n", [0, 1, 4, 9, 11, 12, 13, 16, 18, 20])
def test_ntt_all_modes(cv, log_n):
    ctx, o, _ = cv
    x = fr_mont(o, 1 << log_n, 1000 + log_n)
    for inv, cos in MODES:
        assert np.array_equal(ctx.ntt_host(x.copy(), log_n, inv, cos), o.ntt(x, log_n, inv, cos)), (log_n, inv, cos)


def test_ntt_large_and_zero_padded_batch(cv):
    ctx, o, _ = cv
    log_n = 22
    x = fr_mont(o, 1 << log_n, 77)
    d = to_dev(x)
    ctx.ntt_dev(d, log_n, False, True)
    assert np.array_equal(to_host(d), o.ntt(x, log_n, False, True))
    ctx.ntt_dev(d, log_n, True, True)
    assert np.array_equal(to_host(d), x)
    log_n, length = 18, (1 << 16) + 3                               # the quotient round's shape: coset FFT(4n) of n + 3 coefficients
    xs = [fr_mont(o, 1 << log_n, 2100 + k) for k in range(3)]
    ds = [to_dev(v) for v in xs]
    ctx.ntt_batch_dev(ds, log_n, False, True, length=length)
    for v, dv in zip(xs, ds):
        ref_in = np.zeros_like(v)
        ref_in[:length] = v[:length]
        assert np.array_equal(to_host(dv), o.ntt(ref_in, log_n, False, True))


def test_ntt_domain_limit(cv):
    """log_n beyond what the library supports is Error::InvalidEvalDomainSize (prove.rs:77-81), not a crash"""
    import zkt_plonk_b200 as z
    ctx, _, _ = cv
    with pytest.raises(z.ZkbError) as e:
        ctx.ntt_dev(256, 48, False, False)
    assert e.value.code == -2


@pytest.mark.parametrize("n", [1, 2, 33, 1000, 4097])
def test_msm_arbitrary_bases(cv, n):
    ctx, o, c = cv
    _, pts, _ = gpu_points(ctx, o, n, 40 + n)
    sc = o.rand_fe(cref.FR, n, 50 + n)
    r = c["r"]
    if n >= 33:                                                  # the edge cases of the BN254 tests: 0, 1, r - 1, infinity, repeated base
        sc[0] = 0
        sc[1] = limbs([1], 4)[0]
        sc[2] = limbs([r - 1], 4)[0]
        pts[3] = 0
        pts[5] = pts[4]
        sc[6] = limbs([1], 4)[0]
    got, ginf = ctx.msm_bases(pts, sc)
    exp, einf = o.msm_g1(pts, sc)
    assert ginf == einf and np.array_equal(got, exp)
    got, ginf = ctx.msm_bases(pts, np.zeros_like(sc))
    assert ginf and not got.any()


@pytest.mark.parametrize("log_n,tables", [(10, False), (12, True), (16, False), (16, True)])
def test_msm_resident_key(cv, log_n, tables):
    ctx, o, _ = cv
    n = 1 << log_n
    dpts, pts, _ = gpu_points(ctx, o, n, 60 + log_n)
    if log_n == 10:
        assert np.array_equal(pts[:16], o.g1_mul(ctx.g1_generator(), o.rand_fe(cref.FR, n, 60 + log_n)[:16]))
        assert all(o.g1_on_curve(p) for p in pts[:64])
    ctx.srs_load(dpts)
    ctx.srs_precompute(0 if tables else -1)
    try:
        for seed, skew in ((1, False), (2, True)):
            sc = o.rand_fe(cref.FR, n, 70 + seed)
            if skew:                                             # witness-like: zeros, ones, small values
                rng = np.random.default_rng(seed)
                kind = rng.random(n)
                sc[kind < 0.2] = 0
                sc[(kind >= 0.2) & (kind < 0.4)] = limbs([1], 4)[0]
                small = (kind >= 0.4) & (kind < 0.6)
                sc[small] = 0
                sc[small, 0] = rng.integers(0, 1 << 16, size=int(small.sum()), dtype=np.uint64)
            exp, einf = o.msm_g1(pts, sc)
            got, ginf = ctx.msm(to_dev(sc))
            assert ginf == einf and np.array_equal(got, exp), (log_n, tables, skew)
            got, ginf = ctx.msm(sc)                             # host scalars
            assert np.array_equal(got, exp)
            half = n // 2 + 3                                   # a sub-range of the key: commitments to shorter polynomials
            exp, _ = o.msm_g1(pts[5:5 + half], sc[:half])
            got, _ = ctx.msm(to_dev(sc[:half]), offset=5, n=half)
            assert np.array_equal(got, exp)
            if log_n == 16:                                     # the same MSM as point ranges through shared buckets
                ctx.set_msm_parts(3, 5, 10)
                try:
                    assert np.array_equal(ctx.msm(to_dev(sc[:half]), offset=5, n=half)[0], exp)
                    assert np.array_equal(ctx.msm(np.ascontiguousarray(sc[:half]), offset=5)[0], exp)
                finally:
                    ctx.set_msm_parts(1, 4, 19)
    finally:
        ctx.srs_precompute(-1)


def test_msm_2_20_closed_form_and_commitments(cv):
    """2^20 points k_i G: sum s_i (k_i G) == (sum s_i k_i mod r) G, with the fixed-base tables; then kzg10::commit's path
    (Montgomery coefficients, into_repr on the device) single and as a batch, and the XYZZ partial-sum interface."""
    import zkt_plonk_b200 as z
    ctx, o, c = cv
    n, r = 1 << 20, c["r"]
    dpts, pts, k = gpu_points(ctx, o, n, 5)
    ctx.srs_load(dpts)
    ctx.srs_precompute(0)
    try:
        sc = o.rand_fe(cref.FR, n, 6)
        dot = sum(a * b for a, b in zip(ints(sc, 4), ints(k, 4))) % r
        exp = o.g1_mul(ctx.g1_generator(), limbs([dot], 4))[0]
        got, inf = ctx.msm(to_dev(sc))
        assert not inf and np.array_equal(got, exp)
        got, inf = ctx.msm(sc)
        assert np.array_equal(got, exp)
        t = ctx.msm_last_timing()
        assert t["windows"] * t["c"] >= c["r"].bit_length() and t["accumulate_ms"] > 0
        # commitments: Montgomery coefficients in HBM
        m = 1 << 14
        polys = [fr_mont(o, m + j, 90 + j) for j in range(3)]
        exps = [o.msm_g1(pts[: p.shape[0]], o.from_mont(cref.FR, p)) for p in polys]
        got, inf = ctx.commit_dev(to_dev(polys[0]), 0, polys[0].shape[0])
        assert np.array_equal(got, exps[0][0])
        res = ctx.commit_batch_dev([to_dev(p) for p in polys], [p.shape[0] for p in polys])
        for (g, gi), (e, ei) in zip(res, exps):
            assert gi == ei and np.array_equal(g, e)
        part = ctx.msm_partial(to_dev(o.from_mont(cref.FR, polys[1])), 0, polys[1].shape[0])
        assert part.shape == (24,)
        s, _ = z.context.sum_partials(part.reshape(1, 24), curve=ctx.curve)
        assert np.array_equal(s, exps[1][0])
    finally:
        ctx.srs_precompute(-1)


@pytest.mark.parametrize("log_n", [3, 10])
def test_grand_products_and_quotient(cv, log_n):
    """z1 / z2 evaluations and the fused quotient kernel over this curve's Fr (generator 7 / 22, its own roots of unity)"""
    import torch
    ctx, o, _ = cv
    n = 1 << log_n
    cols = [fr_mont(o, n, 300 + 10 * log_n + j) for j in range(6)]
    beta, gamma, delta, eps = (fr_mont(o, 1, 400 + j)[0].copy() for j in range(4))
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    ctx.z1_evals_dev(log_n, beta, gamma, *[to_dev(v) for v in cols], out)
    assert not ctx.grand_product_failed()
    assert np.array_equal(to_host(out), o.z1_evals(log_n, beta, gamma, *cols))
    f, t, h1, h2 = cols[:4]
    ctx.z2_evals_dev(log_n, delta, eps, to_dev(f), to_dev(t), to_dev(h1), to_dev(h2), out)
    assert np.array_equal(to_host(out), o.z2_evals(log_n, delta, eps, f, t, h1, h2))
    n4 = 4 * n
    wit = {k: fr_mont(o, n4, 500 + i) for i, k in enumerate(o.WIT_ORDER)}
    epk = {k: fr_mont(o, n4, 600 + i) for i, k in enumerate(o.EPK_ORDER)}
    epk["x"], epk["zh"], epk["l1"] = o.epk_free_tables(log_n)
    ch = fr_mont(o, 5, 700 + log_n)
    exp = o.quotient_evals(log_n, ch, wit, epk)
    l1 = torch.empty((n4, 4), dtype=torch.int64, device="cuda")
    ctx.l1_coset_dev(log_n, l1)
    assert np.array_equal(to_host(l1), epk["l1"])
    from zkt_plonk_b200.prover_ops import EPK_ORDER, WIT_ORDER
    dw = [to_dev(wit[k]) for k in WIT_ORDER]
    de = [to_dev(epk[k]) if k != "l1" else l1 for k in EPK_ORDER]
    q = torch.empty((n4, 4), dtype=torch.int64, device="cuda")
    ctx.quotient_evals_dev(log_n, ch, dw, de, q)
    assert np.array_equal(to_host(q), exp)


def test_polynomial_utilities(cv):
    """evaluate / linear combination / division by (X - z): linearization_poly.rs:55-111 and kzg10::open's witness polynomial"""
    import torch
    ctx, o, _ = cv
    n = 5000
    p = fr_mont(o, n, 31)
    z = fr_mont(o, 1, 32)[0].copy()
    assert np.array_equal(ctx.poly_eval_dev(to_dev(p), n, z), o.poly_eval(p, z))
    polys = [fr_mont(o, n - 7 * j, 33 + j) for j in range(4)]
    sc = fr_mont(o, 4, 40)
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    ctx.poly_lincomb_dev([to_dev(v) for v in polys], [v.shape[0] for v in polys], sc, out, n)
    assert np.array_equal(to_host(out), o.poly_lincomb(polys, sc, n))
    quot = torch.empty((n - 1, 4), dtype=torch.int64, device="cuda")
    ev = ctx.poly_divide_linear_dev(to_dev(p), n, z, quot)
    eq, eev = o.poly_divide_linear(p, z)
    assert np.array_equal(ev, eev) and np.array_equal(to_host(quot), eq)


def test_key_files_drive_the_native_prover(cv, tmp_path):
    """keys/mod.rs:29-40,180-203 on Bls12_381 / Bls12_377: the pk / vk files this curve's build writes equal the independent
    Python serialisation of the oracle-backend keys (96-byte commitments); a key loaded back from them, with the committer key
    read from a ck file, proves byte-identically to the key set up from the circuit's columns."""
    import torch
    import zkt_plonk_b200 as z
    from oracle import arkser, plonk_ref, pyref
    from zkt_plonk_b200 import field, prover, synthetic
    ctx, o, c = cv
    field.use_curve(ctx.curve)
    pyref.use_curve(ctx.curve)
    arkser.use_curve(ctx.curve)
    try:
        P = field.R_MOD
        tau = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P
        circ = synthetic.make_circuit(6, seed=33, table_size=16)
        n_powers = 4 * circ.n + 1
        powers, x = [], 1
        for _ in range(n_powers):
            powers.append(x)
            x = x * tau % P
        d_srs = torch.empty((n_powers, ctx.aff_words), dtype=torch.int64, device="cuda")
        ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), to_dev(limbs(powers, 4)), n_powers, d_srs)
        torch.cuda.synchronize()
        h_srs = to_host(d_srs)
        kzg = z.GpuKZG10(ctx)
        kzg.load_committer_key(d_srs)
        blinders = list(range(900, 919))
        native = prover.NativeProver(ctx, circ)
        raw = native.prove_bytes(blinders)
        pk_path, vk_path, ck_path = tmp_path / "pk", tmp_path / "vk", tmp_path / "ck"
        native.save_keys(pk_path, vk_path)
        native.close()
        obe = plonk_ref.OracleBackend(h_srs)
        opk, ovk = prover.setup(obe, circ)
        polys = {name: prover.mont_array_to_ints(opk.polys[name].data[: opk.polys[name].len]) for name in arkser.PK_ORDER}
        assert pk_path.read_bytes() == arkser.prover_key(polys)
        assert vk_path.read_bytes() == arkser.verifier_key(ovk.n, ovk.pi_roots, ovk.commits)
        pts = [prover.point_to_ints(row, not row.any()) for row in h_srs]
        ck_path.write_bytes(arkser.committer_key(pts, pts[:2], n_powers - 1))
        ctx.srs_load(d_srs[:8].contiguous())                               # forget the key, then read it from the file
        kzg.load_committer_key_file(ck_path)
        assert ctx.srs_size() == n_powers
        loaded = prover.NativeProver(ctx, circ, key_files=(pk_path, vk_path))
        try:
            assert loaded.vk().commits == ovk.commits and loaded.vk().pi_roots == ovk.pi_roots
            assert loaded.prove_bytes(blinders) == raw
            assert raw == prover.prove(obe, opk, ovk, circ, blinders).to_bytes()
        finally:
            loaded.close()
    finally:
        field.use_curve("bn254")
        pyref.use_curve("bn254")
        arkser.use_curve("bn254")


# ------------------------------------------------------------------------------------------------ whole proofs on the BLS12 curves
@pytest.mark.parametrize("log_n,fixed_base", [(5, False), (10, True), (14, True)])
def test_gpu_proof_is_byte_identical_and_verifies(cv, log_n, fixed_base):
    """plonk.rs:226-254 (test_full on Bls12_381 / Bls12_377): the five-round schedule (zkt_plonk_b200.prover) over this curve's
    build of the library against the same schedule over the per-curve CPU oracle -- same circuit, witness, SRS and blinders:
    identical verifier keys, byte-identical 1010-byte proofs, accepted by the restated verifier (PC::check through the synthetic
    SRS's trapdoor).  At 2^14 the CPU prover is skipped: the GPU proof is verified only."""
    import random
    import torch
    import zkt_plonk_b200 as z
    from oracle import plonk_ref, pyref
    from zkt_plonk_b200 import field, prover, synthetic
    ctx, o, c = cv
    field.use_curve(ctx.curve)
    pyref.use_curve(ctx.curve)
    try:
        P = field.R_MOD
        tau = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P
        circ = synthetic.make_circuit(log_n, seed=40 + log_n, table_size=min(64, (1 << log_n) // 4))
        assert synthetic.check_gates(circ)
        powers, x = [], 1
        for _ in range(circ.n + 8):
            powers.append(x)
            x = x * tau % P
        d_srs = torch.empty((circ.n + 8, ctx.aff_words), dtype=torch.int64, device="cuda")
        ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), to_dev(limbs(powers, 4)), circ.n + 8, d_srs)
        torch.cuda.synchronize()
        kzg = z.GpuKZG10(ctx)
        kzg.load_committer_key(d_srs)
        if fixed_base:
            ctx.srs_precompute(0)
        gbe = prover.GpuBackend(kzg)
        rnd = random.Random(99)
        blinders = [rnd.randrange(P) for _ in range(19)]
        gpk, gvk = prover.setup(gbe, circ)
        gproof = prover.prove(gbe, gpk, gvk, circ, blinders)
        raw = gproof.to_bytes()
        assert len(raw) == 1010
        pub = list(circ.pi.values())
        assert plonk_ref.verify(gvk, gproof, pub, tau) == 0
        assert plonk_ref.verify(gvk, prover.proof_from_bytes(raw), pub, tau) == 0
        from zkt_plonk_b200 import verifier                            # the library's own verifier (pairings on this curve)
        assert verifier.verify(gvk, raw, pub, verifier.make_cvk(tau)) == 0
        assert verifier.verify(gvk, raw, pub, verifier.make_cvk(tau + 1)) == 1
        assert plonk_ref.verify(gvk, gproof, [(pub[0] + 1) % P] + pub[1:], tau) != 0
        # the C++ round driver (zkb_plonk_setup / zkb_plonk_prove) of this curve's build: the same bytes, also from the variables
        npv = prover.NativeProver(ctx, circ)
        try:
            assert npv.vk().commits == gvk.commits
            assert npv.prove_bytes(blinders) == raw
            if circ.wiring is not None:
                npv.set_wiring()
                assert npv.prove_bytes(blinders, from_vars=True) == raw
        finally:
            npv.close()
        if log_n <= 10:
            obe = plonk_ref.OracleBackend(to_host(d_srs))
            opk, ovk = prover.setup(obe, circ)
            assert gvk.commits == ovk.commits and gvk.pi_roots == ovk.pi_roots
            assert prover.prove(obe, opk, ovk, circ, blinders).to_bytes() == raw
    finally:
        ctx.srs_precompute(-1)
        field.use_curve("bn254")
        pyref.use_curve("bn254")


@pytest.mark.parametrize("world,tables", [(2, False), (3, True)])
def test_range_commitments_add_up(cv, world, tables):
    """SURVEY.md 8e on these curves: contexts that each hold one point range of the committer key (zkb_srs_set_range) commit to the
    overlap of every polynomial with their range; the partial commitments (XYZZ sums of 24 words) add up to the commitment of
    the unsharded key and to the oracle's MSM."""
    import torch
    import zkt_plonk_b200 as z
    from zkt_plonk_b200.context import sum_partials
    from zkt_plonk_b200.parallel import shard_bounds
    ctx, o, _ = cv
    n = 4099
    d_pts, h_pts, _ = gpu_points(ctx, o, n, 11)
    polys = [fr_mont(o, n, 21), fr_mont(o, n - 37, 22), fr_mont(o, 5, 23)]
    polys[1][:100] = 0
    lens = [p.shape[0] for p in polys]
    devs = [to_dev(p) for p in polys]
    ctx.srs_load(d_pts)
    want = ctx.commit_batch_dev(devs, lens)
    b = shard_bounds(n, world)
    parts = []
    for r in range(world):
        c = z.Context(0, curve=ctx.curve)
        c.set_stream(torch.cuda.current_stream())
        c.srs_load(d_pts[b[r]:b[r + 1]].contiguous())
        c.srs_set_range(b[r], n)
        if tables:
            c.srs_precompute(0)
        for k in range(len(polys)):
            c.commit_push(devs[k], lens[k])
        parts.append(c.commit_finish_partials(len(polys)))
        c.close()
    for k in range(len(polys)):
        got, inf = sum_partials(np.stack([parts[r][k] for r in range(world)]), curve=ctx.curve)
        assert inf == want[k][1] and np.array_equal(got, want[k][0])
        exp, einf = o.msm_g1(h_pts[:lens[k]], o.from_mont(cref.FR, polys[k]))
        assert inf == einf and np.array_equal(got, exp)
