"""The inner-product-argument commitment scheme (the reference's second PC: plonk-core/src/commitment.rs:49-86, run by
plonk-core/src/test.rs:73,84).

CPU: the restated protocol (oracle/ipa_ref.py) against its defining properties -- an opening checks, anything tampered does
not, the folded key has its closed form.  GPU (`-m gpu`): the library's round entry points (csrc/ipa.cu, through the C ABI)
against big-integer definitions bit for bit, and whole openings by zkt_plonk_b200.ipa.GpuIPA equal to the restated ones on
every curve the reference instantiates.
"""
import random

import numpy as np
import pytest

from oracle import ipa_ref, pyref

CURVES = ["bn254", "bls12_381", "bls12_377"]


@pytest.fixture
def on_curve(request):
    from zkt_plonk_b200 import field
    field.use_curve(request.param)
    pyref.use_curve(request.param)
    yield request.param
    field.use_curve("bn254")
    pyref.use_curve("bn254")


def host_key(n, seed):
    rnd = random.Random(seed)
    ks = [rnd.randrange(1, pyref.R_MOD) for _ in range(n + 1)]
    pts = [pyref.g1_mul(k, pyref.G1_GEN) for k in ks]
    return pts[:n], pts[n]


# ------------------------------------------------------------------------------------------------ CPU: the restated protocol
@pytest.mark.parametrize("on_curve", ["bn254", "bls12_381"], indirect=True)
def test_restated_protocol_properties(on_curve):
    r = pyref.R_MOD
    rnd = random.Random(3)
    n = 8
    key, h = host_key(n, 11)
    coeffs = [rnd.randrange(r) for _ in range(6)]                  # shorter than the key: padded with zeros
    point = rnd.randrange(r)
    value = sum(c * pow(point, i, r) for i, c in enumerate(coeffs)) % r
    C = ipa_ref.commit(key, coeffs)
    assert C == pyref.msm_naive(key[:6], coeffs)
    proof, challenges = ipa_ref.open_(key, h, coeffs, C, point)
    l_vec, r_vec, final_key, c = proof
    assert len(l_vec) == len(r_vec) == 3 and len(set(challenges)) == 3
    # closed forms of what the folding leaves: G_final = <h-coefficients, G>, c = <coeffs, h-coefficients with inverted challenges>
    hc = ipa_ref.check_poly_coeffs(challenges)
    assert final_key == pyref.msm_naive(key, hc)
    assert ipa_ref.check_poly_eval(challenges, point) == sum(v * pow(point, i, r) for i, v in enumerate(hc)) % r
    inv_hc = ipa_ref.check_poly_coeffs([pow(x, -1, r) for x in challenges])
    assert c == sum(a * b for a, b in zip(coeffs + [0, 0], inv_hc)) % r
    assert ipa_ref.check(key, h, C, point, value, proof)
    # soundness smoke: every part of the statement and of the proof matters
    assert not ipa_ref.check(key, h, C, point, (value + 1) % r, proof)
    assert not ipa_ref.check(key, h, C, (point + 1) % r, value, proof)
    assert not ipa_ref.check(key, h, pyref.g1_add(C, key[0]), point, value, proof)
    assert not ipa_ref.check(key, h, C, point, value, (l_vec, r_vec, final_key, (c + 1) % r))
    assert not ipa_ref.check(key, h, C, point, value, ([l_vec[1], l_vec[0], l_vec[2]], r_vec, final_key, c))
    assert not ipa_ref.check(key, h, C, point, value, (l_vec, r_vec, pyref.g1_add(final_key, key[1]), c))
    assert not ipa_ref.check(key, h, C, point, value, (l_vec[:2], r_vec[:2], final_key, c))
    # the hash is injectable (the byte encodings in front of it are the one recalled part): any oracle gives a sound protocol
    ctr = [0]

    def other(data):
        ctr[0] += 1
        return (int.from_bytes(data[:16], "little") * 2654435761 + ctr[0]) % r or 1

    C2 = ipa_ref.commit(key, coeffs)
    ctr[0] = 0
    proof2, ch2 = ipa_ref.open_(key, h, coeffs, C2, point, oracle=other)
    ctr[0] = 0
    assert ipa_ref.check(key, h, C2, point, value, proof2, oracle=other) and ch2 != challenges
    # several polynomials at one point (how the prover calls PC::open): one opening of sum_j xi^(2 j) p_j
    polys = [coeffs, [rnd.randrange(r) for _ in range(8)], [rnd.randrange(r) for _ in range(3)]]
    comms = [ipa_ref.commit(key, p) for p in polys]
    vals = [sum(c * pow(point, i, r) for i, c in enumerate(p)) % r for p in polys]
    xi = rnd.randrange(r)
    cc, CC, vv = ipa_ref.combine(polys, comms, vals, xi)
    assert CC == ipa_ref.commit(key, cc) and vv == sum(c * pow(point, i, r) for i, c in enumerate(cc)) % r
    assert cc[7] == (pow(xi, 2, r) * polys[1][7]) % r
    proof_c, _ = ipa_ref.open_(key, h, cc, CC, point)
    assert ipa_ref.check(key, h, CC, point, vv, proof_c)
    assert not ipa_ref.check(key, h, ipa_ref.combine(polys, comms, vals, xi + 1)[1], point, vv, proof_c)
    # the zero polynomial: identity commitment, identity cross terms
    proof0, _ = ipa_ref.open_(key, h, [], None, point)
    assert proof0[3] == 0 and all(p is None for p in proof0[0]) and ipa_ref.check(key, h, None, point, 0, proof0)


def test_challenge_derivation():
    """compute_random_oracle_challenge: masked Blake2s digests, retried with the next counter until below the modulus."""
    import hashlib
    r = pyref.R_MOD
    seen_retry = False
    for k in range(64):
        data = bytes([k]) * 7
        v = ipa_ref.random_oracle_challenge(data)
        assert 0 <= v < r
        first = int.from_bytes(hashlib.blake2s(data + (0).to_bytes(8, "little")).digest(), "little") & ((1 << r.bit_length()) - 1)
        if first >= r:
            seen_retry = True
            assert v != first
        else:
            assert v == first
    assert seen_retry                                              # BN254's r is 0.76 of 2^254: a quarter of the digests retry
    from zkt_plonk_b200 import ipa
    assert all(ipa.random_oracle_challenge(bytes([k]) * 7) == ipa_ref.random_oracle_challenge(bytes([k]) * 7) for k in range(16))
    assert ipa.g1_bytes(None) == ipa_ref.g1_bytes(None) and ipa.g1_bytes((5, 9)) == ipa_ref.g1_bytes((5, 9))


@pytest.mark.parametrize("curve", CURVES)
def test_glv_split_of_the_key_fold(curve):
    """The key fold multiplies by a challenge through the curve's endomorphism: the library's split k = k1 + lambda k2 (host
    multi-word arithmetic, csrc/ipa.cu glv_split) against the generator's big-integer version (tools/gen_curves.py glv_params):
    same halves, short, and lambda acts on G1 as (x, y) -> (beta x, y)."""
    import ctypes
    import importlib.util
    import os
    from zkt_plonk_b200 import _lib
    spec = importlib.util.spec_from_file_location("gen_curves", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_curves.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    c = [x for x in gen.CURVES if x["name"] == curve][0]
    gen.check(c)
    g = gen.glv_params(c, trials=50)
    r, lam = c["r"], g["lam"]
    assert gen.ec_mul(lam, c["gen"], c["q"]) == (g["beta"] * c["gen"][0] % c["q"], c["gen"][1])
    lib = _lib.lib(curve)
    rnd = random.Random(8)
    for i in range(400):
        k = [0, 1, r - 1, lam, r - lam, (r - 1) // 2, 2, (1 << 128) % r][i] if i < 8 else rnd.randrange(r)
        kk = np.array([(k >> (64 * j)) & (2**64 - 1) for j in range(4)], dtype=np.uint64)
        m1, m2 = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        n1, n2 = ctypes.c_int(0), ctypes.c_int(0)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        assert lib.zkb_test_glv_split(vp(kk), vp(m1), ctypes.byref(n1), vp(m2), ctypes.byref(n2)) == 1
        a = sum(int(v) << (64 * j) for j, v in enumerate(m1)) * (-1 if n1.value else 1)
        b = sum(int(v) << (64 * j) for j, v in enumerate(m2)) * (-1 if n2.value else 1)
        assert (a + lam * b) % r == k and abs(a) < 1 << 130 and abs(b) < 1 << 130
        k2 = (((k * g["g1"]) >> 384) * g["m1"] + ((k * g["g2"]) >> 384) * g["m2"]) % r
        k1 = (k - lam * k2) % r
        assert a % r == k1 and b % r == k2
    bad = np.array([2**64 - 1] * 4, dtype=np.uint64)                 # not below r
    assert lib.zkb_test_glv_split(vp(bad), vp(m1), ctypes.byref(n1), vp(m2), ctypes.byref(n2)) == _lib.ZKB_ERR_INVALID


# ------------------------------------------------------------------------------------------------ GPU
def _ctx(curve):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import zkt_plonk_b200 as z
    c = z.Context(0, curve=curve)
    c.set_stream(torch.cuda.current_stream())
    return c


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to("cuda")


def _fr_arr(vals, mont=True):
    from zkt_plonk_b200 import field
    return np.array([field.int_to_limbs(field.to_mont(v) if mont else v) for v in vals], dtype=np.uint64).reshape(-1, 4)


def _fr_ints(arr):
    from zkt_plonk_b200 import field
    a = np.ascontiguousarray(arr).view(np.uint64).reshape(-1, 4)
    return [field.from_mont(field.limbs_to_int(row)) for row in a]


def _device_key(ctx, n, seed):
    """k_i * G built in HBM; returns (device tensor of n + 1 points, the points as canonical ints)"""
    import torch
    from zkt_plonk_b200 import field
    rnd = random.Random(seed)
    ks = [rnd.randrange(1, field.R_MOD) for _ in range(n + 1)]
    out = torch.empty((n + 1, ctx.aff_words), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), _dev(_fr_arr(ks, mont=False)), n + 1, out)
    torch.cuda.synchronize()
    w = field.FQ_WORDS
    host = out.cpu().numpy().view(np.uint64)
    pts = [(field.from_mont(field.limbs_to_int(row[:w]), field.Q_MOD), field.from_mont(field.limbs_to_int(row[w:]), field.Q_MOD)) for row in host]
    return out, pts, ks


@pytest.mark.gpu
@pytest.mark.parametrize("on_curve", CURVES, indirect=True)
def test_gpu_rounds_match_the_definitions(on_curve):
    """zkb_ipa_round_lr_dev / zkb_ipa_round_fold_dev against Python integers: cross terms, inner products and all three folds,
    over two consecutive rounds (the second one runs on the folded vectors), n = 2 included."""
    import torch
    from zkt_plonk_b200 import field
    from zkt_plonk_b200.ipa import GpuIPA
    ctx = _ctx(on_curve)
    try:
        r = field.R_MOD
        rnd = random.Random(17)
        helper = GpuIPA(ctx)
        import os
        plain, binary = {"ZKB_IPA_GLV": "0"}, {"ZKB_IPA_NAF": "0"}       # A/B knobs of the key fold, read per call: same results
        for n, knobs in ((2, {}), (16, {}), (64, {}), (16, plain), (16, binary), (8, {**plain, **binary})):
            for name in ("ZKB_IPA_GLV", "ZKB_IPA_NAF"):
                os.environ.pop(name, None)
            os.environ.update(knobs)
            d_key, pts, ks = _device_key(ctx, n, 100 + n)
            assert pts[0] == pyref.g1_mul(ks[0], pyref.G1_GEN)
            key = list(pts[:n])
            c = [rnd.randrange(r) for _ in range(n)]
            z = [rnd.randrange(r) for _ in range(n)]
            c[0], z[1 % n] = 0, r - 1
            d_c, d_z, d_k = _dev(_fr_arr(c)), _dev(_fr_arr(z)), d_key[:n].clone()
            m = n
            while m > 1:
                half = m // 2
                (l_xy, l_inf), (r_xy, r_inf), ip_l, ip_r = ctx.ipa_round_lr_dev(d_c, d_z, d_k, m)
                exp_l, exp_r = pyref.msm_naive(key[:half], c[half:m]), pyref.msm_naive(key[half:m], c[:half])
                assert helper._pt_ints(l_xy, l_inf) == exp_l and helper._pt_ints(r_xy, r_inf) == exp_r
                ipl, ipr = ipa_ref.inner(c[half:m], z[:half]), ipa_ref.inner(c[:half], z[half:m])
                assert _fr_ints(ip_l) == [ipl] and _fr_ints(ip_r) == [ipr]
                # with h': L = <c_r, G_l> + <c_r, z_l> h', R = <c_l, G_r> + <c_l, z_r> h' (h' = the key's spare point)
                (l_xy, l_inf), (r_xy, r_inf), ip_l2, ip_r2 = ctx.ipa_round_lr_dev(d_c, d_z, d_k, m, helper._pt_array([pts[n]])[0])
                assert helper._pt_ints(l_xy, l_inf) == pyref.g1_add(exp_l, pyref.g1_mul(ipl, pts[n]))
                assert helper._pt_ints(r_xy, r_inf) == pyref.g1_add(exp_r, pyref.g1_mul(ipr, pts[n]))
                assert np.array_equal(ip_l2, ip_l) and np.array_equal(ip_r2, ip_r)
                x = rnd.randrange(1, r) if m != 16 else 1            # x = 1: the fold is a plain addition
                xi = pow(x, -1, r)
                ctx.ipa_round_fold_dev(d_c, d_z, d_k, m, _fr_arr([x])[0], _fr_arr([xi])[0])
                c = [(a + xi * b) % r for a, b in zip(c[:half], c[half:m])]
                z = [(a + x * b) % r for a, b in zip(z[:half], z[half:m])]
                key = [pyref.g1_add(a, pyref.g1_mul(x, b)) for a, b in zip(key[:half], key[half:m])]
                torch.cuda.synchronize()
                assert _fr_ints(d_c[:half].cpu().numpy()) == c and _fr_ints(d_z[:half].cpu().numpy()) == z
                got = d_k[:half].cpu().numpy().view(np.uint64)
                assert [helper._pt_ints(row, not row.any()) for row in got] == key
                m = half
        for name in ("ZKB_IPA_GLV", "ZKB_IPA_NAF"):
            os.environ.pop(name, None)
        # the fold's special cases: G_l = -x G_r gives the identity (zeros), G_l = x G_r doubles
        d_key, pts, ks = _device_key(ctx, 3, 5)
        x = 12345
        a = pyref.g1_mul(x, pts[2])
        arr = helper._pt_array([pyref.g1_neg(a), a, pts[2], pts[2]])
        d_k = _dev(arr)
        d_c, d_z = _dev(_fr_arr([1, 2, 3, 4])), _dev(_fr_arr([5, 6, 7, 8]))
        ctx.ipa_round_fold_dev(d_c, d_z, d_k, 4, _fr_arr([x])[0], _fr_arr([pow(x, -1, r)])[0])
        torch.cuda.synchronize()
        got = d_k[:2].cpu().numpy().view(np.uint64)
        assert not got[0].any() and helper._pt_ints(got[1], False) == pyref.g1_add(a, a)
        d_k = _dev(helper._pt_array([pts[0], pts[1], None, None]))       # identity entries in G_r: G_l stays
        ctx.ipa_round_fold_dev(d_c, d_z, d_k, 4, _fr_arr([x])[0], _fr_arr([pow(x, -1, r)])[0])
        torch.cuda.synchronize()
        got = d_k[:2].cpu().numpy().view(np.uint64)
        assert [helper._pt_ints(row, False) for row in got] == [pts[0], pts[1]]
        # argument checks: not a power of two, x * x_inv != 1
        from zkt_plonk_b200._lib import ZkbError
        with pytest.raises(ZkbError):
            ctx.ipa_round_lr_dev(d_c, d_z, d_k, 3)
        with pytest.raises(ZkbError):
            ctx.ipa_round_fold_dev(d_c, d_z, d_k, 4, _fr_arr([5])[0], _fr_arr([6])[0])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("on_curve", CURVES, indirect=True)
def test_gpu_opening_equals_the_restated_one(on_curve):
    """GpuIPA.open over the device rounds gives the proof the restated CPU protocol gives (same L, R, final key, c), its check
    accepts it and rejects tampered statements; a 2^12 opening round-trips on the GPU alone."""
    import torch
    from zkt_plonk_b200 import field
    from zkt_plonk_b200.ipa import GpuIPA, IpaProof
    ctx = _ctx(on_curve)
    try:
        r = field.R_MOD
        rnd = random.Random(23)
        for n, length in ((16, 16), (32, 21)):
            d_key, pts, _ = _device_key(ctx, n, 200 + n)
            pc = GpuIPA(ctx)
            pc.load_committer_key(d_key[:n].contiguous(), pts[n])
            coeffs = [rnd.randrange(r) for _ in range(length)]
            d_coeffs = _dev(_fr_arr(coeffs))
            point = rnd.randrange(r)
            C = pc.commit_dev(d_coeffs, length)
            assert C == ipa_ref.commit(pts[:n], coeffs)
            proof, value = pc.open(d_coeffs, length, C, point)
            assert value == sum(c * pow(point, i, r) for i, c in enumerate(coeffs)) % r
            ref, challenges = ipa_ref.open_(pts[:n], pts[n], coeffs, C, point)
            assert (proof.l_vec, proof.r_vec, proof.final_comm_key, proof.c) == ref
            # the verifier's linear-time step on the device: <h-coefficients, G> for the check polynomial of the challenges
            fk = pc._pt_ints(*ctx.ipa_final_key_dev(pc.key, n, _fr_arr(challenges)))
            assert fk == pyref.msm_naive(pts[:n], ipa_ref.check_poly_coeffs(challenges)) == proof.final_comm_key
            assert pc.check(C, point, value, proof) and ipa_ref.check(pts[:n], pts[n], C, point, value, ref)
            assert not pc.check(C, point, (value + 1) % r, proof)
            assert not pc.check(C, (point + 1) % r, value, proof)
            assert not pc.check(C, point, value, IpaProof(proof.l_vec, proof.r_vec, proof.final_comm_key, (proof.c + 1) % r))
            assert not pc.check(C, point, value, IpaProof(proof.r_vec, proof.l_vec, proof.final_comm_key, proof.c))
            assert not pc.check(C, point, value, IpaProof(proof.l_vec, proof.r_vec, pts[0], proof.c))
            assert _fr_ints(d_coeffs.cpu().numpy()) == coeffs          # the caller's polynomial is untouched
        # several polynomials at one point, as the prover calls PC::open: equal to the restated combination's opening
        n = 16
        d_key, pts, _ = _device_key(ctx, n, 300)
        pc = GpuIPA(ctx)
        pc.load_committer_key(d_key[:n].contiguous(), pts[n])
        polys = [[rnd.randrange(r) for _ in range(m)] for m in (16, 9, 1)]
        d_polys = [_dev(_fr_arr(p)) for p in polys]
        comms = [pc.commit_dev(d, len(p)) for d, p in zip(d_polys, polys)]
        point, xi = rnd.randrange(r), rnd.randrange(r)
        proof, values = pc.open_many(d_polys, [len(p) for p in polys], comms, point, xi)
        assert values == [sum(c * pow(point, i, r) for i, c in enumerate(p)) % r for p in polys]
        cc, CC, vv = ipa_ref.combine(polys, comms, values, xi)
        ref, _ = ipa_ref.open_(pts[:n], pts[n], cc, CC, point)
        assert (proof.l_vec, proof.r_vec, proof.final_comm_key, proof.c) == ref
        assert pc.check_many(comms, point, values, proof, xi)
        assert not pc.check_many(comms, point, [values[0], values[2], values[1]], proof, xi)
        assert not pc.check_many(comms, point, values, proof, (xi + 1) % r)
        n = 1 << 12
        d_key, pts, _ = _device_key(ctx, n, 77)
        pc = GpuIPA(ctx)
        pc.load_committer_key(d_key[:n].contiguous(), pts[n])
        big = np.random.default_rng(5).integers(0, 1 << 62, size=(n, 4), dtype=np.int64)
        big[:, 3] &= (1 << 58) - 1                                     # below every curve's r: valid Montgomery forms
        d_coeffs = torch.from_numpy(big).to("cuda")
        point = rnd.randrange(r)
        C = pc.commit_dev(d_coeffs, n)
        proof, value = pc.open(d_coeffs, n, C, point)
        assert len(proof.l_vec) == 12 and pc.check(C, point, value, proof)
        assert not pc.check(C, point, (value + 1) % r, proof)
    finally:
        ctx.close()
