"""CPU-only: pins the C oracle against the golden vectors and the independent Python restatement."""
import os
import random

import numpy as np
import pytest

from oracle import cref, pyref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "vectors.npz"))


def mont(field, a):
    return cref.to_mont(field, np.ascontiguousarray(a))


def canon(field, a):
    return cref.from_mont(field, np.ascontiguousarray(a))


def test_constants():
    assert pyref.TWO_ADIC_ROOT == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert pow(pyref.TWO_ADIC_ROOT, 1 << 28, pyref.R_MOD) == 1
    assert pow(pyref.TWO_ADIC_ROOT, 1 << 27, pyref.R_MOD) == pyref.R_MOD - 1
    one = cref.limbs_to_ints(mont(cref.FR, cref.ints_to_limbs([1])))[0]
    assert one == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb


def test_field_ops_vs_python():
    for field, p in ((cref.FR, pyref.R_MOD), (cref.FQ, pyref.Q_MOD)):
        a, b = cref.rand_fe(field, 200, 1), cref.rand_fe(field, 200, 2)
        a[0], b[0] = 0, 0
        a[1] = cref.ints_to_limbs([p - 1])[0]
        b[1] = a[1]
        ai, bi = cref.limbs_to_ints(a), cref.limbs_to_ints(b)
        am, bm = mont(field, a), mont(field, b)
        assert cref.limbs_to_ints(canon(field, cref.binop(field, 0, am, bm))) == [x * y % p for x, y in zip(ai, bi)]
        assert cref.limbs_to_ints(canon(field, cref.binop(field, 1, am, bm))) == [(x + y) % p for x, y in zip(ai, bi)]
        assert cref.limbs_to_ints(canon(field, cref.binop(field, 2, am, bm))) == [(x - y) % p for x, y in zip(ai, bi)]
        assert cref.limbs_to_ints(canon(field, cref.binop(field, 4, am[1:]))) == [pow(x, -1, p) for x in ai[1:]]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_ntt_golden(log_n):
    x = mont(cref.FR, GOLD[f"ntt_in_{log_n}"])
    for key, inv, cos in (("fwd", 0, 0), ("inv", 1, 0), ("cfwd", 0, 1), ("cinv", 1, 1)):
        got = canon(cref.FR, cref.ntt(x, log_n, inv, cos))
        assert np.array_equal(got, GOLD[f"ntt_{key}_{log_n}"]), (log_n, key)


def test_ntt_zero_padding_golden():
    x = np.zeros((16, 4), dtype=np.uint64)
    x[:5] = mont(cref.FR, GOLD["ntt_short_in"])
    assert np.array_equal(canon(cref.FR, cref.ntt(x, 4)), GOLD["ntt_short_fwd_4"])
    assert np.array_equal(canon(cref.FR, cref.ntt(x, 4, coset=True)), GOLD["ntt_short_cfwd_4"])


def test_ntt_vs_python_fast_and_roundtrip():
    log_n = 10
    x = cref.rand_fe(cref.FR, 1 << log_n, 7)
    xm = mont(cref.FR, x)
    assert cref.limbs_to_ints(canon(cref.FR, cref.ntt(xm, log_n))) == pyref.ntt(cref.limbs_to_ints(x), log_n)
    for cos in (False, True):
        back = cref.ntt(cref.ntt(xm, log_n, False, cos), log_n, True, cos)
        assert np.array_equal(back, xm)
    # threads do not change results
    assert np.array_equal(cref.ntt(xm, log_n, threads=1), cref.ntt(xm, log_n, threads=4))


def test_g1_known_answers():
    G = mont(cref.FQ, cref.ints_to_limbs([1, 2])).reshape(8)
    pts = cref.g1_mul(G, cref.ints_to_limbs([2, 3, pyref.R_MOD, pyref.R_MOD + 1]))
    assert np.array_equal(canon(cref.FQ, pts[0].reshape(2, 4)), GOLD["g1_2g"])
    assert np.array_equal(canon(cref.FQ, pts[1].reshape(2, 4)), GOLD["g1_3g"])
    assert not pts[2].any()                                  # r * G = infinity, encoded (0, 0)
    assert np.array_equal(pts[3], G)
    assert cref.g1_on_curve(pts[0]) and cref.g1_on_curve(pts[1])


def test_msm_golden():
    P = mont(cref.FQ, GOLD["msm_points"]).reshape(-1, 8)
    out, inf = cref.msm_g1(P, np.ascontiguousarray(GOLD["msm_scalars"]))
    assert not inf
    assert np.array_equal(canon(cref.FQ, out.reshape(2, 4)), GOLD["msm_result"])
    Pc = mont(cref.FQ, GOLD["msm_cancel_points"]).reshape(-1, 8)
    out, inf = cref.msm_g1(Pc, np.ascontiguousarray(GOLD["msm_cancel_scalars"]))
    assert inf and not out.any()


@pytest.mark.parametrize("n", [1, 31, 32, 257])
def test_msm_vs_definition(n):
    """Pippenger (arkworks window rule, incl. the n < 32 -> c = 3 branch) against sum of double-and-add."""
    G = mont(cref.FQ, cref.ints_to_limbs([1, 2])).reshape(8)
    P = cref.g1_mul(G, cref.rand_fe(cref.FR, n, 100 + n))
    s = cref.rand_fe(cref.FR, n, 200 + n)
    terms = np.stack([cref.g1_mul(P[i], s[i:i + 1])[0] for i in range(n)])
    expect = cref.g1_sum(terms)
    out, inf = cref.msm_g1(P, s)
    assert not inf and np.array_equal(out, expect)
    assert np.array_equal(cref.msm_g1(P, s, threads=1)[0], out)


def _rand_ints(rnd, n):
    return [rnd.randrange(pyref.R_MOD) for _ in range(n)]


def test_grand_products_vs_python():
    rnd = random.Random(5)
    log_n, n = 4, 16
    beta, gamma, delta, eps = _rand_ints(rnd, 4)
    cols = [_rand_ints(rnd, n) for _ in range(6)]
    m = lambda v: mont(cref.FR, cref.ints_to_limbs(v))
    got = cref.z1_evals(log_n, m([beta]), m([gamma]), *[m(c) for c in cols])
    assert cref.limbs_to_ints(canon(cref.FR, got)) == pyref.z1_evals(log_n, beta, gamma, *cols)
    f, t, h1, h2 = (_rand_ints(rnd, n) for _ in range(4))
    got = cref.z2_evals(log_n, m([delta]), m([eps]), m(f), m(t), m(h1), m(h2))
    assert cref.limbs_to_ints(canon(cref.FR, got)) == pyref.z2_evals(log_n, delta, eps, f, t, h1, h2)


def test_z1_identity_like_reference_test():
    """permutation/mod.rs:328-392 checks z1(1) = 1 and the product telescopes for a valid permutation."""
    rnd = random.Random(9)
    log_n, n = 3, 8
    w = pyref.root_of_unity(log_n)
    roots = [pow(w, i, pyref.R_MOD) for i in range(n)]
    ids = [roots, [pyref.K1 * r % pyref.R_MOD for r in roots], [pyref.K2 * r % pyref.R_MOD for r in roots]]
    flat = [v for col in ids for v in col]
    perm = list(range(3 * n))
    rnd.shuffle(perm)
    # values constant on permutation cycles => the grand product closes to 1
    vals = [None] * (3 * n)
    for i in range(3 * n):
        if vals[i] is None:
            v, j = rnd.randrange(pyref.R_MOD), i
            while vals[j] is None:
                vals[j] = v
                j = perm[j]
    sig = [flat[perm[i]] for i in range(3 * n)]
    a, b, c = vals[:n], vals[n:2 * n], vals[2 * n:]
    s1, s2, s3 = sig[:n], sig[n:2 * n], sig[2 * n:]
    beta, gamma = rnd.randrange(pyref.R_MOD), rnd.randrange(pyref.R_MOD)
    z = pyref.z1_evals(log_n, beta, gamma, a, b, c, s1, s2, s3)
    assert z[0] == 1
    p = pyref.R_MOD
    i = n - 1
    num = (beta * roots[i] + a[i] + gamma) * (pyref.K1 * beta * roots[i] + b[i] + gamma) * (pyref.K2 * beta * roots[i] + c[i] + gamma)
    den = (beta * s1[i] + a[i] + gamma) * (beta * s2[i] + b[i] + gamma) * (beta * s3[i] + c[i] + gamma)
    assert z[n - 1] * num % p * pow(den, -1, p) % p == 1


def test_quotient_and_free_tables_vs_python():
    rnd = random.Random(11)
    log_n = 2
    n4 = 4 << log_n
    x, zh, l1 = pyref.epk_free_tables(log_n)
    cx, czh, cl1 = cref.epk_free_tables(log_n)
    assert cref.limbs_to_ints(canon(cref.FR, cx)) == x
    assert cref.limbs_to_ints(canon(cref.FR, czh)) == zh
    assert cref.limbs_to_ints(canon(cref.FR, cl1)) == l1
    assert len(set(zh)) == 4                                   # zh takes 4 values on the 4n coset
    ch = dict(zip(("alpha", "beta", "gamma", "delta", "epsilon"), _rand_ints(rnd, 5)))
    wit = {k: _rand_ints(rnd, n4) for k in cref.WIT_ORDER}
    epk = {k: _rand_ints(rnd, n4) for k in cref.EPK_ORDER}
    epk["x"], epk["zh"], epk["l1"] = x, zh, l1
    m = lambda v: mont(cref.FR, cref.ints_to_limbs(v))
    got = cref.quotient_evals(log_n, m([ch[k] for k in ("alpha", "beta", "gamma", "delta", "epsilon")]),
                              {k: m(v) for k, v in wit.items()}, {k: m(v) for k, v in epk.items()})
    assert cref.limbs_to_ints(canon(cref.FR, got)) == pyref.quotient_coset_evals(log_n, ch, wit, epk)


def test_c_polynomial_helpers_match_python_integers():
    """zko_poly_eval / zko_poly_lincomb / zko_poly_divide_linear against the definitions in Python integers."""
    import random
    from oracle import plonk_ref
    from zkt_plonk_b200.prover import P, Poly, ints_to_mont_array, mont_array_to_ints
    rnd = random.Random(5)
    be = plonk_ref.OracleBackend(np.zeros((1, 8), dtype=np.uint64))
    ref = plonk_ref.PythonIntPolyOps
    for n in (1, 2, 7, 64, 257):
        cf = [rnd.randrange(P) for _ in range(n)]
        z = rnd.randrange(P)
        poly = Poly(ints_to_mont_array(cf), n)
        assert be.evaluate(poly, z) == ref.evaluate(cf, z)
        quot, ev = be.divide_linear(poly, z)
        w, e = ref.divide_linear(cf, z)
        assert ev == e and quot.len == n - 1 and mont_array_to_ints(quot.data[: quot.len]) == w
        # (X - z) * quot + p(z) == p
        if n == 1:
            assert e == cf[0]
            continue
        back = [(-(z * w[0]) + e) % P] + [(w[k - 1] - z * w[k]) % P for k in range(1, n - 1)] + [w[-1]]
        assert back == cf
    polys = [[rnd.randrange(P) for _ in range(m)] for m in (5, 9, 1, 9)]
    scalars = [rnd.randrange(P) for _ in polys]
    got = be.lincomb([Poly(ints_to_mont_array(p), len(p)) for p in polys], scalars, cap=12)
    assert got.len == 9 and mont_array_to_ints(got.data[:9]) == ref.lincomb(polys, scalars) and not got.data[9:].any()
    assert be.evaluate(Poly(np.zeros((1, 4), dtype=np.uint64), 0), 5) == 0


def test_permutation_coset_constants_like_the_reference():
    """permutation/constants.rs:33-50 (test_constants) on BN254: K1 = 7 and K2 = 13 generate valid, distinct cosets of the
    largest radix-2 subgroup (size 2^28): K1^(2^28) != 1, K2^(2^28) != 1 and (K1 / K2)^(2^28) != 1; also the domain generator
    the NTTs use is the one arkworks derives (g = 5, omega = 5^((r - 1) / 2^28))."""
    r = pyref.R_MOD
    n = 1 << 28
    assert (r - 1) % n == 0 and (r - 1) % (2 * n) != 0                      # two-adicity of Fr is exactly 28
    assert pyref.K1 == 7 and pyref.K2 == 13
    assert pow(pyref.K1, n, r) != 1 and pow(pyref.K2, n, r) != 1
    assert pow(pyref.K1 * pow(pyref.K2, -1, r) % r, n, r) != 1
    w = pow(5, (r - 1) // n, r)
    assert w == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert pow(w, n, r) == 1 and pow(w, n // 2, r) == r - 1
    from zkt_plonk_b200 import field
    assert field.root_of_unity(28) == w and field.root_of_unity(10) == pow(w, 1 << 18, r)


def test_z2_identity_with_the_reference_tests_vectors():
    """lookup/mod.rs:101-164 (test_compute_z2_poly): t = {0,0,1,2,3,4,5,6}, f = {3,6,0,5,4,3,2,0}, (h1, h2) = combine_split;
    for every domain element  (1+d)(e+f)(d t(wx) + e(1+d) + t) z2(x) == (d h2 + e(1+d) + h1)(d h1(wx) + e(1+d) + h2) z2(wx)
    and z2(1) = 1.  On the domain the polynomial evaluations are the vectors themselves (next row = index + 1 mod n)."""
    from zkt_plonk_b200 import prover
    p = pyref.R_MOD
    rnd = random.Random(31)
    t, f = [0, 0, 1, 2, 3, 4, 5, 6], [3, 6, 0, 5, 4, 3, 2, 0]
    h1, h2 = prover.combine_split(t, f)
    assert len(h1) == len(h2) == 8 and sorted(h1 + h2) == sorted(t + f)
    delta, eps = rnd.randrange(p), rnd.randrange(p)
    z2 = cref.limbs_to_ints(cref.from_mont(cref.FR, cref.z2_evals(
        3, cref.to_mont(cref.FR, cref.ints_to_limbs([delta]))[0], cref.to_mont(cref.FR, cref.ints_to_limbs([eps]))[0],
        *[cref.to_mont(cref.FR, cref.ints_to_limbs(v)) for v in (f, t, h1, h2)])))
    assert z2 == pyref.z2_evals(3, delta, eps, f, t, h1, h2) and z2[0] == 1
    opd, eopd = (1 + delta) % p, eps * (1 + delta) % p
    for i in range(8):
        j = (i + 1) % 8
        part_1 = opd * (eps + f[i]) % p * (delta * t[j] + eopd + t[i]) % p * z2[i] % p
        part_2 = (delta * h2[i] + eopd + h1[i]) % p * (delta * h1[j] + eopd + h2[i]) % p * z2[j] % p
        assert part_1 == part_2, i
