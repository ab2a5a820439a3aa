"""CPU-only: pins oracle/pairing_bls.py (test infrastructure: the BLS12-381 / BLS12-377 pairing the restated verifier's
PC::check uses on those curves) by the defining properties of a pairing and by the curves' parameter identities."""
import random

import pytest

from oracle import pairing_bls


@pytest.mark.parametrize("curve", ["bls12_381", "bls12_377"])
def test_parameters_and_groups(curve):
    e = pairing_bls.Pairing(curve)
    x, q, r = e.x, e.q, e.r
    assert r == x ** 4 - x ** 2 + 1 and q == (x - 1) ** 2 * r // 3 + x                 # the BLS12 family
    assert (e.g1[1] ** 2 - e.g1[0] ** 3 - e.b) % q == 0 and e.g1_mul(r - 1, e.g1) == (e.g1[0], -e.g1[1] % q)
    assert e.g2_on_curve(e.g2) and e.g2_add(e.g2_mul(r - 1, e.g2), e.g2) is None       # order r on the twist
    # xi = xi0 + i is neither a square nor a cube in Fq2 (w^6 = xi is irreducible) and w^12 = A w^6 + B follows from it
    w6 = [0] * 6 + [1] + [0] * 5
    xi_sq = e.f12_mul(w6, w6)
    assert xi_sq == [e.B] + [0] * 5 + [e.A] + [0] * 5


@pytest.mark.parametrize("curve", ["bls12_381", "bls12_377"])
def test_bilinear_nondegenerate_order_r(curve):
    e = pairing_bls.Pairing(curve)
    rnd = random.Random(3)
    a, b = rnd.randrange(1, e.r), rnd.randrange(1, e.r)
    g = e.pairing(e.g1, e.g2)
    assert g != e.f12_one() and e.f12_pow(g, e.r) == e.f12_one()
    assert e.pairing(e.g1_mul(a, e.g1), e.g2) == e.f12_pow(g, a)
    assert e.pairing(e.g1, e.g2_mul(b, e.g2)) == e.f12_pow(g, b)
    assert e.pairing(e.g1_mul(a, e.g1), e.g2_mul(b, e.g2)) == e.f12_pow(g, a * b % e.r)
    # the shape of PC::check: e(A, H) * e(-W, tau H) == 1  <=>  A == tau W
    tau, k = rnd.randrange(1, e.r), rnd.randrange(1, e.r)
    W = e.g1_mul(k, e.g1)
    A = e.g1_mul(tau * k % e.r, e.g1)
    negW = (W[0], -W[1] % e.q)
    assert e.pairing_product_is_one([(A, e.g2), (negW, e.g2_mul(tau, e.g2))])
    assert not e.pairing_product_is_one([(A, e.g2), (negW, e.g2_mul(tau + 1, e.g2))])
    assert e.pairing(None, e.g2) == e.f12_one() and e.pairing(e.g1, None) == e.f12_one()


# ------------------------------------------------------------------------------------------------ the library's pairing (csrc/verify.cu)
@pytest.fixture
def on_curve(request):
    from oracle import pyref
    from zkt_plonk_b200 import field
    field.use_curve(request.param)
    pyref.use_curve(request.param)
    yield request.param
    field.use_curve("bn254")
    pyref.use_curve("bn254")


@pytest.mark.parametrize("on_curve", ["bls12_381", "bls12_377"], indirect=True)
def test_library_pairing_matches_the_python_restatement(on_curve):
    """zkb_pairing / zkb_g2_mul / zkb_pairing_product_is_one of the curve's build (host code: no GPU) against oracle/pairing_bls.py:
    the twelve Fq12 coefficients of e(aP, bQ) are equal one by one (same tower, same Miller function), the product check has the
    shape of PC::check, points off their curves are refused."""
    import numpy as np
    from zkt_plonk_b200 import verifier, _lib
    e = pairing_bls.Pairing(on_curve)
    rnd = random.Random(11)
    a, b = rnd.randrange(1, e.r), rnd.randrange(1, e.r)
    P, Q = e.g1_mul(a, e.g1), e.g2_mul(b, e.g2)
    assert verifier.pairing(e.g1, e.g2) == e.pairing(e.g1, e.g2)
    assert verifier.pairing(P, Q) == e.pairing(P, Q)
    got = verifier.g2_mul(b)                                      # b * (the curve's G2 point) through the ABI
    assert np.array_equal(got, verifier.g2_array(Q))
    tau, k = rnd.randrange(1, e.r), rnd.randrange(1, e.r)
    W, A = e.g1_mul(k, e.g1), e.g1_mul(tau * k % e.r, e.g1)
    negW = (W[0], -W[1] % e.q)
    assert verifier.pairing_product_is_one([(A, e.g2), (negW, e.g2_mul(tau, e.g2))])
    assert not verifier.pairing_product_is_one([(A, e.g2), (negW, e.g2_mul(tau + 1, e.g2))])
    assert verifier.pairing_product_is_one([(None, e.g2), (A, None)])
    with pytest.raises(_lib.ZkbError):
        verifier.pairing((e.g1[0], (e.g1[1] + 1) % e.q), e.g2)
    with pytest.raises(_lib.ZkbError):
        verifier.pairing(e.g1, (e.g2[0], ((e.g2[1][0] + 1) % e.q, e.g2[1][1])))
