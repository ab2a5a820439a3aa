"""CPU backend for zkt_plonk_b200.prover's round schedule and a restated verifier.  TEST INFRASTRUCTURE ONLY.

`OracleBackend` implements every heavy step of the prover with the C oracle (oracle/zkb_oracle.c) or plain Python
integers, so that tests can run proof_system::prove (plonk-core/src/proof_system/prove.rs:59-470) entirely on the
CPU and compare the proof bytes with the CUDA path.  `verify` restates Proof::verify
(plonk-core/src/proof_system/proof.rs:285-503), including compute_r0 (:163-217) and
compute_linearization_commitment (:220-282).  PC::check (SonicKZG10::check's accumulate_elems / check_elems with no
degree bounds and no hiding) is available in two forms: with `cvk = (H, beta H)` it is the reference's own product of
pairings e(C - v*G + z*W, H) * e(-W, beta*H) == 1 on the restated pairing of the selected curve (oracle/pairing.py for BN254,
oracle/pairing_bls.py for BLS12-381 / BLS12-377); without it,
because the synthetic SRS's trapdoor tau is known, the same equation is checked in G1 as  tau*W == C - v*G + z*W.
PARITY UNPINNED against the Rust binary: see zkb_oracle.c's header.
"""
import numpy as np

from oracle import cref, pairing, pairing_bls, pyref
from zkt_plonk_b200 import field
from zkt_plonk_b200.prover import (Poly, Proof, fr_to_limbs, ints_to_mont_array, limbs_to_fr, mont_array_to_ints,
                                    point_to_ints)
from zkt_plonk_b200.transcript import TRANSCRIPTS

def _o():
    """the C oracle of the curve the Python mirror currently works on (field.use_curve)"""
    return cref.oracle(field.CURVE)


EPK_ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_lookup", "q_table", "sigma1", "sigma2", "sigma3", "l1")
WIT_ORDER = ("z1", "z2", "a", "b", "c", "pi", "t", "h1", "h2")


class OracleBackend:
    def __init__(self, srs_points):
        self.srs = np.ascontiguousarray(srs_points, dtype=np.uint64)       # (N, 8) Montgomery affine powers of tau

    # -- storage
    def from_host(self, a, cap=None):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        buf = np.zeros((a.shape[0] if cap is None else cap, 4), dtype=np.uint64)
        buf[: a.shape[0]] = a
        return buf

    def to_host(self, buf, length):
        return buf[:length].copy()

    def slice_copy(self, buf, lo, hi, cap):
        out = np.zeros((cap, 4), dtype=np.uint64)
        out[: hi - lo] = buf[lo:hi]
        return out

    def get(self, buf, i):
        return limbs_to_fr(buf[i])

    def put(self, buf, i, value):
        buf[i] = fr_to_limbs(value)

    # -- transforms / commitments
    def ifft(self, evals, log_n, cap):
        n = 1 << log_n
        out = np.zeros((cap, 4), dtype=np.uint64)
        out[:n] = _o().ntt(np.ascontiguousarray(evals[:n]), log_n, True)
        return out

    def effective_len(self, buf, n):
        nz = np.flatnonzero(buf[:n].any(axis=1))
        return int(nz[-1]) + 1 if nz.size else 0

    def add_blinders(self, poly, blinders):
        for i, b in enumerate(blinders):
            poly.data[poly.len + i] = fr_to_limbs(b)
        for i, b in enumerate(blinders):        # prove.rs:472-483: extend first, then coeffs[i] -= b_i for EVERY i < k
            poly.data[i] = fr_to_limbs((limbs_to_fr(poly.data[i]) - b) % field.R_MOD)
        poly.len += len(blinders)

    def commit(self, poly):
        if poly.len == 0:
            return None
        sc = _o().from_mont(cref.FR, np.ascontiguousarray(poly.data[: poly.len]))
        xy, inf = _o().msm_g1(self.srs[: poly.len], sc)
        return point_to_ints(xy, inf)

    # -- argument math
    def z1_poly(self, log_n, beta, gamma, a, b, c, s1, s2, s3, cap):
        n = 1 << log_n
        ev = _o().z1_evals(log_n, fr_to_limbs(beta), fr_to_limbs(gamma), *[np.ascontiguousarray(x[:n]) for x in (a, b, c, s1, s2, s3)])
        return self.ifft(ev, log_n, cap)

    def z2_poly(self, log_n, delta, eps, f, t, h1, h2, cap):
        n = 1 << log_n
        ev = _o().z2_evals(log_n, fr_to_limbs(delta), fr_to_limbs(eps), *[np.ascontiguousarray(x[:n]) for x in (f, t, h1, h2)])
        return self.ifft(ev, log_n, cap)

    def _coset4(self, poly, log_n):
        buf = np.zeros((4 << log_n, 4), dtype=np.uint64)
        buf[: poly.len] = poly.data[: poly.len]
        return _o().ntt(buf, log_n + 2, False, True)

    def extend_prover_key(self, log_n, polys):
        epk = {name: self._coset4(polys[name], log_n) for name in EPK_ORDER[:-1]}
        epk["x"], epk["zh"], epk["l1"] = _o().epk_free_tables(log_n)
        return epk

    def quotient(self, log_n, epk, ch, polys):
        wit = {name: self._coset4(polys[name], log_n) for name in WIT_ORDER}
        chal = np.stack([fr_to_limbs(x) for x in ch])
        ev = _o().quotient_evals(log_n, chal, wit, epk)
        return _o().ntt(ev, log_n + 2, True, True)

    def evaluate(self, poly, z):
        return limbs_to_fr(_o().poly_eval(poly.data[: poly.len], fr_to_limbs(z))) if poly.len else 0

    def lincomb(self, polys, scalars, cap=None):
        m = max(p.len for p in polys)
        sc = np.stack([fr_to_limbs(x % field.R_MOD) for x in scalars])
        buf = np.zeros((cap or m, 4), dtype=np.uint64)
        buf[:m] = _o().poly_lincomb([p.data[: p.len] for p in polys], sc, m)
        return Poly(buf, m)

    def divide_linear(self, poly, z):
        m = poly.len
        buf = np.zeros((max(m - 1, 1), 4), dtype=np.uint64)
        if m == 0:
            return Poly(buf, 0), 0
        quot, ev = _o().poly_divide_linear(poly.data[:m], fr_to_limbs(z))
        if m > 1:
            buf[: m - 1] = quot
        return Poly(buf, m - 1), limbs_to_fr(ev)


class PythonIntPolyOps:
    """The three polynomial helpers above in plain Python integers (the definitions): tests check the C versions against
    these on random polynomials."""

    @staticmethod
    def evaluate(coeffs, z):
        acc = 0
        for cf in reversed(coeffs):
            acc = (acc * z + cf) % field.R_MOD
        return acc

    @staticmethod
    def lincomb(polys, scalars):
        m = max(len(p) for p in polys)
        acc = [0] * m
        for p, s in zip(polys, scalars):
            for k, cf in enumerate(p):
                acc[k] = (acc[k] + s * cf) % field.R_MOD
        return acc

    @staticmethod
    def divide_linear(cf, z):
        m = len(cf)
        w, carry = [0] * max(m - 1, 0), 0
        for k in range(m - 1, 0, -1):
            carry = (cf[k] + z * carry) % field.R_MOD
            w[k - 1] = carry
        return w, ((cf[0] if m else 0) + z * carry) % field.R_MOD


# ------------------------------------------------------------------------------------------------ verifier
def _lagrange(n, point, zh_eval, tau):
    """util.rs:185-195 compute_lagrange_evaluation(n, point, zh_eval, tau)."""
    return zh_eval * point % field.R_MOD * pow(n * (tau - point) % field.R_MOD, -1, field.R_MOD) % field.R_MOD


def _lin_comb_points(points, scalars):
    acc = None
    for pt, s in zip(points, scalars):
        acc = pyref.g1_add(acc, pyref.g1_mul(s % field.R_MOD, pt) if pt is not None else None)
    return acc


_BLS_PAIRINGS = {}


def _bls_pairing():
    if field.CURVE not in _BLS_PAIRINGS:
        _BLS_PAIRINGS[field.CURVE] = pairing_bls.Pairing(field.CURVE)
    return _BLS_PAIRINGS[field.CURVE]


def make_cvk(tau):
    """The G2 half of sonic_pc::VerifierKey for a synthetic SRS: (h, beta_h) = (H, tau * H), H the alt_bn128 G2 generator on
    BN254, the standard G2 generator on BLS12-381, a derived point of order r on BLS12-377 (oracle/pairing_bls.py).
    KZG10::setup draws h at random; any h gives the same accept / reject decisions."""
    if field.CURVE != "bn254":
        e = _bls_pairing()
        return (e.g2, e.g2_mul(tau % field.R_MOD, e.g2))
    return (pairing.G2_GEN, pairing.g2_mul(tau % field.R_MOD, pairing.G2_GEN))


def _kzg_check(commits, point, values, w, eta, tau, cvk=None):
    """SonicKZG10::check for one query point: A = sum eta^i C_i - (sum eta^i v_i) G + z W, then
    e(A, h) * e(-W, beta_h) == 1 (cvk given) or, with the trapdoor in place of the pairing, tau * W == A."""
    chal = [pow(eta, i, field.R_MOD) for i in range(len(commits))]
    c = _lin_comb_points(commits, chal)
    v = sum(e * x for e, x in zip(chal, values)) % field.R_MOD
    rhs = pyref.g1_add(pyref.g1_add(c, pyref.g1_neg(pyref.g1_mul(v, pyref.G1_GEN))), pyref.g1_mul(point, w) if w else None)
    if cvk is not None:
        check = pairing.pairing_product_is_one if field.CURVE == "bn254" else _bls_pairing().pairing_product_is_one
        return check([(rhs, cvk[0]), (pyref.g1_neg(w) if w else None, cvk[1])])
    lhs = pyref.g1_mul(tau, w) if w else None
    return lhs == rhs


def verify(vk, proof, pub_inputs, tau=None, transcript="merlin", cvk=None):
    """Proof::verify (proof.rs:285-503).  Returns 0 if accepted, else the failing step (1 or 2).
    transcript: "merlin" (transcript.rs:49-109) or "ethereum" (gadgets/src/transcript.rs:8-90).
    cvk = make_cvk(tau): PC::check by pairings, as the reference does; otherwise tau itself is needed."""
    assert tau is not None or cvk is not None
    n = vk.n
    log_n = n.bit_length() - 1
    assert len(pub_inputs) == len(vk.pi_roots), "invalid length of public inputs"
    tr = TRANSCRIPTS[transcript][1]("ZKT Plonk")
    vk.seed_transcript(tr)
    tr.append_scalars("pi", pub_inputs)
    C, E = proof.commits, proof.evals
    for k in ("a", "b", "c", "t", "h1", "h2"):
        tr.append_commitment(k + "_commit", C[k])
    beta, gamma = tr.challenge_scalar("beta"), tr.challenge_scalar("gamma")
    delta, epsilon = tr.challenge_scalar("delta"), tr.challenge_scalar("epsilon")
    assert len({beta, gamma, delta, epsilon}) == 4
    tr.append_commitment("z1_commit", C["z1"])
    tr.append_commitment("z2_commit", C["z2"])
    alpha = tr.challenge_scalar("alpha")
    for k in ("q_lo", "q_mid", "q_hi"):
        tr.append_commitment(k + "_commit", C[k])
    xi = tr.challenge_scalar("xi")
    zh = (pow(xi, n, field.R_MOD) - 1) % field.R_MOD
    l1 = _lagrange(n, 1, zh, xi)
    al2 = alpha * alpha % field.R_MOD
    opd = (1 + delta) % field.R_MOD
    eopd = epsilon * opd % field.R_MOD
    # compute_r0 (proof.rs:163-217)
    part1 = (-sum(_lagrange(n, pt, zh, xi) * pi for pi, pt in zip(pub_inputs, vk.pi_roots))) % field.R_MOD
    part2 = alpha * E["z1_next"] % field.R_MOD * (E["a"] + beta * E["sigma1"] + gamma) % field.R_MOD * (E["b"] + beta * E["sigma2"] + gamma) % field.R_MOD * (E["c"] + gamma) % field.R_MOD
    part3 = l1 * al2 % field.R_MOD
    part4 = al2 * alpha % field.R_MOD * E["z2_next"] % field.R_MOD * (eopd + delta * E["h2"]) % field.R_MOD * (eopd + E["h2"] + delta * E["h1_next"]) % field.R_MOD
    part5 = l1 * al2 % field.R_MOD * al2 % field.R_MOD
    r0 = (part1 + part2 + part3 + part4 + part5) % field.R_MOD
    # compute_linearization_commitment (proof.rs:220-282 with keys/*::compute_linearization_commitment)
    V = vk.commits
    bz = beta * xi % field.R_MOD
    al3, al4, al5 = al2 * alpha % field.R_MOD, al2 * al2 % field.R_MOD, al2 * al2 % field.R_MOD * alpha % field.R_MOD
    scalars = [E["a"] * E["b"] % field.R_MOD, E["a"], E["b"], E["c"], 1,
               (alpha * (bz + E["a"] + gamma) % field.R_MOD * (bz * field.K1 + E["b"] + gamma) % field.R_MOD * (bz * field.K2 + E["c"] + gamma) + l1 * al2) % field.R_MOD,
               (-alpha * beta % field.R_MOD * E["z1_next"] % field.R_MOD * (beta * E["sigma1"] + E["a"] + gamma) % field.R_MOD * (beta * E["sigma2"] + E["b"] + gamma)) % field.R_MOD,
               (al3 * opd % field.R_MOD * (epsilon + E["q_lookup"] * E["c"]) % field.R_MOD * (eopd + E["t"] + delta * E["t_next"]) + al4 * l1) % field.R_MOD,
               (-al3 * E["z2_next"] % field.R_MOD * (eopd + E["h2"] + delta * E["h1_next"])) % field.R_MOD,
               al5 * E["t"] % field.R_MOD]
    points = [V["q_m"], V["q_l"], V["q_r"], V["q_o"], V["q_c"], C["z1"], V["sigma3"], C["z2"], C["h1"], V["q_table"]]
    xn2 = (zh + 1) * xi % field.R_MOD * xi % field.R_MOD
    scalars += [(-zh) % field.R_MOD, (-zh * xn2) % field.R_MOD, (-zh * xn2 % field.R_MOD * xn2) % field.R_MOD]
    points += [C["q_lo"], C["q_mid"], C["q_hi"]]
    r_commit = _lin_comb_points(points, scalars)
    for k in Proof.EVALS:
        tr.append_scalar(k + "_eval", E[k])
    eta = tr.challenge_scalar("eta")
    ok1 = _kzg_check([r_commit, C["a"], C["b"], C["c"], V["sigma1"], V["sigma2"], V["q_lookup"], C["t"], C["h2"]], xi,
                     [r0, E["a"], E["b"], E["c"], E["sigma1"], E["sigma2"], E["q_lookup"], E["t"], E["h2"]], proof.aw, eta, tau, cvk)
    if not ok1:
        return 1
    w_n = field.root_of_unity(log_n)
    ok2 = _kzg_check([C["z1"], C["z2"], C["t"], C["h1"]], xi * w_n % field.R_MOD,
                     [E["z1_next"], E["z2_next"], E["t_next"], E["h1_next"]], proof.saw, eta, tau, cvk)
    return 0 if ok2 else 2


def make_srs_host(n_points, tau):
    """[tau^i] G for i < n_points, Montgomery affine (n_points, 2 * FQ_WORDS) -- small sizes only (double-and-add per point)."""
    w = field.FQ_WORDS
    G = _o().to_mont(cref.FQ, np.array([field.int_to_limbs(v, w) for v in field.G1_GENERATOR], dtype=np.uint64)).reshape(2 * w)
    powers, x = [], 1
    for _ in range(n_points):
        powers.append(x)
        x = x * tau % field.R_MOD
    return _o().g1_mul(G, cref.ints_to_limbs(powers))
