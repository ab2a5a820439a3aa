"""TEST INFRASTRUCTURE (oracle): the inner-product-argument polynomial commitment in Python integers.

Restates ark-poly-commit 0.3 `ipa_pc::InnerProductArgPC` -- the reference's second `PC`, `IPA<G, D>` of
plonk-core/src/commitment.rs:49-86 with D = Blake2s (plonk-core/src/test.rs:73,84) -- for ONE polynomial without hiding and
without degree bounds, which is how the prover calls its PC (prove.rs passes `None` for the rng; commitment.rs:81 leaves
shifted_comm = None).  The dependency is not vendored under /root/reference (Cargo.toml: ark-poly-commit = "0.3"); its open /
succinct_check / check are restated from the published algorithm (Buenz-Chiesa-Mishra-Spooner 2020, section 3 and appendix A):
PARITY UNPINNED against arkworks -- there is no golden vector for this scheme in the reference and no cargo here.  What pins
it: completeness and soundness properties (tests/test_ipa.py: an opening checks, a wrong value / tampered L / wrong point does
not), the closed form of the final key (G_final = <h, G> with h(X) = prod (1 + x_i X^(2^(k-1-i)))), and agreement with the GPU
rounds bit for bit.  The byte encodings that feed the hash (`to_bytes!` of affine points and field elements) are recalled from
ark-ec / ark-ff 0.3 and are the one part no property can check; the hash is injectable for that reason.

Only tests/ may import this module.
"""
import hashlib

from oracle import pyref

PROTOCOL_NAME = b"PC-DL-2020"


def _fq_bytes():
    return 8 * ((pyref.Q_MOD.bit_length() + 63) // 64)


def fr_bytes(v):
    return int(v % pyref.R_MOD).to_bytes(32, "little")


def g1_bytes(pt):
    """ark-ec 0.3 `impl ToBytes for GroupAffine`: x, y (canonical little endian), then the infinity flag as one byte."""
    nb = _fq_bytes()
    if pt is None:
        return (0).to_bytes(nb, "little") + (1).to_bytes(nb, "little") + b"\x01"
    return int(pt[0]).to_bytes(nb, "little") + int(pt[1]).to_bytes(nb, "little") + b"\x00"


def random_oracle_challenge(data):
    """compute_random_oracle_challenge: Blake2s(data || i as u64) for i = 0, 1, .. until Fr::from_random_bytes accepts (the
    digest with the bits above the modulus' length cleared, if below the modulus)."""
    bits = pyref.R_MOD.bit_length()
    i = 0
    while True:
        h = hashlib.blake2s(data + i.to_bytes(8, "little")).digest()
        v = int.from_bytes(h, "little") & ((1 << bits) - 1)
        if v < pyref.R_MOD:
            return v
        i += 1


def msm(points, scalars):
    acc = None
    for p, s in zip(points, scalars):
        if s % pyref.R_MOD:
            acc = pyref.g1_add(acc, pyref.g1_mul(s % pyref.R_MOD, p))
    return acc


def commit(comm_key, coeffs):
    """cm_commit(comm_key, coeffs, None, None)"""
    return msm(comm_key[: len(coeffs)], coeffs)


def inner(a, b):
    return sum(x * y for x, y in zip(a, b)) % pyref.R_MOD


def check_poly_coeffs(challenges):
    """SuccinctCheckPolynomial::compute_coeffs: h(X) = prod_i (1 + x_i X^(2^(k-1-i)))"""
    coeffs = [1]
    for ch in reversed(challenges):
        coeffs = coeffs + [c * ch % pyref.R_MOD for c in coeffs]
    return coeffs


def check_poly_eval(challenges, point):
    r, k, acc = pyref.R_MOD, len(challenges), 1
    for i, ch in enumerate(challenges):
        acc = acc * (1 + ch * pow(point, 1 << (k - 1 - i), r)) % r
    return acc


def open_(comm_key, h, coeffs, commitment, point, oracle=random_oracle_challenge):
    """InnerProductArgPC::open for one polynomial: returns (l_vec, r_vec, final_comm_key, c) and the round challenges."""
    r = pyref.R_MOD
    n = len(comm_key)
    assert n & (n - 1) == 0 and len(coeffs) <= n
    c = [v % r for v in coeffs] + [0] * (n - len(coeffs))
    z = [pow(point, i, r) for i in range(n)]
    value = inner(c, z)
    x = oracle(g1_bytes(commitment) + fr_bytes(point) + fr_bytes(value))
    h_prime = pyref.g1_mul(x, h)
    key = list(comm_key)
    l_vec, r_vec, challenges = [], [], []
    while n > 1:
        half = n // 2
        c_l, c_r, z_l, z_r, k_l, k_r = c[:half], c[half:n], z[:half], z[half:n], key[:half], key[half:n]
        L = pyref.g1_add(msm(k_l, c_r), pyref.g1_mul(inner(c_r, z_l), h_prime) if inner(c_r, z_l) else None)
        R = pyref.g1_add(msm(k_r, c_l), pyref.g1_mul(inner(c_l, z_r), h_prime) if inner(c_l, z_r) else None)
        l_vec.append(L)
        r_vec.append(R)
        x = oracle(fr_bytes(x) + g1_bytes(L) + g1_bytes(R))
        challenges.append(x)
        xi = pow(x, -1, r)
        c = [(a + xi * b) % r for a, b in zip(c_l, c_r)]
        z = [(a + x * b) % r for a, b in zip(z_l, z_r)]
        key = [pyref.g1_add(a, pyref.g1_mul(x, b)) for a, b in zip(k_l, k_r)]
        n = half
    return (l_vec, r_vec, key[0], c[0]), challenges


def combine(polys, commitments, values, opening_challenge):
    """What open / check do with several polynomials at one point: weights xi^(2 j) for polynomial j -- ipa_pc draws two opening
    challenges per polynomial, the odd one for the shifted polynomial of a degree bound (none here); opening_challenges(i) =
    xi^i (ark-poly-commit 0.3 PolynomialCommitment::open)."""
    r = pyref.R_MOD
    n = max((len(p) for p in polys), default=0)
    coeffs, C, v = [0] * n, None, 0
    for j, p in enumerate(polys):
        w = pow(opening_challenge, 2 * j, r)
        for i, c in enumerate(p):
            coeffs[i] = (coeffs[i] + w * c) % r
        C = pyref.g1_add(C, pyref.g1_mul(w, commitments[j]))
        v = (v + w * values[j]) % r
    return coeffs, C, v


def check(comm_key, h, commitment, point, value, proof, oracle=random_oracle_challenge):
    """succinct_check followed by the linear-time check of the final key."""
    r = pyref.R_MOD
    l_vec, r_vec, final_key, c = proof
    if 1 << len(l_vec) != len(comm_key) or len(r_vec) != len(l_vec):
        return False
    x = oracle(g1_bytes(commitment) + fr_bytes(point) + fr_bytes(value))
    h_prime = pyref.g1_mul(x, h)
    acc = pyref.g1_add(commitment, pyref.g1_mul(value % r, h_prime) if value % r else None)
    challenges = []
    for L, R in zip(l_vec, r_vec):
        x = oracle(fr_bytes(x) + g1_bytes(L) + g1_bytes(R))
        challenges.append(x)
        acc = pyref.g1_add(acc, pyref.g1_add(pyref.g1_mul(pow(x, -1, r), L) if L else None, pyref.g1_mul(x, R) if R else None))
    v_prime = check_poly_eval(challenges, point) * c % r
    if acc != msm([final_key, h_prime], [c, v_prime]):
        return False
    return final_key == msm(comm_key, check_poly_coeffs(challenges))
