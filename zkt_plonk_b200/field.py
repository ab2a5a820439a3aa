"""Scalar- and base-field metadata the host side needs (O(1) Python-int arithmetic, never bulk data).

Mirrors ark-{bn254,bls12-381,bls12-377} 0.3 FrParameters / ark-poly 0.3 Radix2EvaluationDomain::new, which plonk-core
reaches through `D::new` (plonk-core/src/proof_system/prove.rs:77, quotient_poly.rs:46).  The reference is generic over
the curve (`ZKTPlonk<F, D, PC, ..>`; plonk.rs:226-254 tests Bls12_381 and Bls12_377); here the Python mirror works on ONE
curve at a time: BN254 unless `use_curve` / `with curve(..)` selects another (the round driver, the transcript encodings and
the synthetic circuits read the constants below at call time).
"""
import contextlib

CURVES = {
    "bn254": dict(
        r=0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
        q=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
        two_adicity=28, generator=5, b=3, g1=(1, 2)),
    "bls12_381": dict(
        r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
        q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
        two_adicity=32, generator=7, b=4,
        g1=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
            0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)),
    "bls12_377": dict(
        r=0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
        q=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
        two_adicity=47, generator=22, b=1,
        g1=(0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
            0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6)),
}
K1, K2 = 7, 13                     # plonk-core/src/permutation/constants.rs:13-20 (the same on every curve)
MONT_R = 1 << 256                  # Fr is four 64-bit limbs on every curve


def use_curve(name):
    """Select the curve the Python mirror works on (process-wide)."""
    global CURVE, R_MOD, Q_MOD, TWO_ADICITY, GENERATOR, TWO_ADIC_ROOT_OF_UNITY, CURVE_B, G1_GENERATOR
    global FQ_WORDS, FQ_BYTES, MONT_RQ, RINV_R, RINV_Q
    c = CURVES[name]
    CURVE = name
    R_MOD, Q_MOD = c["r"], c["q"]
    TWO_ADICITY, GENERATOR = c["two_adicity"], c["generator"]     # Fr::multiplicative_generator(): the coset shift of coset_fft
    TWO_ADIC_ROOT_OF_UNITY = pow(GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
    CURVE_B, G1_GENERATOR = c["b"], c["g1"]
    FQ_WORDS = (Q_MOD.bit_length() + 63) // 64                    # 4 (BN254) or 6 (BLS12-381 / 377)
    FQ_BYTES = 8 * FQ_WORDS
    MONT_RQ = 1 << (64 * FQ_WORDS)
    RINV_R, RINV_Q = pow(MONT_R, -1, R_MOD), pow(MONT_RQ, -1, Q_MOD)


@contextlib.contextmanager
def curve(name):
    prev = CURVE
    use_curve(name)
    try:
        yield
    finally:
        use_curve(prev)


use_curve("bn254")


def to_mont(x, p=None):
    if p is None or p == R_MOD:
        return (x % R_MOD) * MONT_R % R_MOD
    return (x % p) * MONT_RQ % p


def from_mont(x, p=None):
    if p is None or p == R_MOD:
        return x * RINV_R % R_MOD
    return x * RINV_Q % p


def int_to_limbs(x, words=4):
    return [(x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(words)]


def limbs_to_int(l):
    return sum(int(v) << (64 * k) for k, v in enumerate(l))


def root_of_unity(log_n):
    if log_n > TWO_ADICITY:
        return None
    w = TWO_ADIC_ROOT_OF_UNITY
    for _ in range(log_n, TWO_ADICITY):
        w = w * w % R_MOD
    return w


def sqrt_q(a):
    """A square root of a in Fq, or None (q = 3 mod 4 on BN254 and BLS12-381; Tonelli-Shanks on BLS12-377, q = 1 mod 2^46)."""
    q = Q_MOD
    a %= q
    if a == 0:
        return 0
    if pow(a, (q - 1) // 2, q) != 1:
        return None
    if q % 4 == 3:
        return pow(a, (q + 1) // 4, q)
    s, t = 0, q - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    z = 2
    while pow(z, (q - 1) // 2, q) != q - 1:
        z += 1
    m, c, x, b = s, pow(z, t, q), pow(a, (t + 1) // 2, q), pow(a, t, q)
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2, i = b2 * b2 % q, i + 1
        e = pow(c, 1 << (m - i - 1), q)
        m, c, x, b = i, e * e % q, x * e % q, b * e % q * e % q
    return x
