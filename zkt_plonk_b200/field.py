"""BN254 scalar-field metadata the host side needs (O(1) Python-int arithmetic, never bulk data).

Mirrors ark-bn254 0.3 FrParameters / ark-poly 0.3 Radix2EvaluationDomain::new, which plonk-core reaches
through `D::new` (plonk-core/src/proof_system/prove.rs:77, quotient_poly.rs:46).
"""
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
TWO_ADICITY = 28
GENERATOR = 5                      # Fr::multiplicative_generator(): the coset shift of coset_fft
TWO_ADIC_ROOT_OF_UNITY = pow(GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
MONT_R = 1 << 256
K1, K2 = 7, 13                     # plonk-core/src/permutation/constants.rs:13-20
G1_GENERATOR = (1, 2)


def to_mont(x, p=R_MOD):
    return (x % p) * MONT_R % p


def from_mont(x, p=R_MOD):
    return x * pow(MONT_R, -1, p) % p


def int_to_limbs(x):
    return [(x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]


def limbs_to_int(l):
    return sum(int(v) << (64 * k) for k, v in enumerate(l))


def root_of_unity(log_n):
    if log_n > TWO_ADICITY:
        return None
    w = TWO_ADIC_ROOT_OF_UNITY
    for _ in range(log_n, TWO_ADICITY):
        w = w * w % R_MOD
    return w
