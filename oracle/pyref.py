"""Independent big-int restatement of the zkt-plonk prover hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the arithmetic of this path lives in un-vendored crates.io dependencies of the
reference (ark-ff / ark-ec / ark-poly / ark-poly-commit / ark-bn254, all pinned "0.3" in
/root/reference/plonk-core/Cargo.toml:19-24); neither rustc nor those crates exist in this image and
the reference's own tests hold no golden vector for MSM/NTT/commitments.  This
file restates the *published definitions* (DFT over the 2-adic subgroup, sum of scalar multiples on
BN254 G1, the formulas in plonk-core's widget files) with Python integers in canonical (non-Montgomery)
form.  It is the second, independent implementation the C oracle (oracle/zkb_oracle.c, Montgomery
4x64 limbs) is checked against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import it; the product path (zkt_plonk_b200/) never does.

Call sites restated (reference file:line):
  * NTT / INTT / coset variants  -> plonk-core/src/util.rs:63-140 (wrappers over ark-poly
    Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place)
  * MSM                          -> plonk-core/src/commitment.rs:31-46 (VariableBaseMSM::multi_scalar_mul)
  * z1 grand product             -> plonk-core/src/permutation/mod.rs:181-257
  * z2 grand product             -> plonk-core/src/lookup/mod.rs:25-85
  * quotient on the 4n coset     -> plonk-core/src/proof_system/quotient_poly.rs:20-227 and
    keys/arithmetic.rs:67-81, keys/permutation.rs:97-137, keys/lookup.rs:81-122
"""

# ---------------------------------------------------------------- curve constants (ark-bn254 / ark-bls12-381 / ark-bls12-377 0.3)
# BN254 unless use_curve selects another; the definitions below read these globals at call time.
_CURVES = {
    "bn254": (0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
              0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47, 28, 5, 3, (1, 2)),
    "bls12_381": (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
                  0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 32, 7, 4,
                  (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
                   0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)),
    "bls12_377": (0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
                  0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001, 47, 22, 1,
                  (0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
                   0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6)),
}
K1 = 7   # plonk-core/src/permutation/constants.rs:13-15
K2 = 13  # plonk-core/src/permutation/constants.rs:18-20
MONT_R = 1 << 256   # Fr (and BN254's Fq): four 64-bit limbs


def use_curve(name):
    global CURVE, R_MOD, Q_MOD, TWO_ADICITY, FR_GENERATOR, TWO_ADIC_ROOT, CURVE_B, G1_GEN
    CURVE = name
    R_MOD, Q_MOD, TWO_ADICITY, FR_GENERATOR, CURVE_B, G1_GEN = _CURVES[name]   # FR_GENERATOR: Fr::multiplicative_generator()
    TWO_ADIC_ROOT = pow(FR_GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)


use_curve("bn254")


def to_mont(x, p):
    return (x * MONT_R) % p


def from_mont(x, p):
    return (x * pow(MONT_R, -1, p)) % p


def inv(x, p):
    return pow(x, -1, p)


# ---------------------------------------------------------------- domain
def root_of_unity(log_n):
    """Radix2EvaluationDomain::new: group_gen = TWO_ADIC_ROOT squared (TWO_ADICITY - log_n) times."""
    assert 0 <= log_n <= TWO_ADICITY
    w = TWO_ADIC_ROOT
    for _ in range(log_n, TWO_ADICITY):
        w = w * w % R_MOD
    return w


def dft_naive(a, log_n, inverse=False):
    """O(n^2) definition: A_k = sum_j a_j w^{jk}; short inputs are zero padded (fft_in_place resizes)."""
    n = 1 << log_n
    a = list(a) + [0] * (n - len(a))
    w = root_of_unity(log_n)
    if inverse:
        w = inv(w, R_MOD)
    out = []
    for k in range(n):
        wk = pow(w, k, R_MOD)
        acc, x = 0, 1
        for j in range(n):
            acc = (acc + a[j] * x) % R_MOD
            x = x * wk % R_MOD
        out.append(acc)
    if inverse:
        ninv = inv(n, R_MOD)
        out = [v * ninv % R_MOD for v in out]
    return out


def ntt(a, log_n, inverse=False):
    """O(n log n) recursive radix-2, natural order in and out."""
    n = 1 << log_n
    a = list(a) + [0] * (n - len(a))
    w = root_of_unity(log_n)
    if inverse:
        w = inv(w, R_MOD)

    def rec(v, w):
        m = len(v)
        if m == 1:
            return v
        e = rec(v[0::2], w * w % R_MOD)
        o = rec(v[1::2], w * w % R_MOD)
        out = [0] * m
        x = 1
        for k in range(m // 2):
            t = x * o[k] % R_MOD
            out[k] = (e[k] + t) % R_MOD
            out[k + m // 2] = (e[k] - t) % R_MOD
            x = x * w % R_MOD
        return out

    out = rec(a, w)
    if inverse:
        ninv = inv(n, R_MOD)
        out = [v * ninv % R_MOD for v in out]
    return out


def coset_ntt(a, log_n):
    """coset_fft_in_place: distribute_powers(coeffs, g=5) then fft."""
    n = 1 << log_n
    a = list(a) + [0] * (n - len(a))
    g, x = FR_GENERATOR, 1
    sc = []
    for v in a:
        sc.append(v * x % R_MOD)
        x = x * g % R_MOD
    return ntt(sc, log_n)


def coset_intt(a, log_n):
    """coset_ifft_in_place: ifft then distribute_powers(evals, g^-1)."""
    out = ntt(a, log_n, inverse=True)
    gi, x = inv(FR_GENERATOR, R_MOD), 1
    res = []
    for v in out:
        res.append(v * x % R_MOD)
        x = x * gi % R_MOD
    return res


# ---------------------------------------------------------------- G1 (affine, None = infinity)
def g1_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q_MOD)


def g1_add(P, Qp):
    if P is None:
        return Qp
    if Qp is None:
        return P
    x1, y1 = P
    x2, y2 = Qp
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * inv(2 * y1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * inv(x2 - x1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def g1_mul(k, P):
    acc = None
    k %= R_MOD
    while k:
        if k & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


def msm_naive(points, scalars):
    """Definition of VariableBaseMSM::multi_scalar_mul: sum_i s_i * P_i over min(len) pairs."""
    acc = None
    for P, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(s, P))
    return acc


# ---------------------------------------------------------------- grand products (evaluation form)
def z1_evals(log_n, beta, gamma, a, b, c, s1, s2, s3):
    """permutation/mod.rs:181-254, before the final iFFT."""
    n = 1 << log_n
    p = R_MOD
    w = root_of_unity(log_n)
    out, state, root = [1], 1, 1
    for i in range(n - 1):
        num = (beta * root + a[i] + gamma) * (K1 * beta * root + b[i] + gamma) % p \
            * (K2 * beta * root + c[i] + gamma) % p
        den = (beta * s1[i] + a[i] + gamma) * (beta * s2[i] + b[i] + gamma) % p \
            * (beta * s3[i] + c[i] + gamma) % p
        state = state * num % p * inv(den, p) % p
        out.append(state)
        root = root * w % p
    return out


def z2_evals(log_n, delta, epsilon, f, t, h1, h2):
    """lookup/mod.rs:25-82, before the final iFFT."""
    n = 1 << log_n
    p = R_MOD
    opd = (1 + delta) % p
    eopd = epsilon * opd % p
    out, state = [1], 1
    for i in range(n - 1):
        num = opd * (epsilon + f[i]) % p * (delta * t[i + 1] + eopd + t[i]) % p
        den = (delta * h2[i] + eopd + h1[i]) * (delta * h1[i + 1] + eopd + h2[i]) % p
        state = state * num % p * inv(den, p) % p
        out.append(state)
    return out


# ---------------------------------------------------------------- quotient on the 4n coset
def quotient_coset_evals(log_n, ch, wit, epk):
    """quotient_poly.rs:98-224.  `wit`/`epk` map names to length-4n coset evaluation lists.

    ch  = dict(alpha, beta, gamma, delta, epsilon)
    wit = dict(z1, z2, a, b, c, pi, t, h1, h2)
    epk = dict(q_m, q_l, q_r, q_o, q_c, q_lookup, q_table, sigma1, sigma2, sigma3, x, l1, zh)
    "next" (x*omega_n) is index i+4 mod 4n (quotient_poly.rs:53-94).
    """
    p = R_MOD
    n4 = 4 << log_n
    al, be, ga, de, ep = (ch[k] for k in ("alpha", "beta", "gamma", "delta", "epsilon"))
    al2 = al * al % p
    al3 = al2 * al % p
    al4 = al3 * al % p
    al5 = al4 * al % p
    opd = (1 + de) % p
    eopd = ep * opd % p
    out = []
    for i in range(n4):
        j = (i + 4) % n4
        a, b, c = wit["a"][i], wit["b"][i], wit["c"][i]
        arith = (a * b % p * epk["q_m"][i] + a * epk["q_l"][i] + b * epk["q_r"][i] + c * epk["q_o"][i]
                 + epk["q_c"][i] + wit["pi"][i]) % p
        bx = be * epk["x"][i] % p
        z1, z1n, l1 = wit["z1"][i], wit["z1"][j], epk["l1"][i]
        perm = (al * z1 % p * (bx + a + ga) % p * (bx * K1 + b + ga) % p * (bx * K2 + c + ga)
                - al * z1n % p * (be * epk["sigma1"][i] + a + ga) % p * (be * epk["sigma2"][i] + b + ga) % p
                * (be * epk["sigma3"][i] + c + ga)
                + (z1 - 1) * l1 % p * al2) % p
        t, tn, h1, h1n, h2 = wit["t"][i], wit["t"][j], wit["h1"][i], wit["h1"][j], wit["h2"][i]
        z2, z2n = wit["z2"][i], wit["z2"][j]
        look = (al3 * z2 % p * opd % p * (ep + epk["q_lookup"][i] * c) % p * (eopd + t + de * tn)
                - al3 * z2n % p * (eopd + h1 + de * h2) % p * (eopd + h2 + de * h1n)
                + al4 * (z2 - 1) % p * l1
                + al5 * epk["q_table"][i] % p * t) % p
        out.append((arith + perm + look) * inv(epk["zh"][i], p) % p)
    return out


def epk_free_tables(log_n):
    """x_coset, zh_coset, l_1_coset of keys/mod.rs:109-119 as plain definitions on the coset 5*<w_4n>."""
    n = 1 << log_n
    n4 = 4 * n
    p = R_MOD
    w = root_of_unity(log_n + 2)
    x, xs = FR_GENERATOR, []
    for _ in range(n4):
        xs.append(x)
        x = x * w % p
    zh = [(pow(v, n, p) - 1) % p for v in xs]
    ninv = inv(n, p)
    l1 = [z * ninv % p * inv((v - 1) % p, p) % p for z, v in zip(zh, xs)]
    return xs, zh, l1


# ---------------------------------------------------------------- limb helpers (4 x u64 little endian)
def to_limbs(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))
