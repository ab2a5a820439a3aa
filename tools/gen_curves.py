#!/usr/bin/env python3
"""Generates zkt_plonk_b200/csrc/curve_params.h and zkt_plonk_b200/csrc/ff_wide.cuh.

The reference is generic over the pairing engine (`ZKTPlonk<F, D, PC, ..>`; its own test_full runs on Bls12_381 and
Bls12_377, plonk-core/src/plonk.rs:226-254, the CLI fixes Bn254, bin/src/instance.rs:7-10).  Rust monomorphises per curve;
this library is compiled once per curve (-DZKB_CURVE=0/1/2 -> libzkb200.so / libzkb200_bls12_381.so /
libzkb200_bls12_377.so, same C ABI, element widths reported by zkb_curve_info).

curve_params.h: Montgomery constants of Fr and Fq (32-bit limbs for the device, 64-bit limbs for the host), the
two-adicity data ark-ff 0.3 FftParameters hold (GENERATOR, TWO_ADICITY, T = (r - 1) >> s) and the G1 generator, all
derived here from the moduli by Python big integers.  Checked by this script before anything is written: the generators
lie on their curves and have order r, 2^s | r - 1, GENERATOR^((r-1)/2) = -1, and GENERATOR^T equals the
TWO_ADIC_ROOT_OF_UNITY constants published in ark-bn254 / ark-bls12-381 / ark-bls12-377 0.3.

ff_wide.cuh: the carry chains of ff.cuh (one asm statement per chain) for 12 x 32-bit limbs.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "zkt_plonk_b200", "csrc")

CURVES = [
    dict(
        id=0, name="bn254",
        r=0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
        q=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
        b=3, gen=(1, 2), fr_gen=5,
        root=19103219067921713944291392827692070036145651957329286315305642004821462161904,
        # pairing: BN family, x = 4965661367192848881; Fq2 = Fq[i] / (i^2 + 1), xi = 9 + i, D-type twist
        family="bn", x=4965661367192848881, neg_beta=1, xi0=9, twist="D",
    ),
    dict(
        id=1, name="bls12_381",
        r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
        q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
        b=4,
        gen=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
             0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
        fr_gen=7,
        root=10238227357739495823651030575849232062558860180284477541189508159991286009131,
        # pairing: BLS12 family, x = -0xd201000000010000; Fq2 = Fq[i] / (i^2 + 1), xi = 1 + i, M-type twist
        family="bls12", x=-0xd201000000010000, neg_beta=1, xi0=1, twist="M",
    ),
    dict(
        id=2, name="bls12_377",
        r=0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
        q=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
        b=1,
        gen=(0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
             0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6),
        fr_gen=22,
        root=8065159656716812877374967518403273466521432693661810619979959746626482506078,
        # pairing: BLS12 family, x = 0x8508c00000000001; Fq2 = Fq[i] / (i^2 + 5), xi = i, D-type twist
        family="bls12", x=0x8508c00000000001, neg_beta=5, xi0=0, twist="D",
    ),
]


def ec_add(P, Q, p):
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0:
            return None
        lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
    x = (lam * lam - P[0] - Q[0]) % p
    return (x, (lam * (P[0] - x) - P[1]) % p)


def ec_mul(k, P, p):
    R = None
    while k:
        if k & 1:
            R = ec_add(R, P, p)
        P = ec_add(P, P, p)
        k >>= 1
    return R


def check(c):
    r, q = c["r"], c["q"]
    gx, gy = c["gen"]
    assert (gy * gy - gx ** 3 - c["b"]) % q == 0, "generator not on the curve"
    assert ec_mul(r, (gx, gy), q) is None, "generator order"
    s, t = 0, r - 1
    while t % 2 == 0:
        t //= 2
        s += 1
    c["s"], c["t"] = s, t
    assert pow(c["fr_gen"], (r - 1) // 2, r) == r - 1, "GENERATOR is a square"
    assert pow(c["fr_gen"], t, r) == c["root"], "TWO_ADIC_ROOT_OF_UNITY"


def f2_mul(a, b, q, neg_beta):
    return ((a[0] * b[0] - neg_beta * a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)


def f2_pow(a, e, q, neg_beta):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a, q, neg_beta)
        a = f2_mul(a, a, q, neg_beta)
        e >>= 1
    return r


def pairing_params(c):
    """Constants of the pairing (csrc/verify.cu), derived and checked: the family's polynomials give q and r from x; xi is
    neither a square nor a cube in Fq2 (so w^6 = xi defines Fq12); zeta = xi^((q^2 - 1) / 6) lies in Fq (w^(q^2) = zeta w)."""
    q, r, x = c["q"], c["r"], c["x"]
    if c["family"] == "bn":
        assert r == 36 * x ** 4 + 36 * x ** 3 + 18 * x ** 2 + 6 * x + 1 and q == 36 * x ** 4 + 36 * x ** 3 + 24 * x ** 2 + 6 * x + 1
    else:
        assert r == x ** 4 - x ** 2 + 1 and q == (x - 1) ** 2 * r // 3 + x
    xi = (c["xi0"], 1)
    assert f2_pow(xi, (q * q - 1) // 2, q, c["neg_beta"]) != (1, 0) and f2_pow(xi, (q * q - 1) // 3, q, c["neg_beta"]) != (1, 0)
    zeta = f2_pow(xi, (q * q - 1) // 6, q, c["neg_beta"])
    assert zeta[1] == 0
    assert (q ** 4 - q ** 2 + 1) % r == 0
    return dict(zeta=zeta[0], hard=(q ** 4 - q ** 2 + 1) // r, final=(q ** 12 - 1) // r)


def limbs(x, n, bits):
    return [(x >> (bits * i)) & ((1 << bits) - 1) for i in range(n)]


def field(p):
    n64 = (p.bit_length() + 63) // 64
    R = 1 << (64 * n64)
    return dict(p=p, n64=n64, n32=2 * n64, bits=p.bit_length(), one=R % p, r2=R * R % p,
                inv64=(-pow(p, -1, 1 << 64)) % (1 << 64), inv32=(-pow(p, -1, 1 << 32)) % (1 << 32))


def dev_struct(name, f):
    def arr(x):
        ls = limbs(x, f["n32"], 32)
        rows = [", ".join("0x%08xu" % v for v in ls[i:i + 4]) for i in range(0, len(ls), 4)]
        return (",\n" + " " * 35).join(rows)

    out = ["struct %s {" % name,
           "    static constexpr int N = %d;                  // 32-bit limbs" % f["n32"],
           "    static constexpr int BITS = %d;              // bit length of the modulus" % f["bits"],
           "    static constexpr bool LAZY = %s;          // two spare bits in 8 limbs: dedicated squaring and shared reductions (ff.cuh)"
           % ("true" if f["n32"] == 8 and f["bits"] <= 254 else "false")]
    for fn, val, note in (("mod", f["p"], "the modulus"), ("one", f["one"], "R mod p"), ("r2", f["r2"], "R^2 mod p")):
        out += ["    static __host__ __device__ __forceinline__ constexpr uint32_t %s(int i) {   // %s" % (fn, note),
                "        constexpr uint32_t m[%d] = {%s};" % (f["n32"], arr(val)),
                "        return m[i];", "    }"]
    out += ["    static constexpr uint32_t INV = 0x%08xu;     // -p^-1 mod 2^32" % f["inv32"], "};"]
    return "\n".join(out)


def host_arr(x, n):
    return "{" + ", ".join("0x%016xULL" % v for v in limbs(x, n, 64)) + "}"


def host_params(name, f):
    return ("static const ParamsT<%d> %s = {\n    %s,\n    %s,\n    %s,\n    0x%016xULL};"
            % (f["n64"], name, host_arr(f["p"], f["n64"]), host_arr(f["one"], f["n64"]), host_arr(f["r2"], f["n64"]), f["inv64"]))


def glv_params(c, trials=2000):
    """Constants of the G1 endomorphism phi(x, y) = (beta x, y) = lambda (x, y) (q = r = 1 mod 3 on all three curves) and of the
    split k = k1 + lambda k2 (mod r) with |k1|, |k2| ~ sqrt(r) (Gallant-Lambert-Vanstone): a short basis (a1, b1), (a2, b2) of the
    lattice {(a, b): a + b lambda = 0 mod r} from the extended Euclid sequence of (r, lambda); rounding by multiply-and-shift,
    C1 = (k G1) >> 384 with G1 = round(2^384 |b2| / r), C2 likewise with |b1|; k2 = C1 M1 + C2 M2, k1 = k - lambda k2 (mod r) with
    the signs folded into M1 = -sign(b2) b1, M2 = sign(b1) b2.  Checked here on random and edge scalars: the halves stay below
    2^130 (csrc/ipa.cu re-checks every split at run time and falls back to the plain scalar if one is not short)."""
    import random
    q, r, G = c["q"], c["r"], c["gen"]

    def root3(p):
        g = 2
        while pow(g, (p - 1) // 3, p) == 1:
            g += 1
        return pow(g, (p - 1) // 3, p)

    beta, lam = root3(q), root3(r)
    if ec_mul(lam, G, q) != (beta * G[0] % q, G[1]):
        beta = beta * beta % q
    assert ec_mul(lam, G, q) == (beta * G[0] % q, G[1]) and (lam * lam + lam + 1) % r == 0 and (beta * beta + beta + 1) % q == 0
    rows, r0, r1, t0, t1 = [(r, 0)], r, lam, 0, 1
    while r1:
        k = r0 // r1
        r0, r1, t0, t1 = r1, r0 - k * r1, t1, t0 - k * t1
        rows.append((r0, t0))
    l = max(i for i, (ri, _) in enumerate(rows) if ri * ri >= r)
    v1 = (rows[l + 1][0], -rows[l + 1][1])
    cands = [(rows[l][0], -rows[l][1])] + ([(rows[l + 2][0], -rows[l + 2][1])] if l + 2 < len(rows) else [])
    v2 = min(cands, key=lambda v: v[0] * v[0] + v[1] * v[1])
    (a1, b1), (a2, b2) = v1, v2
    assert (a1 + b1 * lam) % r == 0 and (a2 + b2 * lam) % r == 0 and a1 * b2 - a2 * b1 in (r, -r)
    sg = lambda v: -1 if v < 0 else 1
    g1, g2 = ((abs(b2) << 384) + r // 2) // r, ((abs(b1) << 384) + r // 2) // r
    m1, m2 = (-sg(b2) * b1) % r, (sg(b1) * b2) % r
    assert g1 < 1 << 320 and g2 < 1 << 320
    rnd = random.Random(c["id"])
    worst = 0
    for i in range(trials):
        k = [0, 1, r - 1, lam, r - lam, (r - 1) // 2][i] if i < 6 else rnd.randrange(r)
        k2 = (((k * g1) >> 384) * m1 + ((k * g2) >> 384) * m2) % r
        k1 = (k - lam * k2) % r
        s1, s2 = min(k1, r - k1), min(k2, r - k2)
        worst = max(worst, s1, s2)
        assert (k1 + lam * k2) % r == k
    assert worst < 1 << 130, worst.bit_length()
    return dict(beta=beta, lam=lam, g1=g1, g2=g2, m1=m1, m2=m2, bits=worst.bit_length())


def gen_params():
    o = ["// curve_params.h -- GENERATED by tools/gen_curves.py (do not edit): field and group constants of the curves the",
         "// library can be compiled for.  The reference is generic over the pairing engine (plonk.rs:226-254 runs its full test on",
         "// Bls12_381 and Bls12_377; the CLI fixes Bn254, bin/src/instance.rs:7-10); here the curve is a compile-time choice:",
         "// -DZKB_CURVE=0 (BN254, default), 1 (BLS12-381), 2 (BLS12-377).  Montgomery form, R = 2^(32 N).",
         "#pragma once", "#include <stdint.h>", "",
         "#define ZKB_CURVE_BN254 0", "#define ZKB_CURVE_BLS12_381 1", "#define ZKB_CURVE_BLS12_377 2",
         "#ifndef ZKB_CURVE", "#define ZKB_CURVE ZKB_CURVE_BN254", "#endif", "",
         "#ifndef __CUDACC__", "#ifndef __host__", "#define __host__", "#define __device__", "#define __forceinline__ inline",
         "#define ZKB_UNDEF_CUDA_QUALIFIERS", "#endif", "#endif", "",
         "namespace zkb {", "namespace host {",
         "template <int L> struct ParamsT { uint64_t p[L], one[L], r2[L], inv; };   // modulus, R mod p, R^2 mod p, -p^-1 mod 2^64",
         "}  // namespace host", ""]
    for k, c in enumerate(CURVES):
        check(c)
        fr, fq = field(c["r"]), field(c["q"])
        o.append(("#if" if k == 0 else "#elif") + " ZKB_CURVE == %d   // ---------------------------------------- %s" % (c["id"], c["name"]))
        o.append('#define ZKB_CURVE_NAME "%s"' % c["name"])
        o.append(dev_struct("FrP", fr))
        o.append(dev_struct("FqP", fq))
        o.append("namespace host {")
        o.append("constexpr int FR_L = %d, FQ_L = %d;             // 64-bit limbs of a scalar / a base-field element" % (fr["n64"], fq["n64"]))
        o.append(host_params("FR", fr))
        o.append(host_params("FQ", fq))
        o.append("// ark-ff 0.3 FftParameters of Fr: GENERATOR (also the coset generator ark-poly uses), TWO_ADICITY, T = (r - 1) >> TWO_ADICITY")
        o.append("constexpr uint64_t FR_GENERATOR = %d;" % c["fr_gen"])
        o.append("constexpr unsigned FR_TWO_ADICITY = %d;" % c["s"])
        o.append("static const uint64_t FR_T[%d] = %s;" % (fr["n64"], host_arr(c["t"], fr["n64"])))
        o.append("// G1: y^2 = x^3 + %d; the generator in Montgomery form" % c["b"])
        o.append("constexpr uint64_t G1_COEFF_B = %d;" % c["b"])
        o.append("static const uint64_t G1_GEN_X[%d] = %s;" % (fq["n64"], host_arr(c["gen"][0] * (1 << (64 * fq["n64"])) % c["q"], fq["n64"])))
        o.append("static const uint64_t G1_GEN_Y[%d] = %s;" % (fq["n64"], host_arr(c["gen"][1] * (1 << (64 * fq["n64"])) % c["q"], fq["n64"])))
        gl = glv_params(c)
        R4 = 1 << 256
        o.append("// GLV split for the key fold of csrc/ipa.cu: phi(x, y) = (beta x, y) = lambda (x, y); k2 = C1 M1 + C2 M2, k1 = k - lambda k2 with")
        o.append("// C = (k G) >> 384; halves below 2^%d on %d generator-checked scalars (tools/gen_curves.py glv_params)" % (gl["bits"], 2000))
        o.append("static const uint64_t GLV_BETA[%d] = %s;   // Montgomery" % (fq["n64"], host_arr(gl["beta"] * (1 << (64 * fq["n64"])) % c["q"], fq["n64"])))
        o.append("static const uint64_t GLV_LAMBDA_MONT[4] = %s;" % host_arr(gl["lam"] * R4 % c["r"], 4))
        o.append("static const uint64_t GLV_M1_MONT[4] = %s;" % host_arr(gl["m1"] * R4 % c["r"], 4))
        o.append("static const uint64_t GLV_M2_MONT[4] = %s;" % host_arr(gl["m2"] * R4 % c["r"], 4))
        o.append("static const uint64_t GLV_G1[5] = %s;" % host_arr(gl["g1"], 5))
        o.append("static const uint64_t GLV_G2[5] = %s;" % host_arr(gl["g2"], 5))
        pp = pairing_params(c)
        nh, nf = (pp["hard"].bit_length() + 63) // 64, (pp["final"].bit_length() + 63) // 64
        o.append("// pairing (csrc/verify.cu): Fq2 = Fq[i] / (i^2 + %d), w^6 = xi = %d + i, %s-type twist, %s family with x = %s0x%x"
                 % (c["neg_beta"], c["xi0"], c["twist"], c["family"].upper(), "-" if c["x"] < 0 else "", abs(c["x"])))
        o.append("constexpr int PAIRING_IS_BLS12 = %d, PAIRING_TWIST_M = %d, PAIRING_X_NEG = %d;" % (c["family"] == "bls12", c["twist"] == "M", c["x"] < 0))
        o.append("constexpr uint64_t FQ2_NEG_BETA = %d, PAIRING_XI0 = %d, PAIRING_X_ABS = 0x%016xULL;" % (c["neg_beta"], c["xi0"], abs(c["x"])))
        o.append("static const uint64_t PAIRING_ZETA[%d] = %s;   // xi^((q^2 - 1) / 6), canonical: w^(q^2) = zeta w" % (fq["n64"], host_arr(pp["zeta"], fq["n64"])))
        o.append("constexpr int PAIRING_HARD_EXP_LIMBS = %d, PAIRING_FINAL_EXP_LIMBS = %d;" % (nh, nf))
        o.append("static const uint64_t PAIRING_HARD_EXP[%d] = %s;   // (q^4 - q^2 + 1) / r" % (nh, host_arr(pp["hard"], nh)))
        o.append("static const uint64_t PAIRING_FINAL_EXP[%d] = %s;   // (q^12 - 1) / r" % (nf, host_arr(pp["final"], nf)))
        o.append("}  // namespace host")
    o += ["#else", '#error "ZKB_CURVE must be 0 (BN254), 1 (BLS12-381) or 2 (BLS12-377)"', "#endif", "",
          "}  // namespace zkb", "",
          "#ifdef ZKB_UNDEF_CUDA_QUALIFIERS", "#undef __host__", "#undef __device__", "#undef __forceinline__",
          "#undef ZKB_UNDEF_CUDA_QUALIFIERS", "#endif", ""]
    return "\n".join(o)


# ---------------------------------------------------------------------------------------------- 12-limb carry chains
def gen_wide(N=12):
    H = N // 2   # products per chain
    o = ["// ff_wide.cuh -- GENERATED by tools/gen_curves.py (do not edit): the carry chains of ff.cuh for %d x 32-bit limbs" % N,
         "// (BLS12-381 / BLS12-377 base fields).  Same layout as the 8-limb chains: an `even' accumulator E (limb positions",
         "// 0 .. %d) and an `odd' accumulator O (positions 1 .. %d); a chain multiplies the limbs a[PAR], a[PAR + 2], .. by one" % (N - 1, N),
         "// word and adds the %d aligned 64-bit products, ONE asm statement per chain (the carry flag never crosses statements)." % H,
         "#pragma once", "#include <stdint.h>", "", "namespace zkb {", "#ifdef __CUDACC__", ""]
    acc_out = ", ".join('"=r"(acc[%d])' % i for i in range(N))
    acc_io = ", ".join('"+r"(acc[%d])' % i for i in range(N))

    def a_ops():
        return ", ".join('"r"(a[PAR + %d])' % (2 * k) for k in range(H))

    # row_mul
    body = []
    for k in range(H):
        body.append('"mul.lo.u32 %%%d, %%%d, %%%d;\\n\\t" "mul.hi.u32 %%%d, %%%d, %%%d;%s"'
                    % (2 * k, N + k, N + H, 2 * k + 1, N + k, N + H, "" if k == H - 1 else "\\n\\t"))
    o += ["// acc = {a[PAR], a[PAR+2], ..} * b as %d aligned 64-bit products (no carries)" % H,
          "template <int PAR>",
          "__device__ __forceinline__ void wrow_mul(uint32_t (&acc)[%d], const uint32_t (&a)[%d], uint32_t b) {" % (N, N),
          "    asm(" + "\n        ".join(body),
          "        : " + acc_out, "        : " + a_ops() + ', "r"(b));', "}", ""]
    # row_mad_cin: operands: acc 0..N-1, lo = N, x = N+1, a = N+2.., b = N+2+H
    body = ['"add.cc.u32 %%%d, %%%d, %%%d;\\n\\t"' % (N, N, N + 1)]
    for k in range(H):
        last = k == H - 1
        body.append('"madc.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t" "madc.hi%s.u32 %%%d, %%%d, %%%d, %%%d;%s"'
                    % (2 * k, N + 2 + k, N + 2 + H, 2 * k, "" if last else ".cc", 2 * k + 1, N + 2 + k, N + 2 + H, 2 * k + 1,
                       "" if last else "\\n\\t"))
    o += ["// lo += x (carry c);  acc += {a[PAR], ..} * b + c   (the carry out of the top limb is provably zero)",
          "template <int PAR>",
          "__device__ __forceinline__ void wrow_mad_cin(uint32_t &lo, uint32_t x, uint32_t (&acc)[%d], const uint32_t (&a)[%d], uint32_t b) {" % (N, N),
          "    asm(" + "\n        ".join(body),
          "        : " + acc_io + ', "+r"(lo)', '        : "r"(x), ' + a_ops() + ', "r"(b));', "}", ""]
    # row_mad: acc 0..N-1, a = N.., b = N+H
    body = []
    for k in range(H):
        last = k == H - 1
        body.append('"%s.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t" "madc.hi%s.u32 %%%d, %%%d, %%%d, %%%d;%s"'
                    % ("mad" if k == 0 else "madc", 2 * k, N + k, N + H, 2 * k, "" if last else ".cc", 2 * k + 1, N + k, N + H, 2 * k + 1,
                       "" if last else "\\n\\t"))
    o += ["// acc += {a[PAR], ..} * b   (no carry in; the carry out of the top limb is provably zero)",
          "template <int PAR>",
          "__device__ __forceinline__ void wrow_mad(uint32_t (&acc)[%d], const uint32_t (&a)[%d], uint32_t b) {" % (N, N),
          "    asm(" + "\n        ".join(body),
          "        : " + acc_io, "        : " + a_ops() + ', "r"(b));', "}", ""]
    # row_mad_cout: acc 0..N-1, top = N, a = N+1.., b = N+1+H
    body = []
    for k in range(H):
        body.append('"%s.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t" "madc.hi.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t"'
                    % ("mad" if k == 0 else "madc", 2 * k, N + 1 + k, N + 1 + H, 2 * k, 2 * k + 1, N + 1 + k, N + 1 + H, 2 * k + 1))
    body.append('"addc.u32 %%%d, %%%d, 0;"' % (N, N))
    o += ["// acc += {a[PAR], ..} * b;  top += the carry out of the top limb",
          "template <int PAR>",
          "__device__ __forceinline__ void wrow_mad_cout(uint32_t (&acc)[%d], uint32_t &top, const uint32_t (&a)[%d], uint32_t b) {" % (N, N),
          "    asm(" + "\n        ".join(body),
          "        : " + acc_io + ', "+r"(top)', "        : " + a_ops() + ', "r"(b));', "}", ""]
    # add / sub: r 0..N-1, c = N, a = N+1.., b = 2N+1..
    for nm, op in (("wadd", "add"), ("wsub", "sub")):
        body = []
        for k in range(N):
            body.append('"%s.cc.u32 %%%d, %%%d, %%%d;\\n\\t"' % (op if k == 0 else op + "c", k, N + 1 + k, 2 * N + 1 + k))
        body.append('"%sc.u32 %%%d, 0, 0;"' % (op, N))
        o += ["// r = a %s b over %d limbs; returns the carry (add: 0 / 1, sub: 0 / 0xffffffff)" % ("+" if op == "add" else "-", N),
              "__device__ __forceinline__ uint32_t %s(uint32_t (&r)[%d], const uint32_t (&a)[%d], const uint32_t (&b)[%d]) {" % (nm, N, N, N),
              "    uint32_t c;",
              "    asm(" + "\n        ".join(body),
              "        : " + ", ".join('"=r"(r[%d])' % i for i in range(N)) + ', "=r"(c)',
              "        : " + ", ".join('"r"(a[%d])' % i for i in range(N)) + ", " + ", ".join('"r"(b[%d])' % i for i in range(N)) + ");",
              "    return c;", "}", ""]
    # triangular chains of the dedicated squaring: pairs below K0 hold no product of this row (their multiplicands are zero)
    # and only ripple the carry; same operand numbering as wrow_mad_cin / wrow_mad_cout
    o += ["// wtri_mad_cin<PAR, K0>: as wrow_mad_cin, but the products start at pair K0 (the pairs below it ripple the carry of lo += x)",
          "template <int PAR, int K0>",
          "__device__ __forceinline__ void wtri_mad_cin(uint32_t &lo, uint32_t x, uint32_t (&acc)[%d], const uint32_t (&a)[%d], uint32_t b) {" % (N, N),
          "    static_assert(K0 >= 0 && K0 < %d, \"a chain has %d pairs\");" % (H, H)]
    for K0 in range(H):
        body = ['"add.cc.u32 %%%d, %%%d, %%%d;\\n\\t"' % (N, N, N + 1)]
        for k in range(H):
            last = k == H - 1
            if k < K0:
                body.append('"addc.cc.u32 %%%d, %%%d, 0;\\n\\t" "addc.cc.u32 %%%d, %%%d, 0;\\n\\t"' % (2 * k, 2 * k, 2 * k + 1, 2 * k + 1))
            else:
                body.append('"madc.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t" "madc.hi%s.u32 %%%d, %%%d, %%%d, %%%d;%s"'
                            % (2 * k, N + 2 + k, N + 2 + H, 2 * k, "" if last else ".cc", 2 * k + 1, N + 2 + k, N + 2 + H, 2 * k + 1,
                               "" if last else "\\n\\t"))
        o += ["    %sif constexpr (K0 == %d) {" % ("" if K0 == 0 else "else ", K0),
              "        asm(" + "\n            ".join(body),
              "            : " + acc_io + ', "+r"(lo)', '            : "r"(x), ' + a_ops() + ', "r"(b));', "    }"]
    o += ["}", ""]
    o += ["// wtri_mad_cout<PAR, K0>: as wrow_mad_cout, the products start at pair K0 (K0 == %d: the row has no limb of this parity left)" % H,
          "template <int PAR, int K0>",
          "__device__ __forceinline__ void wtri_mad_cout(uint32_t (&acc)[%d], uint32_t &top, const uint32_t (&a)[%d], uint32_t b) {" % (N, N),
          "    static_assert(K0 >= 0 && K0 <= %d, \"a chain has %d pairs\");" % (H, H)]
    for K0 in range(H):
        body = []
        for k in range(K0, H):
            body.append('"%s.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t" "madc.hi.cc.u32 %%%d, %%%d, %%%d, %%%d;\\n\\t"'
                        % ("mad" if k == K0 else "madc", 2 * k, N + 1 + k, N + 1 + H, 2 * k, 2 * k + 1, N + 1 + k, N + 1 + H, 2 * k + 1))
        body.append('"addc.u32 %%%d, %%%d, 0;"' % (N, N))
        o += ["    %sif constexpr (K0 == %d) {" % ("" if K0 == 0 else "else ", K0),
              "        asm(" + "\n            ".join(body),
              "            : " + acc_io + ', "+r"(top)', "            : " + a_ops() + ', "r"(b));', "    }"]
    o += ["}", ""]
    # merge: r = E + (O << 32) + x : r 0..N-1, E = N..2N-1, x = 2N, O = 2N+1 .. 3N-1 (O[0..N-2])
    body = []
    for k in range(N):
        body.append('"%s%s.u32 %%%d, %%%d, %%%d;%s"' % ("add" if k == 0 else "addc", "" if k == N - 1 else ".cc", k, N + k, 2 * N + k,
                                                       "" if k == N - 1 else "\\n\\t"))
    o += ["// r = E + (O << 32) + x   (the result is below 2p: O's top limb is zero and nothing carries out)",
          "__device__ __forceinline__ void wmerge(uint32_t (&r)[%d], const uint32_t (&E)[%d], uint32_t x, const uint32_t (&O)[%d]) {" % (N, N, N),
          "    asm(" + "\n        ".join(body),
          "        : " + ", ".join('"=r"(r[%d])' % i for i in range(N)),
          "        : " + ", ".join('"r"(E[%d])' % i for i in range(N)) + ', "r"(x), ' + ", ".join('"r"(O[%d])' % i for i in range(N - 1)) + ");",
          "}", "", "#endif  // __CUDACC__", "}  // namespace zkb", ""]
    return "\n".join(o)


def gen_oracle_params():
    o = ["/* zko_curve_params.h -- GENERATED by tools/gen_curves.py (do not edit): the constants of oracle/zkb_oracle.c for the curve",
         " * it is compiled for (-DZKO_CURVE=0 BN254 (default), 1 BLS12-381, 2 BLS12-377).  TEST INFRASTRUCTURE ONLY. */",
         "#ifndef ZKO_CURVE", "#define ZKO_CURVE 0", "#endif"]
    for k, c in enumerate(CURVES):
        check(c)
        fr, fq = field(c["r"]), field(c["q"])
        o.append(("#if" if k == 0 else "#elif") + " ZKO_CURVE == %d   /* %s */" % (c["id"], c["name"]))
        o.append("#define ZKO_FQ_L %d              /* 64-bit limbs of a base-field element */" % fq["n64"])
        o.append("#define ZKO_FR_BITS %d" % fr["bits"])
        o.append("#define ZKO_FR_GENERATOR %d      /* Fr::multiplicative_generator(), ark-poly's coset generator */" % c["fr_gen"])
        o.append("#define ZKO_FR_TWO_ADICITY %d" % c["s"])
        o.append("#define ZKO_G1_B %d              /* y^2 = x^3 + b */" % c["b"])
        o.append("#define ZKO_FR_T_INIT %s   /* (r - 1) >> TWO_ADICITY */" % host_arr(c["t"], fr["n64"]))
        for nm, f in (("FR", fr), ("FQ", fq)):
            o.append("#define ZKO_%s_INIT {%s, \\\n    %s, \\\n    %s, \\\n    0x%016xULL}"
                     % (nm, host_arr(f["p"], f["n64"]), host_arr(f["one"], f["n64"]), host_arr(f["r2"], f["n64"]), f["inv64"]))
    o += ["#else", '#error "ZKO_CURVE must be 0, 1 or 2"', "#endif", ""]
    return "\n".join(o)


if __name__ == "__main__":
    with open(os.path.join(HERE, "..", "oracle", "zko_curve_params.h"), "w") as f:
        f.write(gen_oracle_params())
    with open(os.path.join(OUT_DIR, "curve_params.h"), "w") as f:
        f.write(gen_params())
    with open(os.path.join(OUT_DIR, "ff_wide.cuh"), "w") as f:
        f.write(gen_wide())
    print("wrote curve_params.h, ff_wide.cuh and oracle/zko_curve_params.h")
