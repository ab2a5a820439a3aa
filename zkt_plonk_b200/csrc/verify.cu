// verify.cu -- Proof::verify of zkt-plonk on the host.  No GPU work: the verifier is a 13-point linear
// combination, two KZG checks and a handful of field operations (milliseconds in the reference as well).
//
//   zkb_plonk_verify   plonk-core/src/proof_system/proof.rs:285-503 (compute_r0 :163-217, compute_linearization_commitment
//                      :220-282 with keys/{arithmetic,permutation,lookup}.rs::compute_linearization_commitment), then
//                      PC::check twice (:441-502) = SonicKZG10::check without degree bounds or hiding:
//                      e(sum eta^i C_i - (sum eta^i v_i) G + z W, h) * e(-W, beta_h) == 1
//   zkb_pairing        e(P, Q) on the curve of this build (ark-ec 0.3 / ark-bn254 / ark-bls12-381 / ark-bls12-377 0.3 are crates.io
//                      dependencies, un-vendored): optimal ate, restated in its simplest exact form -- Fq2 = Fq[i] / (i^2 + b)
//                      (b = 1, 1, 5), w^6 = xi = xi0 + i (9 + i, 1 + i, i), Fq12 = Fq[w] / (w^12 - 2 xi0 w^6 + xi0^2 + b) as twelve
//                      coefficients in Fq (BN254: w^12 - 18 w^6 + 82), G2 arithmetic on the twist in affine Fq2 coordinates,
//                      sparse line functions (D-type twist: BN254, BLS12-377; M-type: BLS12-381, lines times w^3), the Miller
//                      loop over 6x + 2 with the two Frobenius-twisted additions (BN) or over |x| (BLS12), final exponentiation
//                      (q^6 - 1)(q^2 + 1) and one power (q^4 - q^2 + 1) / r.  A product-of-pairings check has the same outcome
//                      under any correct pairing, so arkworks' internal conventions cannot change accept / reject.
// tests/test_verify.py compares zkb_pairing with the independent Python restatements (oracle/pairing.py, oracle/pairing_bls.py)
// coefficient by coefficient and zkb_plonk_verify with oracle/plonk_ref.py on accepted and tampered proofs, on every curve.
#include <string.h>

#include <vector>

#include "../../include/zkb200.h"
#include "host_plonk.h"

namespace {

using host::FQ;
using host::FR;

// ================================================================================================ constants
// (q^12 - 1) / r, (q^4 - q^2 + 1) / r, zeta = xi^((q^2 - 1) / 6), the twist and the family: curve_params.h (generated, checked there)
using host::PAIRING_FINAL_EXP;
using host::PAIRING_HARD_EXP;
#if ZKB_CURVE == ZKB_CURVE_BN254
// BN254 only: the Frobenius twists of the last two Miller steps and the loop count 6x + 2
const uint64_t FROB_X_C0[4] = {0x99e39557176f553dULL, 0xb78cc310c2c3330cULL, 0x4c0bec3cf559b143ULL, 0x2fb347984f7911f7ULL};
const uint64_t FROB_X_C1[4] = {0x1665d51c640fcba2ULL, 0x32ae2a1d0b7c9dceULL, 0x4ba4cc8bd75a0794ULL, 0x16c9e55061ebae20ULL};
const uint64_t FROB_Y_C0[4] = {0xdc54014671a0135aULL, 0xdbaae0eda9c95998ULL, 0xdc5ec698b6e2f9b9ULL, 0x063cf305489af5dcULL};
const uint64_t FROB_Y_C1[4] = {0x82d37f632623b0e3ULL, 0x21807dc98fa25bd2ULL, 0x0704b5a7ec796f2bULL, 0x07c03cbcac41049aULL};
const uint64_t FROB2_X_C0[4] = {0xe4bd44e5607cfd48ULL, 0xc28f069fbb966e3dULL, 0x5e6dd9e7e0acccb0ULL, 0x30644e72e131a029ULL};
const uint64_t ATE_LOOP_LOW = 0x9d797039be763ba8ULL;      // 6x + 2 = 2^64 + this (x = 4965661367192848881): top bit implicit

#endif

inline Fq q_from_canon(const uint64_t *c /* FQ_L words */) { Fq f, r2; memcpy(f.l, c, FQB); memcpy(r2.l, FQ.r2, FQB); return host::mul(f, r2, FQ); }
inline Fq q_small(uint64_t v) { return host::from_u64(v, FQ); }
inline Fq qadd(const Fq &a, const Fq &b) { return host::add(a, b, FQ); }
inline Fq qsub(const Fq &a, const Fq &b) { return host::sub(a, b, FQ); }
inline Fq qmul(const Fq &a, const Fq &b) { return host::mul(a, b, FQ); }
inline Fq qzero() { Fq z; memset(z.l, 0, sizeof z.l); return z; }
inline Fq qneg(const Fq &a) { return host::sub(qzero(), a, FQ); }
inline bool qeq(const Fq &a, const Fq &b) { return memcmp(a.l, b.l, FQB) == 0; }
inline Fq q_canonical(const Fq &m) { Fq one = qzero(); one.l[0] = 1; return host::mul(m, one, FQ); }

// ================================================================================================ Fq2 = Fq[i] / (i^2 + b), b = FQ2_NEG_BETA
inline Fq q_times_b(const Fq &a) { return host::FQ2_NEG_BETA == 1 ? a : qmul(q_small(host::FQ2_NEG_BETA), a); }
struct F2 { Fq c0, c1; };
inline F2 f2_add(const F2 &a, const F2 &b) { return {qadd(a.c0, b.c0), qadd(a.c1, b.c1)}; }
inline F2 f2_sub(const F2 &a, const F2 &b) { return {qsub(a.c0, b.c0), qsub(a.c1, b.c1)}; }
inline F2 f2_mul(const F2 &a, const F2 &b) {
    return {qsub(qmul(a.c0, b.c0), q_times_b(qmul(a.c1, b.c1))), qadd(qmul(a.c0, b.c1), qmul(a.c1, b.c0))};
}
inline F2 f2_conj(const F2 &a) { return {a.c0, qneg(a.c1)}; }
inline F2 f2_inv(const F2 &a) {
    Fq d = host::inv(qadd(qmul(a.c0, a.c0), q_times_b(qmul(a.c1, a.c1))), FQ);
    return {qmul(a.c0, d), qneg(qmul(a.c1, d))};
}
inline bool f2_is_zero(const F2 &a) { return host::is_zero(a.c0) && host::is_zero(a.c1); }
inline bool f2_eq(const F2 &a, const F2 &b) { return qeq(a.c0, b.c0) && qeq(a.c1, b.c1); }

struct G2 { F2 x, y; bool inf; };

// the twist: y^2 = x^3 + b / xi (D-type: BN254 3 / (9 + i), BLS12-377 1 / i) or y^2 = x^3 + b xi (M-type: BLS12-381 4 (1 + i))
bool g2_on_curve(const G2 &p) {
    if (p.inf) return true;
    F2 xi = {q_small(host::PAIRING_XI0), q_small(1)}, bb = {q_small(host::G1_COEFF_B), qzero()};
    F2 b2 = f2_mul(bb, host::PAIRING_TWIST_M ? xi : f2_inv(xi));
    return f2_eq(f2_sub(f2_mul(p.y, p.y), f2_mul(f2_mul(p.x, p.x), p.x)), b2);
}

// ================================================================================================ Fq12 = Fq[w] / (w^12 - A w^6 + C)
// w^6 = xi0 + i and i^2 = -b give (w^6 - xi0)^2 = -b:  A = 2 xi0, C = xi0^2 + b  (BN254: 18, 82; BLS12-381: 2, 2; BLS12-377: 0, 5)
constexpr uint64_t F12_A = 2 * host::PAIRING_XI0, F12_C = host::PAIRING_XI0 * host::PAIRING_XI0 + host::FQ2_NEG_BETA;
struct F12 { Fq c[12]; };
F12 f12_one() { F12 r; for (int i = 0; i < 12; ++i) r.c[i] = qzero(); r.c[0] = host::one(FQ); return r; }
bool f12_eq(const F12 &a, const F12 &b) { for (int i = 0; i < 12; ++i) if (!qeq(a.c[i], b.c[i])) return false; return true; }

F12 f12_mul(const F12 &a, const F12 &b) {
    Fq t[23];
    for (int k = 0; k < 23; ++k) t[k] = qzero();
    for (int i = 0; i < 12; ++i) {
        if (host::is_zero(a.c[i])) continue;
        for (int j = 0; j < 12; ++j) t[i + j] = qadd(t[i + j], qmul(a.c[i], b.c[j]));
    }
    const Fq kA = q_small(F12_A), kC = q_small(F12_C);
    for (int k = 22; k >= 12; --k) {                       // w^12 = A w^6 - C
        if (host::is_zero(t[k])) continue;
        if (F12_A) t[k - 6] = qadd(t[k - 6], qmul(kA, t[k]));
        t[k - 12] = qsub(t[k - 12], qmul(kC, t[k]));
    }
    F12 r;
    for (int k = 0; k < 12; ++k) r.c[k] = t[k];
    return r;
}

// a * line: a line function has five non-zero coefficients (60 products instead of 144) -- at w^0, w^1, w^3, w^7, w^9 on a
// D-type twist, at w^0, w^2, w^3, w^6, w^8 (the line times w^3) on an M-type twist
F12 f12_mul_line(const F12 &a, const F12 &l) {
    static const int NZ_D[5] = {0, 1, 3, 7, 9}, NZ_M[5] = {0, 2, 3, 6, 8};
    const int *NZ = host::PAIRING_TWIST_M ? NZ_M : NZ_D;
    Fq t[23];
    for (int k = 0; k < 23; ++k) t[k] = qzero();
    for (int i = 0; i < 12; ++i)
        for (int jj = 0; jj < 5; ++jj) t[i + NZ[jj]] = qadd(t[i + NZ[jj]], qmul(a.c[i], l.c[NZ[jj]]));
    const Fq kA = q_small(F12_A), kC = q_small(F12_C);
    for (int k = 20; k >= 12; --k) {
        if (F12_A) t[k - 6] = qadd(t[k - 6], qmul(kA, t[k]));
        t[k - 12] = qsub(t[k - 12], qmul(kC, t[k]));
    }
    F12 r;
    for (int k = 0; k < 12; ++k) r.c[k] = t[k];
    return r;
}

// q^6-power Frobenius: w -> -w
F12 f12_conj(const F12 &a) {
    F12 r = a;
    for (int k = 1; k < 12; k += 2) r.c[k] = qneg(a.c[k]);
    return r;
}

// q^2-power Frobenius: the coefficients are in Fq and w^(q^2) = zeta w with zeta a sixth root of unity of Fq
F12 f12_frob2(const F12 &a) {
    const Fq zeta = q_from_canon(host::PAIRING_ZETA);
    F12 r = a;
    Fq z = zeta;
    for (int k = 1; k < 12; ++k) { r.c[k] = qmul(a.c[k], z); z = qmul(z, zeta); }
    return r;
}

// a^-1 by the extended Euclidean algorithm in Fq[w] against w^12 - A w^6 + C (a != 0)
F12 f12_inv(const F12 &a) {
    auto deg = [](const Fq *p) { int d = 12; while (d > 0 && host::is_zero(p[d])) --d; return d; };
    Fq lm[13], hm[13], low[13], high[13];
    for (int i = 0; i < 13; ++i) { lm[i] = qzero(); hm[i] = qzero(); low[i] = i < 12 ? a.c[i] : qzero(); high[i] = qzero(); }
    lm[0] = host::one(FQ);
    high[0] = q_small(F12_C); high[6] = qneg(q_small(F12_A)); high[12] = host::one(FQ);
    while (deg(low) > 0) {
        // r = high div low (quotient only), then (lm, low, hm, high) <- (hm - lm r, high - low r, lm, low)
        Fq r[13], work[13];
        for (int i = 0; i < 13; ++i) { r[i] = qzero(); work[i] = high[i]; }
        const int dl = deg(low), dh = deg(high);
        const Fq lead_inv = host::inv(low[dl], FQ);
        for (int i = dh - dl; i >= 0; --i) {
            Fq c = qmul(work[dl + i], lead_inv);
            r[i] = c;
            if (!host::is_zero(c)) for (int j = 0; j <= dl; ++j) work[i + j] = qsub(work[i + j], qmul(c, low[j]));
        }
        Fq nm[13], nw[13];
        for (int i = 0; i < 13; ++i) { nm[i] = hm[i]; nw[i] = high[i]; }
        for (int i = 0; i < 13; ++i)
            for (int j = 0; j + i < 13; ++j) {
                nm[i + j] = qsub(nm[i + j], qmul(lm[i], r[j]));
                nw[i + j] = qsub(nw[i + j], qmul(low[i], r[j]));
            }
        for (int i = 0; i < 13; ++i) { hm[i] = lm[i]; high[i] = low[i]; lm[i] = nm[i]; low[i] = nw[i]; }
    }
    const Fq inv0 = host::inv(low[0], FQ);
    F12 out;
    for (int i = 0; i < 12; ++i) out.c[i] = qmul(lm[i], inv0);
    return out;
}

F12 f12_pow(const F12 &a, const uint64_t *e, int limbs) {
    F12 acc = f12_one();
    bool started = false;
    for (int i = 64 * limbs - 1; i >= 0; --i) {
        if (started) acc = f12_mul(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) { acc = started ? f12_mul(acc, a) : a; started = true; }
    }
    return acc;
}

// Fq2 -> two Fq12 coefficients: a0 + a1 i = (a0 - xi0 a1) + a1 w^6
inline void embed(const F2 &a, Fq *lo, Fq *hi) {
    *lo = host::PAIRING_XI0 ? qsub(a.c0, qmul(q_small(host::PAIRING_XI0), a.c1)) : a.c0;
    *hi = a.c1;
}

// line through the untwisted (xr, yr) with slope embed(lam) w, at P = (xp, yp) in E(Fq):
//   -yp + xp embed(lam) w + embed(yr - lam xr) w^3      (coefficients at w^0, w^1, w^7, w^3, w^9)
F12 sparse_line(const F2 &lam, const F2 &xr, const F2 &yr, const Fq &xp, const Fq &yp) {
    F12 l;
    for (int i = 0; i < 12; ++i) l.c[i] = qzero();
    Fq l0, l1, c0, c1;
    embed(lam, &l0, &l1);
    embed(f2_sub(yr, f2_mul(lam, xr)), &c0, &c1);
    if (host::PAIRING_TWIST_M) {                           // untwist (x', y') -> (x' / w^2, y' / w^3); the line times w^3 (an element
        l.c[3] = qneg(yp);                                 // of Fq4: the final exponentiation kills it):
        l.c[2] = qmul(l0, xp);                             //   -yp w^3 + xp embed(lam) w^2 + embed(yr - lam xr)
        l.c[8] = qmul(l1, xp);
        l.c[0] = c0;
        l.c[6] = c1;
        return l;
    }
    l.c[0] = qneg(yp);
    l.c[1] = qmul(l0, xp);
    l.c[7] = qmul(l1, xp);
    l.c[3] = c0;
    l.c[9] = c1;
    return l;
}

struct Miller {
    F12 f;
    F2 xr, yr;
    Fq xp, yp;
    // one step: tangent at R (square == true) or chord through R and (x2, y2); f <- f^2 * line or f * line; R <- 2R or R + Q
    void step(bool square, F2 x2, F2 y2) {
        F2 lam;
        if (square) {
            F2 three = {q_small(3), qzero()};
            lam = f2_mul(f2_mul(three, f2_mul(xr, xr)), f2_inv(f2_add(yr, yr)));
            x2 = xr;
            y2 = yr;
        } else {
            lam = f2_mul(f2_sub(y2, yr), f2_inv(f2_sub(x2, xr)));
        }
        F12 line = sparse_line(lam, xr, yr, xp, yp);
        f = f12_mul_line(square ? f12_mul(f, f) : f, line);
        F2 x3 = f2_sub(f2_sub(f2_mul(lam, lam), xr), x2);
        F2 y3 = f2_sub(f2_mul(lam, f2_sub(xr, x3)), yr);
        xr = x3;
        yr = y3;
    }
};

// Miller loop of the optimal ate pairing; the identity on either side gives 1.  (P, Q of prime order r: the chord / tangent
// denominators cannot vanish inside the loop.)
F12 miller_loop(const G2 &q, const Fq &xp, const Fq &yp, bool p_inf) {
    if (q.inf || p_inf) return f12_one();
    Miller m;
    m.f = f12_one();
    m.xr = q.x; m.yr = q.y; m.xp = xp; m.yp = yp;
#if ZKB_CURVE != ZKB_CURVE_BN254
    // BLS12: f_{|x|, Q}(P), bits of |x| below the leading one; x < 0: f_{-|x|} = 1 / f_{|x|} up to factors the final power
    // kills, and on the values that survive it the inverse is the q^6 Frobenius w -> -w
    int top = 63;
    while (!((host::PAIRING_X_ABS >> top) & 1)) --top;
    for (int i = top - 1; i >= 0; --i) {
        m.step(true, q.x, q.y);
        if ((host::PAIRING_X_ABS >> i) & 1) m.step(false, q.x, q.y);
    }
    return host::PAIRING_X_NEG ? f12_conj(m.f) : m.f;
#else
    for (int i = 63; i >= 0; --i) {
        m.step(true, q.x, q.y);
        if ((ATE_LOOP_LOW >> i) & 1) m.step(false, q.x, q.y);
    }
    const F2 frob_x = {q_from_canon(FROB_X_C0), q_from_canon(FROB_X_C1)}, frob_y = {q_from_canon(FROB_Y_C0), q_from_canon(FROB_Y_C1)};
    const F2 frob2_x = {q_from_canon(FROB2_X_C0), qzero()};
    F2 x1 = f2_mul(f2_conj(q.x), frob_x), y1 = f2_mul(f2_conj(q.y), frob_y);      // pi(Q)
    F2 x2 = f2_mul(q.x, frob2_x), y2 = q.y;                                        // -pi^2(Q)
    m.step(false, x1, y1);
    F2 lam = f2_mul(f2_sub(y2, m.yr), f2_inv(f2_sub(x2, m.xr)));
    return f12_mul_line(m.f, sparse_line(lam, m.xr, m.yr, xp, yp));
#endif
}

// f^((q^12 - 1) / r) as (q^6 - 1) (q^2 + 1) ((q^4 - q^2 + 1) / r): two Frobenius maps, one inversion and a 761-bit power
// instead of a 2790-bit one.  final_exponentiation_plain is the definition; zkb_pairing's tests see both agree.
F12 final_exponentiation_plain(const F12 &f) { return f12_pow(f, PAIRING_FINAL_EXP, host::PAIRING_FINAL_EXP_LIMBS); }
F12 final_exponentiation(const F12 &f) {
    bool zero = true;
    for (int i = 0; i < 12; ++i) if (!host::is_zero(f.c[i])) zero = false;
    if (zero) return f;                                    // cannot happen for points of order r; keep it total
    F12 t = f12_mul(f12_conj(f), f12_inv(f));
    t = f12_mul(f12_frob2(t), t);
    return f12_pow(t, PAIRING_HARD_EXP, host::PAIRING_HARD_EXP_LIMBS);
}

// ================================================================================================ G1 on the host (XYZZ)
struct X1 { Fq x, y, zz, zzz; };
inline bool x1_inf(const X1 &p) { return host::is_zero(p.zz); }
inline X1 x1_zero() { X1 p; memset(&p, 0, sizeof p); return p; }
inline X1 x1_from(const Pt &a) {
    if (a.inf) return x1_zero();
    return {a.x, a.y, host::one(FQ), host::one(FQ)};
}
X1 x1_dbl(const X1 &p) {
    if (x1_inf(p)) return p;
    Fq u = qadd(p.y, p.y), v = qmul(u, u), w = qmul(u, v), s = qmul(p.x, v);
    Fq xx = qmul(p.x, p.x), m = qadd(qadd(xx, xx), xx);
    X1 r;
    r.x = qsub(qsub(qmul(m, m), s), s);
    r.y = qsub(qmul(m, qsub(s, r.x)), qmul(w, p.y));
    r.zz = qmul(v, p.zz);
    r.zzz = qmul(w, p.zzz);
    return r;
}
X1 x1_add(const X1 &a, const X1 &b) {
    if (x1_inf(a)) return b;
    if (x1_inf(b)) return a;
    Fq u1 = qmul(a.x, b.zz), u2 = qmul(b.x, a.zz), s1 = qmul(a.y, b.zzz), s2 = qmul(b.y, a.zzz);
    Fq p = qsub(u2, u1), r = qsub(s2, s1);
    if (host::is_zero(p)) return host::is_zero(r) ? x1_dbl(a) : x1_zero();
    Fq pp = qmul(p, p), ppp = qmul(p, pp), q = qmul(u1, pp);
    X1 o;
    o.x = qsub(qsub(qsub(qmul(r, r), ppp), q), q);
    o.y = qsub(qmul(r, qsub(q, o.x)), qmul(s1, ppp));
    o.zz = qmul(qmul(a.zz, b.zz), pp);
    o.zzz = qmul(qmul(a.zzz, b.zzz), ppp);
    return o;
}
// scalar (Montgomery Fr) * point
X1 x1_mul(const Fe &k_mont, const Pt &p) {
    Fe one = {{1, 0, 0, 0}};
    Fe k = host::mul(k_mont, one, FR);                    // canonical
    X1 base = x1_from(p), acc = x1_zero();
    for (int i = 255; i >= 0; --i) {
        acc = x1_dbl(acc);
        if ((k.l[i >> 6] >> (i & 63)) & 1) acc = x1_add(acc, base);
    }
    return acc;
}
Pt x1_affine(const X1 &p) {
    Pt r;
    r.inf = x1_inf(p);
    if (r.inf) { r.x = qzero(); r.y = qzero(); return r; }
    Fq zi = host::inv(p.zzz, FQ);
    Fq zzi = qmul(zi, p.zz);
    zzi = qmul(zzi, zzi);
    r.x = qmul(p.x, zzi);
    r.y = qmul(p.y, zi);
    return r;
}
Pt pt_neg(const Pt &p) { Pt r = p; if (!p.inf) r.y = qneg(p.y); return r; }
bool g1_on_curve(const Pt &p) {
    if (p.inf) return true;
    return qeq(qmul(p.y, p.y), qadd(qmul(qmul(p.x, p.x), p.x), q_small(host::G1_COEFF_B)));
}

// a square root in Fq, if there is one: a^((q + 1) / 4) when q = 3 mod 4 (BN254, BLS12-381), Tonelli-Shanks otherwise
// (BLS12-377: q - 1 = 2^46 t); the exponents are derived from the modulus
bool fq_sqrt(const Fq &a, Fq *out) {
    if (host::is_zero(a)) { *out = a; return true; }
    constexpr int L = host::FQ_L;
    auto shr = [](uint64_t *v, unsigned k) {                // v >>= k, 0 < k < 64
        for (int i = 0; i < L; ++i) v[i] = (v[i] >> k) | (i + 1 < L ? v[i + 1] << (64 - k) : 0);
    };
    uint64_t e[L];
    memcpy(e, FQ.p, sizeof e);
    if ((FQ.p[0] & 3) == 3) {
        e[0] += 1;                                          // q + 1: no carry, q ends in ...11
        shr(e, 2);
        Fq y = host::pow(a, e, FQ);
        if (!qeq(qmul(y, y), a)) return false;
        *out = y;
        return true;
    }
    e[0] -= 1;                                              // q - 1 = 2^s t
    unsigned sbits = 0;
    while (!(e[0] & 1)) { shr(e, 1); ++sbits; }             // e = t (odd)
    uint64_t half[L];
    memcpy(half, FQ.p, sizeof half);
    half[0] -= 1;
    shr(half, 1);                                           // (q - 1) / 2: Euler's criterion
    const Fq one = host::one(FQ), minus_one = qneg(one);
    if (!qeq(host::pow(a, half, FQ), one)) return false;
    Fq z = q_small(2);
    while (!qeq(host::pow(z, half, FQ), minus_one)) z = qadd(z, one);
    uint64_t t1[L];
    memcpy(t1, e, sizeof t1);
    t1[0] += 1;                                             // (t + 1) / 2
    if (t1[0] == 0) for (int i = 1; i < L && ++t1[i] == 0; ++i) {}
    shr(t1, 1);
    unsigned m = sbits;
    Fq c = host::pow(z, e, FQ), x = host::pow(a, t1, FQ), b = host::pow(a, e, FQ);
    while (!qeq(b, one)) {
        unsigned i = 0;
        Fq b2 = b;
        while (!qeq(b2, one)) { b2 = qmul(b2, b2); ++i; }
        if (i >= m) return false;
        Fq g = c;
        for (unsigned k = 0; k + i + 1 < m; ++k) g = qmul(g, g);
        m = i;
        c = qmul(g, g);
        x = qmul(x, g);
        b = qmul(b, c);
    }
    *out = x;
    return true;
}

// GroupAffine::deserialize (compressed, ark-ec 0.3): x little endian (32 / 48 bytes), bit 6 of the last byte = infinity,
// bit 7 = (y > -y)
bool g1_decompress(const uint8_t *in /* FQB bytes */, Pt *out) {
    constexpr int L = host::FQ_L;
    Fq x;
    memcpy(x.l, in, FQB);
    const bool inf = (x.l[L - 1] >> 62) & 1, positive = (x.l[L - 1] >> 63) & 1;
    x.l[L - 1] &= ~(3ULL << 62);
    if (inf && positive) return false;                     // SWFlags::from_u8: both flag bits set is no valid flag
    if (host::ge<L>(x.l, FQ.p)) return false;              // the x field must be canonical, also under the infinity flag
    if (inf) { out->inf = true; out->x = qzero(); out->y = qzero(); return true; }
    Fq xm = q_from_canon(x.l);
    Fq rhs = qadd(qmul(qmul(xm, xm), xm), q_small(host::G1_COEFF_B));
    Fq y;
    if (!fq_sqrt(rhs, &y)) return false;                   // x is not on the curve
    Fq yc = q_canonical(y), nyc;
    host::sub_raw<L>(nyc.l, FQ.p, yc.l);
    if (host::is_zero(yc)) memset(nyc.l, 0, sizeof nyc.l);
    bool gt = false;
    for (int i = L - 1; i >= 0; --i) if (yc.l[i] != nyc.l[i]) { gt = yc.l[i] > nyc.l[i]; break; }
    out->inf = false;
    out->x = xm;
    out->y = gt == positive ? y : qneg(y);
    return true;
}

constexpr int G2_W = 4 * host::FQ_L;          // 64-bit words of an affine G2 point at the C boundary: x.c0 x.c1 y.c0 y.c1 (16 / 24)
G2 g2_from(const uint64_t *v /* G2_W words */) {
    constexpr int L = host::FQ_L;
    G2 p;
    p.x = {fq_from(v), fq_from(v + L)};
    p.y = {fq_from(v + 2 * L), fq_from(v + 3 * L)};
    p.inf = f2_is_zero(p.x) && f2_is_zero(p.y);
    return p;
}
bool fq_canonical_mont(const Fq &m) { return !host::ge<host::FQ_L>(m.l, FQ.p); }

bool pairing_product_is_one(const Pt *g1, const G2 *g2, size_t k) {
    F12 f = f12_one();
    for (size_t i = 0; i < k; ++i) f = f12_mul(f, miller_loop(g2[i], g1[i].x, g1[i].y, g1[i].inf));
    return f12_eq(final_exponentiation(f), f12_one());
}

// util.rs:185-195 compute_lagrange_evaluation(n, point, zh_eval, tau) = zh * point / (n (tau - point))
Fe lagrange(uint64_t n, const Fe &point, const Fe &zh, const Fe &tau) {
    Fe den = fmul(host::from_u64(n, FR), fsub(tau, point));
    return fmul(fmul(zh, point), host::inv(den, FR));
}

// SonicKZG10::check for one query point (no degree bounds, no hiding)
bool kzg_check(const Pt *commits, const Fe *values, size_t k, const Fe &point, const Pt &w, const Fe &eta, const G2 &h, const G2 &beta_h) {
    X1 c = x1_zero();
    Fe chal = FR_ONE(), v = {{0, 0, 0, 0}};
    for (size_t i = 0; i < k; ++i) {
        c = x1_add(c, x1_mul(chal, commits[i]));
        v = fadd(v, fmul(chal, values[i]));
        chal = fmul(chal, eta);
    }
    Pt gen;                                               // the G1 generator: powers_of_g[0] of the committer key
    gen.inf = false; gen.x = fq_from(host::G1_GEN_X); gen.y = fq_from(host::G1_GEN_Y);
    X1 a = x1_add(x1_add(c, x1_mul(fneg(v), gen)), x1_mul(point, w));
    Pt g1s[2] = {x1_affine(a), pt_neg(w)};
    G2 g2s[2] = {h, beta_h};
    return pairing_product_is_one(g1s, g2s, 2);
}

}  // namespace

extern "C" {

// k * Q on the twist E'(Fq2) (BN254: y^2 = x^3 + 3 / (9 + i)), affine double-and-add (one Fq2 inversion per step; setup-time only).
// KZG10::setup's beta_h = beta * h for a synthetic SRS whose trapdoor is known (ark-poly-commit 0.3 kzg10::setup).
int zkb_g2_mul(const uint64_t *g2_xy, const uint64_t scalar_canonical[4], uint64_t *out_xy) {
    if (!g2_xy || !scalar_canonical || !out_xy) return ZKB_ERR_INVALID;
    const G2 q = g2_from(g2_xy);
    if (!g2_on_curve(q)) return ZKB_ERR_INVALID;
    G2 acc;
    acc.inf = true;
    acc.x = {qzero(), qzero()};
    acc.y = acc.x;
    auto add = [](const G2 &a, const G2 &b) -> G2 {
        if (a.inf) return b;
        if (b.inf) return a;
        F2 lam;
        if (f2_eq(a.x, b.x)) {
            if (!f2_eq(a.y, b.y) || f2_is_zero(a.y)) { G2 z; z.inf = true; z.x = {qzero(), qzero()}; z.y = z.x; return z; }
            const F2 three = {q_small(3), qzero()};
            lam = f2_mul(f2_mul(three, f2_mul(a.x, a.x)), f2_inv(f2_add(a.y, a.y)));
        } else {
            lam = f2_mul(f2_sub(b.y, a.y), f2_inv(f2_sub(b.x, a.x)));
        }
        G2 r;
        r.inf = false;
        r.x = f2_sub(f2_sub(f2_mul(lam, lam), a.x), b.x);
        r.y = f2_sub(f2_mul(lam, f2_sub(a.x, r.x)), a.y);
        return r;
    };
    for (int bit = 255; bit >= 0; --bit) {
        acc = add(acc, acc);
        if ((scalar_canonical[bit >> 6] >> (bit & 63)) & 1) acc = add(acc, q);
    }
    if (acc.inf) { memset(out_xy, 0, 8 * G2_W); return ZKB_OK; }
    constexpr int L = host::FQ_L;
    memcpy(out_xy, acc.x.c0.l, FQB); memcpy(out_xy + L, acc.x.c1.l, FQB);
    memcpy(out_xy + 2 * L, acc.y.c0.l, FQB); memcpy(out_xy + 3 * L, acc.y.c1.l, FQB);
    return ZKB_OK;
}

int zkb_pairing(const uint64_t *g1_xy, const uint64_t *g2_xy, uint64_t *out_canonical /* 12 x fq_words */) {
    if (!g1_xy || !g2_xy || !out_canonical) return ZKB_ERR_INVALID;
    Pt p;
    p.x = fq_from(g1_xy); p.y = fq_from(g1_xy + AFF_W / 2);
    p.inf = host::is_zero(p.x) && host::is_zero(p.y);
    G2 q = g2_from(g2_xy);
    if (!fq_canonical_mont(p.x) || !fq_canonical_mont(p.y) || !g1_on_curve(p) || !g2_on_curve(q)) return ZKB_ERR_INVALID;
    const F12 ml = miller_loop(q, p.x, p.y, p.inf);
    F12 e = final_exponentiation(ml);
    if (!f12_eq(e, final_exponentiation_plain(ml))) return ZKB_ERR_INVALID;      // the split exponentiation against its definition
    for (int i = 0; i < 12; ++i) {
        Fq c = q_canonical(e.c[i]);
        memcpy(out_canonical + host::FQ_L * i, c.l, FQB);
    }
    return ZKB_OK;
}

int zkb_pairing_product_is_one(const uint64_t *g1_xy, const uint64_t *g2_xy, size_t count, int *is_one) {
    if ((!g1_xy || !g2_xy) && count) return ZKB_ERR_INVALID;
    if (!is_one) return ZKB_ERR_INVALID;
    if (count > (1u << 20)) return ZKB_ERR_INVALID;               // a pairing product has a handful of factors
    std::vector<Pt> ps(count);
    std::vector<G2> qs(count);
    for (size_t i = 0; i < count; ++i) {
        ps[i].x = fq_from(g1_xy + AFF_W * i); ps[i].y = fq_from(g1_xy + AFF_W * i + AFF_W / 2);
        ps[i].inf = host::is_zero(ps[i].x) && host::is_zero(ps[i].y);
        qs[i] = g2_from(g2_xy + G2_W * i);
        if (!g1_on_curve(ps[i]) || !g2_on_curve(qs[i])) return ZKB_ERR_INVALID;
    }
    *is_one = pairing_product_is_one(ps.data(), qs.data(), count) ? 1 : 0;
    return ZKB_OK;
}

int zkb_plonk_verify(size_t n, const uint64_t *pi_roots_mont, size_t n_pi, const uint64_t *vk_xy, const int vk_inf[10],
                     const uint64_t *pub_inputs_mont, const uint8_t *proof, const uint64_t *g2_h, const uint64_t *g2_beta_h,
                     int transcript_kind) {
    if (!vk_xy || !proof || !g2_h || !g2_beta_h || ((!pi_roots_mont || !pub_inputs_mont) && n_pi)) return ZKB_ERR_INVALID;
    if (n < 2 || (n & (n - 1)) || n > ((size_t)1 << 28) || (transcript_kind != 0 && transcript_kind != 1)) return ZKB_ERR_INVALID;
    if (transcript_kind == 1 && ZKB_CURVE != ZKB_CURVE_BN254) return ZKB_ERR_UNSUPPORTED;          // EthereumTranscript is bound to Bn254 upstream
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < n) ++log_n;
    // ---- inputs
    Pt V[10];                                             // q_m q_l q_r q_o q_c sigma1 sigma2 sigma3 q_lookup q_table
    for (int k = 0; k < 10; ++k) {
        V[k].x = fq_from(vk_xy + AFF_W * k); V[k].y = fq_from(vk_xy + AFF_W * k + AFF_W / 2);
        V[k].inf = (vk_inf && vk_inf[k]) || (host::is_zero(V[k].x) && host::is_zero(V[k].y));
        if (!V[k].inf && (!fq_canonical_mont(V[k].x) || !fq_canonical_mont(V[k].y) || !g1_on_curve(V[k]))) return ZKB_ERR_INVALID;
    }
    const G2 h = g2_from(g2_h), beta_h = g2_from(g2_beta_h);
    if (!g2_on_curve(h) || !g2_on_curve(beta_h)) return ZKB_ERR_INVALID;
    Pt C[11], aw, saw;                                    // a b c t h1 h2 z1 z2 q_lo q_mid q_hi  (proof.rs:112-154)
    for (int k = 0; k < 11; ++k) if (!g1_decompress(proof + FQB * k, &C[k])) return ZKB_ERR_INVALID;
    if (!g1_decompress(proof + 11 * FQB, &aw) || !g1_decompress(proof + 12 * FQB + 1, &saw) || proof[12 * FQB] != 0 || proof[13 * FQB + 1] != 0)
        return ZKB_ERR_INVALID;                                   // kzg10::Proof { w, random_v: None } twice
    Fe E[12];                                             // a b c sigma1 sigma2 z1_next q_lookup t t_next z2_next h1_next h2
    {
        Fe r2;
        memcpy(r2.l, FR.r2, 32);
        for (int k = 0; k < 12; ++k) {
            Fe c;
            memcpy(c.l, proof + 13 * FQB + 2 + 32 * k, 32);
            if (host::ge(c.l, FR.p)) return ZKB_ERR_INVALID;
            E[k] = host::mul(c, r2, FR);
        }
    }
    enum { E_A, E_B, E_C, E_S1, E_S2, E_Z1N, E_QLK, E_T, E_TN, E_Z2N, E_H1N, E_H2 };
    enum { C_A, C_B, C_C, C_T, C_H1, C_H2, C_Z1, C_Z2, C_QLO, C_QMID, C_QHI };
    enum { V_QM, V_QL, V_QR, V_QO, V_QC, V_S1, V_S2, V_S3, V_QLK, V_QT };
    const Fe *pi = (const Fe *)pub_inputs_mont, *roots = (const Fe *)pi_roots_mont;
    // ---- transcript (keys/mod.rs:260-275, proof.rs:300-430)
    Transcript tr("ZKT Plonk", transcript_kind);
    tr.append_u64("circuit_size", (uint64_t)n);
    {
        const char *labels[10] = {"q_m_commit", "q_l_commit", "q_r_commit", "q_o_commit", "q_c_commit", "sigma1_commit",
                                  "sigma2_commit", "sigma3_commit", "q_lookup_commit", "q_table_commit"};
        for (int k = 0; k < 10; ++k) tr.append_commitment(labels[k], V[k]);
    }
    tr.append_scalars("pi", pi, n_pi);
    tr.append_commitment("a_commit", C[C_A]);
    tr.append_commitment("b_commit", C[C_B]);
    tr.append_commitment("c_commit", C[C_C]);
    tr.append_commitment("t_commit", C[C_T]);
    tr.append_commitment("h1_commit", C[C_H1]);
    tr.append_commitment("h2_commit", C[C_H2]);
    const Fe beta = tr.challenge_scalar("beta"), gamma = tr.challenge_scalar("gamma");
    const Fe delta = tr.challenge_scalar("delta"), epsilon = tr.challenge_scalar("epsilon");
    tr.append_commitment("z1_commit", C[C_Z1]);
    tr.append_commitment("z2_commit", C[C_Z2]);
    const Fe alpha = tr.challenge_scalar("alpha");
    tr.append_commitment("q_lo_commit", C[C_QLO]);
    tr.append_commitment("q_mid_commit", C[C_QMID]);
    tr.append_commitment("q_hi_commit", C[C_QHI]);
    const Fe xi = tr.challenge_scalar("xi");
    // ---- scalars
    const Fe one = FR_ONE();
    const Fe zh = fsub(host::pow_u64(xi, (uint64_t)n, FR), one);
    const Fe l1 = lagrange(n, one, zh, xi);
    const Fe al2 = fmul(alpha, alpha), al3 = fmul(al2, alpha), al4 = fmul(al2, al2), al5 = fmul(al4, alpha);
    const Fe opd = fadd(one, delta), eopd = fmul(epsilon, opd);
    // compute_r0 (proof.rs:163-217)
    Fe part1 = {{0, 0, 0, 0}};
    for (size_t k = 0; k < n_pi; ++k) part1 = fsub(part1, fmul(lagrange(n, roots[k], zh, xi), pi[k]));
    Fe part2 = fmul(fmul(fmul(fmul(alpha, E[E_Z1N]), fadd(fadd(E[E_A], fmul(beta, E[E_S1])), gamma)),
                         fadd(fadd(E[E_B], fmul(beta, E[E_S2])), gamma)), fadd(E[E_C], gamma));
    Fe part3 = fmul(l1, al2);
    Fe part4 = fmul(fmul(fmul(al3, E[E_Z2N]), fadd(eopd, fmul(delta, E[E_H2]))), fadd(fadd(eopd, E[E_H2]), fmul(delta, E[E_H1N])));
    Fe part5 = fmul(l1, al4);
    const Fe r0 = fadd(fadd(fadd(fadd(part1, part2), part3), part4), part5);
    // compute_linearization_commitment (proof.rs:220-282)
    const Fe bz = fmul(beta, xi), k1 = host::from_u64(7, FR), k2 = host::from_u64(13, FR);
    Fe sc[13];
    const Pt *pts[13] = {&V[V_QM], &V[V_QL], &V[V_QR], &V[V_QO], &V[V_QC], &C[C_Z1], &V[V_S3], &C[C_Z2], &C[C_H1], &V[V_QT],
                         &C[C_QLO], &C[C_QMID], &C[C_QHI]};
    sc[0] = fmul(E[E_A], E[E_B]); sc[1] = E[E_A]; sc[2] = E[E_B]; sc[3] = E[E_C]; sc[4] = one;
    sc[5] = fadd(fmul(fmul(fmul(alpha, fadd(fadd(bz, E[E_A]), gamma)), fadd(fadd(fmul(bz, k1), E[E_B]), gamma)),
                      fadd(fadd(fmul(bz, k2), E[E_C]), gamma)), fmul(l1, al2));
    sc[6] = fneg(fmul(fmul(fmul(fmul(alpha, beta), E[E_Z1N]), fadd(fadd(fmul(beta, E[E_S1]), E[E_A]), gamma)),
                      fadd(fadd(fmul(beta, E[E_S2]), E[E_B]), gamma)));
    sc[7] = fadd(fmul(fmul(fmul(al3, opd), fadd(epsilon, fmul(E[E_QLK], E[E_C]))), fadd(fadd(eopd, E[E_T]), fmul(delta, E[E_TN]))),
                 fmul(al4, l1));
    sc[8] = fneg(fmul(fmul(al3, E[E_Z2N]), fadd(fadd(eopd, E[E_H2]), fmul(delta, E[E_H1N]))));
    sc[9] = fmul(al5, E[E_T]);
    const Fe xn2 = fmul(fmul(fadd(zh, one), xi), xi);     // xi^(n + 2)
    sc[10] = fneg(zh); sc[11] = fneg(fmul(zh, xn2)); sc[12] = fneg(fmul(fmul(zh, xn2), xn2));
    X1 rc = x1_zero();
    for (int k = 0; k < 13; ++k) rc = x1_add(rc, x1_mul(sc[k], *pts[k]));
    const Pt r_commit = x1_affine(rc);
    {
        const char *labels[12] = {"a_eval", "b_eval", "c_eval", "sigma1_eval", "sigma2_eval", "z1_next_eval", "q_lookup_eval", "t_eval",
                                  "t_next_eval", "z2_next_eval", "h1_next_eval", "h2_eval"};
        for (int k = 0; k < 12; ++k) tr.append_scalar(labels[k], E[k]);
    }
    const Fe eta = tr.challenge_scalar("eta");
    // ---- the two openings (proof.rs:441-502)
    {
        const Pt cs[9] = {r_commit, C[C_A], C[C_B], C[C_C], V[V_S1], V[V_S2], V[V_QLK], C[C_T], C[C_H2]};
        const Fe vs[9] = {r0, E[E_A], E[E_B], E[E_C], E[E_S1], E[E_S2], E[E_QLK], E[E_T], E[E_H2]};
        if (!kzg_check(cs, vs, 9, xi, aw, eta, h, beta_h)) return 1;
    }
    {
        const Pt cs[4] = {C[C_Z1], C[C_Z2], C[C_T], C[C_H1]};
        const Fe vs[4] = {E[E_Z1N], E[E_Z2N], E[E_TN], E[E_H1N]};
        if (!kzg_check(cs, vs, 4, fmul(xi, host::fr_root_of_unity(log_n)), saw, eta, h, beta_h)) return 2;
    }
    return 0;
}

}  // extern "C"
