"""BN254 (alt_bn128) optimal ate pairing in plain Python integers.  TEST INFRASTRUCTURE ONLY.

Gives the restated verifier (oracle/plonk_ref.py) the real `PC::check` of the reference -- SonicKZG10::check's
product of pairings e(C - v G + z W, H) * e(-W, beta H) == 1 (ark-poly-commit 0.3 sonic_pc / kzg10, reached from
plonk-core/src/proof_system/proof.rs:441-502) -- instead of the G1 shortcut through the synthetic SRS's trapdoor.
The pairing lives in ark-ec 0.3 / ark-bn254 0.3 (crates.io, un-vendored); this file restates the published
construction in its simplest form: Fq12 = Fq[w] / (w^12 - 18 w^6 + 82) (so that w^6 = 9 + i), G2 on the sextic
twist y^2 = x^3 + 3 / (9 + i) untwisted into E(Fq12), a Miller loop over 6x + 2 = 29793968203157093288 with affine
line functions, the two Frobenius-twisted additions, and the final exponentiation as a single power
(q^12 - 1) / r.  A pairing-product check is value-compatible with any correct implementation (the check is
"product == 1"), so no convention of arkworks' Miller loop can change its outcome.
Pinned by: the G2 generator (EIP-197 constants) lying on the twist and having order r; bilinearity in both
arguments; non-degeneracy (tests/test_pairing.py).
"""
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ATE_LOOP_COUNT = 29793968203157093288                      # 6x + 2, x = 4965661367192848881
# EIP-197 / alt_bn128 G2 generator: x = x0 + x1 i, y = y0 + y1 i
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


# ---- Fq2 = Fq[i] / (i^2 + 1), elements as (c0, c1)
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, -a[1] * d % Q)


B2 = f2_mul((3, 0), f2_inv((9, 1)))                        # twist coefficient 3 / (9 + i)


def g2_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_mul(y, y), f2_mul(f2_mul(x, x), x)) == B2


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        m = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        m = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(m, m), x1), x2)
    return (x3, f2_sub(f2_mul(m, f2_sub(x1, x3)), y1))


def g2_mul(k, p):
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def g2_neg(p):
    return None if p is None else (p[0], ((-p[1][0]) % Q, (-p[1][1]) % Q))


# ---- Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), elements as lists of 12 ints
def f12(c0=0):
    return [c0 % Q] + [0] * 11


F12_ONE = f12(1)


def f12_add(a, b):
    return [(x + y) % Q for x, y in zip(a, b)]


def f12_sub(a, b):
    return [(x - y) % Q for x, y in zip(a, b)]


def f12_scale(a, k):
    return [x * k % Q for x in a]


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):                             # w^12 = 18 w^6 - 82
        v = t[k]
        if v:
            t[k - 6] += 18 * v
            t[k - 12] -= 82 * v
    return [v % Q for v in t[:12]]


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_divmod_lead(a, b):
    """Quotient of polynomial division a / b over Fq (coefficient lists, low degree first)."""
    a = list(a)
    da, db = _poly_deg(a), _poly_deg(b)
    out = [0] * (da - db + 1)
    inv = pow(b[db], -1, Q)
    for i in range(da - db, -1, -1):
        c = a[db + i] * inv % Q
        out[i] = c
        if c:
            for j in range(db + 1):
                a[i + j] = (a[i + j] - c * b[j]) % Q
    return out


def f12_inv(a):
    """Extended Euclid in Fq[w] against the modulus polynomial."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, (-18) % Q, 0, 0, 0, 0, 0, 1]
    while _poly_deg(low):
        r = _poly_divmod_lead(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            if lm[i] or low[i]:
                for j in range(13 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
        nm, new = [x % Q for x in nm], [x % Q for x in new]
        lm, low, hm, high = nm, new, lm, low
    inv = pow(low[0], -1, Q)
    return [x * inv % Q for x in lm[:12]]


def f12_pow(a, e):
    acc, base = F12_ONE, a
    while e:
        if e & 1:
            acc = f12_mul(acc, base)
        base = f12_mul(base, base)
        e >>= 1
    return acc


# ---- curve over Fq12 (affine), untwist, line functions
def _e12_double(p):
    x, y = p
    m = f12_mul(f12_scale(f12_mul(x, x), 3), f12_inv(f12_scale(y, 2)))
    x3 = f12_sub(f12_mul(m, m), f12_scale(x, 2))
    return (x3, f12_sub(f12_mul(m, f12_sub(x, x3)), y)), m


def _e12_add(p, q):
    (x1, y1), (x2, y2) = p, q
    m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    x3 = f12_sub(f12_sub(f12_mul(m, m), x1), x2)
    return (x3, f12_sub(f12_mul(m, f12_sub(x1, x3)), y1)), m


def _line(p, m, t):
    """Line through p with slope m, evaluated at t."""
    return f12_sub(f12_mul(m, f12_sub(t[0], p[0])), f12_sub(t[1], p[1]))


def untwist(pt):
    """E'(Fq2) -> E(Fq12): i = w^6 - 9, then x / w^-2 ... i.e. (x w^2, y w^3) in the w-basis."""
    (x0, x1), (y0, y1) = pt
    nx, ny = [0] * 12, [0] * 12
    nx[2], nx[8] = (x0 - 9 * x1) % Q, x1                    # (a + b i) w^2 = (a - 9 b) w^2 + b w^8
    ny[3], ny[9] = (y0 - 9 * y1) % Q, y1
    return (nx, ny)


def miller_loop(q2, p1):
    """q2 on the twist (Fq2 coordinates), p1 on E(Fq) as (x, y) ints; identity on either side gives 1."""
    if q2 is None or p1 is None:
        return F12_ONE
    qq = untwist(q2)
    pp = (f12(p1[0]), f12(p1[1]))
    r, f = qq, F12_ONE
    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        r2, m = _e12_double(r)
        f = f12_mul(f12_mul(f, f), _line(r, m, pp))
        r = r2
        if (ATE_LOOP_COUNT >> i) & 1:
            r2, m = _e12_add(r, qq)
            f = f12_mul(f, _line(r, m, pp))
            r = r2
    q1 = (f12_pow(qq[0], Q), f12_pow(qq[1], Q))
    nq2 = (f12_pow(q1[0], Q), f12_scale(f12_pow(q1[1], Q), Q - 1))
    r2, m = _e12_add(r, q1)
    f = f12_mul(f, _line(r, m, pp))
    r = r2
    _, m = _e12_add(r, nq2)
    return f12_mul(f, _line(r, m, pp))


def final_exponentiation(f):
    return f12_pow(f, (Q ** 12 - 1) // R)


def pairing(q2, p1):
    return final_exponentiation(miller_loop(q2, p1))


def pairing_product_is_one(pairs):
    """pairs: [(G1 point, G2 point)]; product of pairings == 1 (what PairingEngine::product_of_pairings feeds)."""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(q2, p1))
    return final_exponentiation(f) == F12_ONE


# ---- the same Miller loop with the G2 arithmetic kept on the twist (Fq2) and sparse lines: the shape of the C++ verifier
# (csrc/verify.cu), which tests compare against the generic loop above value for value.
def f2_conj(a):
    return (a[0], (-a[1]) % Q)


def f2_pow(a, e):
    acc = (1, 0)
    while e:
        if e & 1:
            acc = f2_mul(acc, a)
        a = f2_mul(a, a)
        e >>= 1
    return acc


XI = (9, 1)
FROB_X = f2_pow(XI, (Q - 1) // 3)                 # (w^2)^(q-1): x-coordinate factor of the q-power Frobenius on the twist
FROB_Y = f2_pow(XI, (Q - 1) // 2)                 # (w^3)^(q-1)
FROB2_X = f2_pow(XI, (Q * Q - 1) // 3)            # q^2-power: an element of Fq (a cube root of unity); the y factor is -1


def embed(a):
    """Fq2 -> Fq12 coefficients (two of them): a0 + a1 i = (a0 - 9 a1) + a1 w^6."""
    return ((a[0] - 9 * a[1]) % Q, a[1])


def sparse_line(lam, xr, yr, p1):
    """Line through the untwisted (xr, yr) with slope embed(lam) w, evaluated at p1 = (xp, yp) in E(Fq):
    -yp + xp embed(lam) w + embed(yr - lam xr) w^3  ->  coefficients at w^0, w^1, w^7, w^3, w^9."""
    out = [0] * 12
    l0, l1 = embed(lam)
    c0, c1 = embed(f2_sub(yr, f2_mul(lam, xr)))
    out[0] = (-p1[1]) % Q
    out[1], out[7] = l0 * p1[0] % Q, l1 * p1[0] % Q
    out[3], out[9] = c0, c1
    return out


def miller_loop_twist(q2, p1):
    if q2 is None or p1 is None:
        return F12_ONE
    xq, yq = q2
    xr, yr = q2
    f = F12_ONE

    def add_step(f, xr, yr, x2, y2, square):
        if square:
            lam = f2_mul(f2_mul((3, 0), f2_mul(xr, xr)), f2_inv(f2_add(yr, yr)))
            x2, y2 = xr, yr
        else:
            lam = f2_mul(f2_sub(y2, yr), f2_inv(f2_sub(x2, xr)))
        line = sparse_line(lam, xr, yr, p1)
        f = f12_mul(f12_mul(f, f), line) if square else f12_mul(f, line)
        x3 = f2_sub(f2_sub(f2_mul(lam, lam), xr), x2)
        y3 = f2_sub(f2_mul(lam, f2_sub(xr, x3)), yr)
        return f, x3, y3

    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        f, xr, yr = add_step(f, xr, yr, None, None, True)
        if (ATE_LOOP_COUNT >> i) & 1:
            f, xr, yr = add_step(f, xr, yr, xq, yq, False)
    x1, y1 = f2_mul(f2_conj(xq), FROB_X), f2_mul(f2_conj(yq), FROB_Y)           # pi(Q)
    x2, y2 = f2_mul(xq, FROB2_X), yq                                            # -pi^2(Q): y * (-1) negated again
    f, xr, yr = add_step(f, xr, yr, x1, y1, False)
    lam = f2_mul(f2_sub(y2, yr), f2_inv(f2_sub(x2, xr)))
    return f12_mul(f, sparse_line(lam, xr, yr, p1))


# ---- final exponentiation split as (q^6 - 1) (q^2 + 1) ((q^4 - q^2 + 1) / r): the shape of csrc/verify.cu
ZETA = f2_pow(XI, (Q * Q - 1) // 6)               # w^(q^2) = ZETA * w; a sixth root of unity, it lies in Fq
HARD_EXP = (Q ** 4 - Q ** 2 + 1) // R


def f12_conj(a):
    """q^6-power Frobenius: w -> -w."""
    return [(-c) % Q if k & 1 else c for k, c in enumerate(a)]


def f12_frob2(a):
    """q^2-power Frobenius: coefficients are in Fq, w^k -> ZETA^k w^k."""
    out, z = [], 1
    for c in a:
        out.append(c * z % Q)
        z = z * ZETA[0] % Q
    return out


def final_exponentiation_split(f):
    t = f12_mul(f12_conj(f), f12_inv(f))
    t = f12_mul(f12_frob2(t), t)
    return f12_pow(t, HARD_EXP)
