#!/usr/bin/env python
"""Limb-level model of csrc/ff.cuh's Montgomery product and of the dedicated squaring (fsqr), instruction chain by
instruction chain, to validate the carry handling on the CPU before spending GPU time: every carry the device code
drops is asserted to be zero here, and results are compared with a * b * 2^-256 mod p in Python integers.

    python tools/sqr/model.py [iterations]
"""
import random
import sys

M32 = 0xFFFFFFFF
FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def limbs(x):
    return [(x >> (32 * i)) & M32 for i in range(8)]


def chain_mad(acc, k0, ms, b, cin=0):
    """mad.lo.cc / madc.hi.cc chain over pairs k0..3 of acc (8 limbs): acc += sum ms[k] * b << (64 k) + cin << (64 k0).
    Returns the carry out of acc[7]."""
    c = cin
    for k in range(k0, 4):
        prod = ms[k] * b
        lo, hi = prod & M32, prod >> 32
        t = acc[2 * k] + lo + c
        acc[2 * k], c = t & M32, t >> 32
        t = acc[2 * k + 1] + hi + c
        acc[2 * k + 1], c = t & M32, t >> 32
    return c


def reduce_and_shift(E, O, p, inv):
    m = (E[0] * inv) & M32
    pl = limbs(p)
    assert chain_mad(O, 0, [pl[1], pl[3], pl[5], pl[7]], m) == 0          # row_mad: carry out dropped
    c = chain_mad(E, 0, [pl[0], pl[2], pl[4], pl[6]], m)                   # row_mad_cout
    O[7] += c
    assert O[7] <= M32 and E[0] == 0
    x = E[1]
    newE = list(O)
    newO = E[2:8] + [0, 0]
    return newE, newO, x


def finish(E, O, x, p):
    assert O[7] == 0
    r = sum(E[i] << (32 * i) for i in range(8)) + (sum(O[i] << (32 * i) for i in range(7)) << 32) + x
    assert r < 2 * p and r < 1 << 256
    return r - p if r >= p else r


def fmul(a, b, p):
    inv = (-pow(p, -1, 1 << 32)) & M32
    al, bl = limbs(a), limbs(b)
    E, O, x = [0] * 8, [0] * 8, 0
    for i in range(8):
        # row_mad_cin: E[0] += x, carry into the O chain
        t = E[0] + x
        E[0], c = t & M32, t >> 32
        assert chain_mad(O, 0, [al[1], al[3], al[5], al[7]], bl[i], c) == 0
        c = chain_mad(E, 0, [al[0], al[2], al[4], al[6]], bl[i])
        O[7] += c
        assert O[7] <= M32
        E, O, x = reduce_and_shift(E, O, p, inv)
    return finish(E, O, x, p)


def fmul2(a, b, c2, d, p):
    """(a*b + c2*d) * 2^-256 mod p with ONE word-serial reduction (ff.cuh fmadd2): both products' rows are added before
    each reduction row.  Intermediate T < a + c2 + p < 3p < 2^256, final T < p/4 + p/4 + p."""
    inv = (-pow(p, -1, 1 << 32)) & M32
    al, bl, cl, dl = limbs(a), limbs(b), limbs(c2), limbs(d)
    E, O, x = [0] * 8, [0] * 8, 0
    for i in range(8):
        t = E[0] + x
        E[0], c = t & M32, t >> 32
        assert chain_mad(O, 0, [al[1], al[3], al[5], al[7]], bl[i], c) == 0
        c = chain_mad(E, 0, [al[0], al[2], al[4], al[6]], bl[i])
        O[7] += c
        assert O[7] <= M32
        assert chain_mad(O, 0, [cl[1], cl[3], cl[5], cl[7]], dl[i]) == 0
        c = chain_mad(E, 0, [cl[0], cl[2], cl[4], cl[6]], dl[i])
        O[7] += c
        assert O[7] <= M32
        E, O, x = reduce_and_shift(E, O, p, inv)
    return finish(E, O, x, p)


def fmaddn(pairs, p):
    """sum of up to four products with one reduction (ff.cuh fmadd3 / fmadd4): intermediate T < 4p + p < 2^256 (p < 0.19 * 2^256),
    final T < 4 * p/4 * 0.76 + p < 2p."""
    inv = (-pow(p, -1, 1 << 32)) & M32
    E, O, x = [0] * 8, [0] * 8, 0
    ls = [(limbs(a), limbs(b)) for a, b in pairs]
    for i in range(8):
        t = E[0] + x
        E[0], c = t & M32, t >> 32
        for k, (al, bl) in enumerate(ls):
            assert chain_mad(O, 0, [al[1], al[3], al[5], al[7]], bl[i], c if k == 0 else 0) == 0
            c2 = chain_mad(E, 0, [al[0], al[2], al[4], al[6]], bl[i])
            O[7] += c2
            assert O[7] <= M32
        E, O, x = reduce_and_shift(E, O, p, inv)
    return finish(E, O, x, p)


def carry_ripple(acc, upto, c):
    """addc.cc chain through acc[0..upto) with incoming carry c; returns the carry that reaches acc[upto]."""
    for k in range(upto):
        t = acc[k] + c
        acc[k], c = t & M32, t >> 32
    return c


def fsqr(a, p):
    """Triangular rows: row i adds a_i * [a_i, (2a)_{i+1} & ~1, (2a)_{i+2}, ..., (2a)_7] at relative limbs i..7."""
    inv = (-pow(p, -1, 1 << 32)) & M32
    al = limbs(a)
    dl = limbs(2 * a)                                       # a < 2^254: no ninth limb
    E, O, x = [0] * 8, [0] * 8, 0
    for i in range(8):
        v = [0] * 8
        v[i] = al[i]
        if i + 1 < 8:
            v[i + 1] = dl[i + 1] & ~1 & M32
        for j in range(i + 2, 8):
            v[j] = dl[j]
        ke = (i + 1) // 2                                   # first even relative limb >= i is 2 * ke
        ko = i // 2                                         # first odd relative limb >= i is 2 * ko + 1
        # E[0] += x; the carry ripples through O[0 .. 2 ko) and enters the product chain at pair ko
        t = E[0] + x
        E[0], c = t & M32, t >> 32
        c = carry_ripple(O, 2 * ko, c)
        assert chain_mad(O, ko, [v[1], v[3], v[5], v[7]], al[i], c) == 0
        if ke < 4:
            c = chain_mad(E, ke, [v[0], v[2], v[4], v[6]], al[i])
            O[7] += c
            assert O[7] <= M32
        E, O, x = reduce_and_shift(E, O, p, inv)
    return finish(E, O, x, p)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rnd = random.Random(1)
    for p in (FR, FQ):
        rinv = pow(1 << 256, -1, p)
        edge = [0, 1, 2, p - 1, p - 2, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < p else p - 3, 0xFFFFFFFF, 0x80000000,
                int("80000000" * 8, 16) % p, int("7fffffff" * 8, 16) % p, int("ffffffff" * 8, 16) % p,
                sum(0x80000000 << (32 * i) for i in range(0, 8, 2)) % p, sum(0xFFFFFFFF << (32 * i) for i in range(1, 8, 2)) % p]
        vals = edge + [rnd.randrange(p) for _ in range(iters)]
        for a in vals:
            b = rnd.randrange(p)
            assert fmul(a, b, p) == a * b * rinv % p
            assert fsqr(a, p) == a * a * rinv % p, hex(a)
            c2 = rnd.choice(vals[:14]) if rnd.random() < 0.2 else rnd.randrange(p)
            d = rnd.choice(vals[:14]) if rnd.random() < 0.2 else rnd.randrange(p)
            assert fmul2(a, b, c2, d, p) == (a * b + c2 * d) * rinv % p
            for npairs in (3, 4):
                prs = [(rnd.choice(vals[:14]) if rnd.random() < 0.3 else rnd.randrange(p),
                        rnd.choice(vals[:14]) if rnd.random() < 0.3 else rnd.randrange(p)) for _ in range(npairs)]
                assert fmaddn(prs, p) == sum(u * v for u, v in prs) * rinv % p
        assert fmaddn([(p - 1, p - 1)] * 4, p) == 4 * (p - 1) * (p - 1) * rinv % p
        assert fmaddn([(p - 1, p - 1)] * 3 + [(0, 0)], p) == 3 * (p - 1) * (p - 1) * rinv % p
        for a in edge:                                       # all four operands at their extremes
            for b in edge:
                assert fmul2(a, b, p - 1, p - 1, p) == (a * b + (p - 1) * (p - 1)) * rinv % p
                assert fmul2(p - 1, p - 1, a, b, p) == (a * b + (p - 1) * (p - 1)) * rinv % p
    print("model ok:", 2 * (iters + 14), "products and squarings")


if __name__ == "__main__":
    main()
