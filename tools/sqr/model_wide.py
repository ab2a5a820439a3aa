#!/usr/bin/env python
"""Limb-level model of the 12-limb Montgomery product, of the dedicated squaring and of the fused a*b + c*d of csrc/ff.cuh (chains: csrc/ff_wide.cuh),
chain by chain: every carry the device code drops is asserted to be zero, results are compared with Python integers.
Run for the base fields of BLS12-381 and BLS12-377, random and extreme operands.

    python tools/sqr/model_wide.py [iterations]
"""
import random
import sys

M32 = 0xFFFFFFFF
N = 12
Q381 = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
Q377 = 0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001


def limbs(x):
    return [(x >> (32 * i)) & M32 for i in range(N)]


def chain(acc, ms, b, cin=0):
    """mad.lo.cc / madc.hi.cc over the six pairs of acc: acc += sum ms[k] * b << (64 k) + cin.  Returns the carry out."""
    c = cin
    for k in range(N // 2):
        prod = ms[k] * b
        t = acc[2 * k] + (prod & M32) + c
        acc[2 * k], c = t & M32, t >> 32
        t = acc[2 * k + 1] + (prod >> 32) + c
        acc[2 * k + 1], c = t & M32, t >> 32
    return c


def even(v):
    return v[0::2]


def odd(v):
    return v[1::2]


def add_row(E, O, x, v, w, first):
    """T += x + v * w: wrow_mad_cin<1> on O (with the carry of E[0] += x) and wrow_mad_cout<0> on E; plain chains when `first`
    (x is then already inside: the second product's row of the fused version)."""
    if not first:
        t = E[0] + x
        E[0], c = t & M32, t >> 32
    else:
        c = 0
    assert chain(O, odd(v), w, c) == 0, "wrow_mad_cin / wrow_mad dropped a carry"
    O[N - 1] += chain(E, even(v), w)
    assert O[N - 1] <= M32, "the carry into the top limb overflowed"


def reduce_shift(E, O, p, inv):
    m = (E[0] * inv) & M32
    pl = limbs(p)
    assert chain(O, odd(pl), m) == 0
    O[N - 1] += chain(E, even(pl), m)
    assert O[N - 1] <= M32 and E[0] == 0
    x = E[1]
    return list(O), E[2:N] + [0, 0], x


def fmadd(pairs, p):
    """sum of a*b over `pairs` with one reduction (one pair: fmul)"""
    inv = (-pow(p, -1, 1 << 32)) & M32
    E, O, x = [0] * N, [0] * N, 0
    for i in range(N):
        for j, (a, b) in enumerate(pairs):
            add_row(E, O, x, limbs(a), limbs(b)[i], first=(i == 0 or j > 0))
        E, O, x = reduce_shift(E, O, p, inv)
    assert O[N - 1] == 0
    r = sum(E[i] << (32 * i) for i in range(N)) + (sum(O[i] << (32 * i) for i in range(N - 1)) << 32) + x
    assert r < 2 * p and r < 1 << (32 * N), "one conditional subtraction must normalise the result"
    return r - p if r >= p else r


def tri_chain(acc, k0, ms, b, cin):
    """wtri_mad_cin / wtri_mad_cout: the pairs below k0 only ripple the carry, the products start at pair k0"""
    c = cin
    for k in range(N // 2):
        if k < k0:
            for j in (2 * k, 2 * k + 1):
                t = acc[j] + c
                acc[j], c = t & M32, t >> 32
        else:
            prod = ms[k] * b
            t = acc[2 * k] + (prod & M32) + c
            acc[2 * k], c = t & M32, t >> 32
            t = acc[2 * k + 1] + (prod >> 32) + c
            acc[2 * k + 1], c = t & M32, t >> 32
    return c


def fsqr(a, p):
    """the dedicated squaring of csrc/ff.cuh (12 limbs)"""
    inv = (-pow(p, -1, 1 << 32)) & M32
    al, dl = limbs(a), limbs(2 * a)
    assert 2 * a < 1 << (32 * N)
    E, O, x = [0] * N, [0] * N, 0
    for i in range(N):
        v = [0 if j < i else al[i] if j == i else (dl[j] & ~1 & M32) if j == i + 1 else dl[j] for j in range(N)]
        if i == 0:
            assert chain(O, odd(v), al[0]) == 0 and chain(E, even(v), al[0]) == 0          # wrow_mul: no carries at all
        else:
            t = E[0] + x
            E[0], c = t & M32, t >> 32
            assert tri_chain(O, i // 2, odd(v), al[i], c) == 0, "wtri_mad_cin dropped a carry"
            if (i + 1) // 2 < N // 2:
                O[N - 1] += tri_chain(E, (i + 1) // 2, even(v), al[i], 0)                   # wtri_mad_cout: no carry in
            assert O[N - 1] <= M32
        E, O, x = reduce_shift(E, O, p, inv)
    assert O[N - 1] == 0
    r = sum(E[i] << (32 * i) for i in range(N)) + (sum(O[i] << (32 * i) for i in range(N - 1)) << 32) + x
    assert r < 2 * p and r < 1 << (32 * N)
    return r - p if r >= p else r


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rnd = random.Random(1)
    R = 1 << (32 * N)
    for p in (Q381, Q377):
        rinv = pow(R, -1, p)
        extreme = [0, 1, p - 1, p - 2, (1 << 380) % p, M32, p >> 1]
        cases = [(a, b, c, d) for a in extreme for b in extreme for c in (0, p - 1) for d in (1, p - 1)]
        cases += [tuple(rnd.randrange(p) for _ in range(4)) for _ in range(iters)]
        for a, b, c, d in cases:
            assert fmadd([(a, b)], p) == a * b * rinv % p
            assert fmadd([(a, b), (c, d)], p) == (a * b + c * d) * rinv % p
            for v in (a, b, c, d):
                assert fsqr(v, p) == v * v * rinv % p
    print("ff_wide model: fmul, dedicated squaring and fused a*b + c*d exact, no dropped carry, for the base fields of BLS12-381 and BLS12-377")
