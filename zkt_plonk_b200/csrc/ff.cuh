// ff.cuh -- BN254 Fr / Fq arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces ark-ff 0.3 `Fp256<FrParameters/FqParameters>` (crates.io dependency of the reference,
// plonk-core/Cargo.toml:20) on the device.  In-memory layout is identical to arkworks' (4 x u64 little
// endian == 8 x u32 little endian), so buffers cross the C ABI without conversion.
//
// Multiplication is a CIOS Montgomery product over two interleaved accumulators ("even" limbs and
// "odd" limbs shifted by 32 bits), so that every 32x32->64 partial product lands on an aligned register
// pair and each row is ONE carry chain of mad.lo.cc / madc.hi.cc -- ptxas turns each lo/hi pair into one
// IMAD.WIDE.U32(.X) on the fma pipe.  Every carry chain lives inside a single asm statement (the CC
// flag is invisible to the compiler, so chains are never split across statements).
#pragma once
#include <stdint.h>

namespace zkb {

struct fe_t { uint32_t v[8]; };

// ------------------------------------------------------------------ field parameters
struct FrP {
    static __host__ __device__ __forceinline__ constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) {   // R mod r
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) {    // R^2 mod r
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;                                 // -r^-1 mod 2^32
};

struct FqP {
    static __host__ __device__ __forceinline__ constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) {   // R mod q
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) {    // R^2 mod q
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;                                 // -q^-1 mod 2^32
};

#ifdef __CUDACC__

// ------------------------------------------------------------------ carry-chain building blocks
// acc[0..7] = {m0,m1,m2,m3} * b laid out as four aligned 64-bit products (no carries involved).
__device__ __forceinline__ void row_mul(uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
    asm("mul.lo.u32 %0, %8, %12;\n\t"  "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"  "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t" "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t" "mul.hi.u32 %7, %11, %12;"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
}

// lo += x (carry c);  acc[0..7] += {m0..m3}*b + c  -- the carry out of acc[7] is provably zero here.
__device__ __forceinline__ void row_mad_cin(uint32_t &lo, uint32_t x, uint32_t (&acc)[8],
                                            uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t" "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t" "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t" "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t" "madc.hi.u32 %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
          "+r"(lo)
        : "r"(x), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
}

// acc[0..7] += {m0..m3}*b (no carry in, carry out of acc[7] dropped: provably zero).
__device__ __forceinline__ void row_mad(uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"   "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"  "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t" "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t" "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
}

// acc[0..7] += {m0..m3}*b ; top += carry out of acc[7].
__device__ __forceinline__ void row_mad_cout(uint32_t (&acc)[8], uint32_t &top,
                                             uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"    "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"  "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"  "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"  "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
          "+r"(top)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
}

// r = a + b over 8 limbs, returns carry out
__device__ __forceinline__ uint32_t add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"  "addc.cc.u32 %1, %10, %18;\n\t" "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t" "addc.cc.u32 %4, %13, %21;\n\t" "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t" "addc.cc.u32 %7, %16, %24;\n\t" "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}

// r = a - b over 8 limbs, returns borrow (0 or 0xffffffff)
__device__ __forceinline__ uint32_t sub8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"  "subc.cc.u32 %1, %10, %18;\n\t" "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t" "subc.cc.u32 %4, %13, %21;\n\t" "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t" "subc.cc.u32 %7, %16, %24;\n\t" "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}

template <class P>
__device__ __forceinline__ void load_mod(uint32_t (&m)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = P::mod(i);
}

// if t >= p then t -= p   (t < 2p)
template <class P>
__device__ __forceinline__ void reduce_once(uint32_t (&t)[8]) {
    uint32_t m[8], d[8];
    load_mod<P>(m);
    uint32_t borrow = sub8(d, t, m);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = borrow ? t[i] : d[i];
}

// ------------------------------------------------------------------ field operations (inputs and outputs in [0,p))
template <class P>
__device__ __forceinline__ fe_t fadd(const fe_t &a, const fe_t &b) {
    fe_t r;
    add8(r.v, a.v, b.v);          // p < 2^254: the sum never carries out of 256 bits
    reduce_once<P>(r.v);
    return r;
}

template <class P>
__device__ __forceinline__ fe_t fsub(const fe_t &a, const fe_t &b) {
    fe_t r;
    uint32_t m[8], s[8];
    uint32_t borrow = sub8(r.v, a.v, b.v);
    load_mod<P>(m);
    add8(s, r.v, m);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = borrow ? s[i] : r.v[i];
    return r;
}

template <class P>
__device__ __forceinline__ fe_t fdbl(const fe_t &a) { return fadd<P>(a, a); }

template <class P>
__device__ __forceinline__ bool fis_zero(const fe_t &a) {
    return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}

__device__ __forceinline__ bool feq(const fe_t &a, const fe_t &b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.v[i] ^ b.v[i];
    return d == 0;
}

template <class P>
__device__ __forceinline__ fe_t fneg(const fe_t &a) {
    fe_t r;
    uint32_t m[8];
    load_mod<P>(m);
    sub8(r.v, m, a.v);
    bool z = fis_zero<P>(a);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = z ? 0u : r.v[i];
    return r;
}

template <class P>
__device__ __forceinline__ fe_t fzero() {
    fe_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0;
    return r;
}

template <class P>
__device__ __forceinline__ fe_t fone() {
    fe_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = P::one(i);
    return r;
}

// Montgomery product a*b*R^-1 mod p.
template <class P>
__device__ __forceinline__ fe_t fmul(const fe_t &a, const fe_t &b) {
    uint32_t E[8], O[8], x = 0;
    // row 0: even limbs of a -> E (positions 0..7), odd limbs -> O (positions 1..8)
    row_mul(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    row_mul(O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i) {
            // T += x + a * b[i]   (x is the limb left over at position 0 by the previous shift)
            row_mad_cin(E[0], x, O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            row_mad_cout(E, O[7], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
        }
        // T += m * p with m chosen so that limb 0 cancels
        uint32_t m = E[0] * P::INV;
        row_mad(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        row_mad_cout(E, O[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
        // T >>= 32: O becomes the even accumulator, E[2..7] the odd one, E[1] is carried over as x
        x = E[1];
        uint32_t t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = O[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) O[k] = E[k + 2];
        O[6] = 0; O[7] = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) E[k] = t[k];
    }
    // merge: r = E + (O << 32) + x  (< 2p < 2^255, so O[7] == 0 and no carry out)
    fe_t r;
    asm("add.cc.u32 %0, %8, %16;\n\t"  "addc.cc.u32 %1, %9, %17;\n\t"  "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t" "addc.cc.u32 %4, %12, %20;\n\t" "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t" "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(x), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    reduce_once<P>(r.v);
    return r;
}

// ------------------------------------------------------------------ dedicated squaring
// Triangular rows: with d = 2a (a < 2^254, so d fits 8 limbs), row i adds
//     a_i * [ a_i, d_{i+1} & ~1, d_{i+2}, ..., d_7 ]   at relative limbs i .. 7
// (d_{i+1} & ~1 = a_{i+1} << 1: the bit that 2a carries in from a_i belongs to a row already done), i.e. every
// off-diagonal product once, doubled, plus the diagonal: 36 wide products instead of 64 in the a*a half; the
// word-serial reduction (64 + 8) is unchanged.  Same even / odd accumulator layout as fmul, so each row is again
// one carry chain per accumulator; a chain that starts above pair 0 first ripples the incoming carry through the
// limbs below it (ALU pipe, which the product leaves idle).  tools/sqr/model.py replays these chains limb by limb
// and asserts that every dropped carry is zero.
//
// K0 = first pair of the chain.  Operands as in row_mad_cin / row_mad_cout; unused multiplicands are ignored.
template <int K0>
__device__ __forceinline__ void tri_mad_cin(uint32_t &lo, uint32_t x, uint32_t (&acc)[8],
                                            uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
    if (K0 == 0) {
        row_mad_cin(lo, x, acc, m0, m1, m2, m3, b);
    } else if (K0 == 1) {
        asm("add.cc.u32 %8, %8, %9;\n\t"
            "addc.cc.u32 %0, %0, 0;\n\t"            "addc.cc.u32 %1, %1, 0;\n\t"
            "madc.lo.cc.u32 %2, %11, %14, %2;\n\t" "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
            "madc.lo.cc.u32 %4, %12, %14, %4;\n\t" "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
            "madc.lo.cc.u32 %6, %13, %14, %6;\n\t" "madc.hi.u32 %7, %13, %14, %7;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(lo)
            : "r"(x), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
    } else if (K0 == 2) {
        asm("add.cc.u32 %8, %8, %9;\n\t"
            "addc.cc.u32 %0, %0, 0;\n\t"            "addc.cc.u32 %1, %1, 0;\n\t"
            "addc.cc.u32 %2, %2, 0;\n\t"            "addc.cc.u32 %3, %3, 0;\n\t"
            "madc.lo.cc.u32 %4, %12, %14, %4;\n\t" "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
            "madc.lo.cc.u32 %6, %13, %14, %6;\n\t" "madc.hi.u32 %7, %13, %14, %7;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(lo)
            : "r"(x), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
    } else {
        asm("add.cc.u32 %8, %8, %9;\n\t"
            "addc.cc.u32 %0, %0, 0;\n\t"            "addc.cc.u32 %1, %1, 0;\n\t"
            "addc.cc.u32 %2, %2, 0;\n\t"            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"            "addc.cc.u32 %5, %5, 0;\n\t"
            "madc.lo.cc.u32 %6, %13, %14, %6;\n\t" "madc.hi.u32 %7, %13, %14, %7;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(lo)
            : "r"(x), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
    }
}

template <int K0>
__device__ __forceinline__ void tri_mad_cout(uint32_t (&acc)[8], uint32_t &top,
                                             uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
    if (K0 == 0) {
        row_mad_cout(acc, top, m0, m1, m2, m3, b);
    } else if (K0 == 1) {
        asm("mad.lo.cc.u32 %2, %10, %13, %2;\n\t"   "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"  "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"  "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(top)
            : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
    } else if (K0 == 2) {
        asm("mad.lo.cc.u32 %4, %11, %13, %4;\n\t"   "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"  "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(top)
            : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
    } else if (K0 == 3) {
        asm("mad.lo.cc.u32 %6, %12, %13, %6;\n\t"   "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(top)
            : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
    }                                                  // K0 == 4: no even limb left in this row
}

// Montgomery reduction row + one-limb shift shared by fmul's and fsqr's rows.
template <class P>
__device__ __forceinline__ void mont_row_reduce(uint32_t (&E)[8], uint32_t (&O)[8], uint32_t &x) {
    uint32_t m = E[0] * P::INV;
    row_mad(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
    row_mad_cout(E, O[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
    x = E[1];
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = O[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) O[k] = E[k + 2];
    O[6] = 0; O[7] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) E[k] = t[k];
}

template <class P, int I>
__device__ __forceinline__ void sqr_row(uint32_t (&E)[8], uint32_t (&O)[8], uint32_t &x, const uint32_t (&a)[8], const uint32_t (&d)[8]) {
    // v[j] for j >= I: a_I, d_{I+1} & ~1, d_{I+2} ...; entries below I are never read by the chains
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < I ? 0u : (j == I ? a[I] : (j == I + 1 ? (d[j] & ~1u) : d[j]));
    tri_mad_cin<I / 2>(E[0], x, O, v[1], v[3], v[5], v[7], a[I]);
    tri_mad_cout<(I + 1) / 2>(E, O[7], v[0], v[2], v[4], v[6], a[I]);
    mont_row_reduce<P>(E, O, x);
}

template <class P>
__device__ __forceinline__ fe_t fsqr_tri(const fe_t &a) {
    uint32_t d[8], E[8], O[8], x = 0;
    d[0] = a.v[0] << 1;
#pragma unroll
    for (int j = 1; j < 8; ++j) d[j] = __funnelshift_l(a.v[j - 1], a.v[j], 1);
    row_mul(E, a.v[0], d[2], d[4], d[6], a.v[0]);
    row_mul(O, d[1] & ~1u, d[3], d[5], d[7], a.v[0]);
    mont_row_reduce<P>(E, O, x);
    sqr_row<P, 1>(E, O, x, a.v, d);
    sqr_row<P, 2>(E, O, x, a.v, d);
    sqr_row<P, 3>(E, O, x, a.v, d);
    sqr_row<P, 4>(E, O, x, a.v, d);
    sqr_row<P, 5>(E, O, x, a.v, d);
    sqr_row<P, 6>(E, O, x, a.v, d);
    sqr_row<P, 7>(E, O, x, a.v, d);
    fe_t r;
    asm("add.cc.u32 %0, %8, %16;\n\t"  "addc.cc.u32 %1, %9, %17;\n\t"  "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t" "addc.cc.u32 %4, %12, %20;\n\t" "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t" "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(x), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    reduce_once<P>(r.v);
    return r;
}

// Every fsqr is the triangular squaring (bucket accumulation: 2 of the 10 products of an XYZZ mixed addition;
// inversions: 254 of ~380 products).  Measured on B200 (profiles/r01f): bit-exact, bucket accumulation 2.7 % faster.
// -DZKB_DEDICATED_SQR=0 restores the plain product for A/B measurements.
#ifndef ZKB_DEDICATED_SQR
#define ZKB_DEDICATED_SQR 1
#endif
template <class P>
__device__ __forceinline__ fe_t fsqr(const fe_t &a) {
#if ZKB_DEDICATED_SQR
    return fsqr_tri<P>(a);
#else
    return fmul<P>(a, a);
#endif
}

// (a*b + c*d) * R^-1 mod p with ONE word-serial reduction: the rows of both products are added before each
// reduction row (200 instead of 272 multiply-adds).  Intermediate T < a + c + p < 3p < 2^256; the final value is
// below p/4 + p/4 + p, so one conditional subtraction still normalises it (tools/sqr/model.py replays the chains).
template <class P>
__device__ __forceinline__ fe_t fmadd2(const fe_t &a, const fe_t &b, const fe_t &c, const fe_t &d) {
    uint32_t E[8], O[8], x = 0;
    row_mul(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    row_mul(O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
    row_mad(O, c.v[1], c.v[3], c.v[5], c.v[7], d.v[0]);
    row_mad_cout(E, O[7], c.v[0], c.v[2], c.v[4], c.v[6], d.v[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i) {
            row_mad_cin(E[0], x, O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            row_mad_cout(E, O[7], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
            row_mad(O, c.v[1], c.v[3], c.v[5], c.v[7], d.v[i]);
            row_mad_cout(E, O[7], c.v[0], c.v[2], c.v[4], c.v[6], d.v[i]);
        }
        mont_row_reduce<P>(E, O, x);
    }
    fe_t r;
    asm("add.cc.u32 %0, %8, %16;\n\t"  "addc.cc.u32 %1, %9, %17;\n\t"  "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t" "addc.cc.u32 %4, %12, %20;\n\t" "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t" "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(x), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    reduce_once<P>(r.v);
    return r;
}

// sum_{k<N} a[k]*b[k] * R^-1 mod p for N <= 4 with one reduction (p < 0.19 * 2^256: intermediate T < 4p + p < 2^256,
// final T < 4 * 0.19 p + p < 2p; tools/sqr/model.py fmaddn asserts the dropped carries for N = 3, 4).
template <class P, int N>
__device__ __forceinline__ fe_t fmaddn(const fe_t (&a)[N], const fe_t (&b)[N]) {
    static_assert(N >= 1 && N <= 4, "at most four products share a reduction");
    uint32_t E[8], O[8], x = 0;
    row_mul(E, a[0].v[0], a[0].v[2], a[0].v[4], a[0].v[6], b[0].v[0]);
    row_mul(O, a[0].v[1], a[0].v[3], a[0].v[5], a[0].v[7], b[0].v[0]);
#pragma unroll
    for (int k = 1; k < N; ++k) {
        row_mad(O, a[k].v[1], a[k].v[3], a[k].v[5], a[k].v[7], b[k].v[0]);
        row_mad_cout(E, O[7], a[k].v[0], a[k].v[2], a[k].v[4], a[k].v[6], b[k].v[0]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i) {
            row_mad_cin(E[0], x, O, a[0].v[1], a[0].v[3], a[0].v[5], a[0].v[7], b[0].v[i]);
            row_mad_cout(E, O[7], a[0].v[0], a[0].v[2], a[0].v[4], a[0].v[6], b[0].v[i]);
#pragma unroll
            for (int k = 1; k < N; ++k) {
                row_mad(O, a[k].v[1], a[k].v[3], a[k].v[5], a[k].v[7], b[k].v[i]);
                row_mad_cout(E, O[7], a[k].v[0], a[k].v[2], a[k].v[4], a[k].v[6], b[k].v[i]);
            }
        }
        mont_row_reduce<P>(E, O, x);
    }
    fe_t r;
    asm("add.cc.u32 %0, %8, %16;\n\t"  "addc.cc.u32 %1, %9, %17;\n\t"  "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t" "addc.cc.u32 %4, %12, %20;\n\t" "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t" "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(x), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    reduce_once<P>(r.v);
    return r;
}

// a*b - c*d: the second product enters as (p - c)*d.  Used for Y3 = R (Q - X3) - Y1 PPP of every XYZZ addition and
// doubling (ec.cuh): 9 reductions instead of 10 per mixed addition.  Measured on B200 (profiles/r01g): bit-exact, bucket
// accumulation 7.4 % faster (ptxas also settles on 128 registers: 4 CTAs per SM instead of 3).
// -DZKB_FUSED_MADD2=0 restores the two separate products for A/B measurements.
#ifndef ZKB_FUSED_MADD2
#define ZKB_FUSED_MADD2 1
#endif
template <class P>
__device__ __forceinline__ fe_t fmsub2(const fe_t &a, const fe_t &b, const fe_t &c, const fe_t &d) {
#if ZKB_FUSED_MADD2
    return fmadd2<P>(a, b, fneg<P>(c), d);
#else
    return fsub<P>(fmul<P>(a, b), fmul<P>(c, d));
#endif
}

// a * small constant k (k <= 16) by an addition chain -- used for K1 = 7, K2 = 13, 3*X^2 ...
template <class P>
__device__ __forceinline__ fe_t fmul_small(const fe_t &a, unsigned k) {
    fe_t acc = fzero<P>(), cur = a;
    while (k) {
        if (k & 1) acc = fadd<P>(acc, cur);
        k >>= 1;
        if (k) cur = fdbl<P>(cur);
    }
    return acc;
}

// a^e for a 256-bit little-endian exponent held in 8 limbs (uniform across the warp in every caller)
template <class P>
__device__ __noinline__ fe_t fpow(const fe_t &a, const uint32_t *e) {
    fe_t acc = fone<P>();
    for (int i = 255; i >= 0; --i) {
        acc = fsqr<P>(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fmul<P>(acc, a);
    }
    return acc;
}

// a^(p-2)
template <class P>
__device__ __noinline__ fe_t finv(const fe_t &a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = P::mod(i);
    e[0] -= 2;                                    // both moduli end in ...01 / ...47: no borrow
    return fpow<P>(a, e);
}

template <class P>
__device__ __forceinline__ fe_t fto_mont(const fe_t &a) {
    fe_t r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.v[i] = P::r2(i);
    return fmul<P>(a, r2);
}

template <class P>
__device__ __forceinline__ fe_t ffrom_mont(const fe_t &a) {
    fe_t one = fzero<P>();
    one.v[0] = 1;
    return fmul<P>(a, one);
}

// ------------------------------------------------------------------ 32-byte loads / stores
__device__ __forceinline__ fe_t fload(const void *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fe_t r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ fe_t fload_ro(const void *p) {          // read-only path
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    fe_t r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ void fstore(void *p, const fe_t &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

#endif  // __CUDACC__
}  // namespace zkb
