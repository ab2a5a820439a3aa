// ff_kara.cuh -- Montgomery product with one level of Karatsuba on the a x b half (product first, reduction after).
//
// Every hot kernel of this library is bound by the SM's wide-multiplier pipe (IMAD.WIDE), while the ALU pipe
// (IADD3 / LOP3) is ~90 % idle.  ff.cuh's CIOS product spends 64 + 64 wide MACs (+ 8 IMAD).  Here the 8 x 8-limb product
// is three 4 x 4 products (48 wide MACs) whose recombination is carry-chain additions on the ALU pipe, followed by the
// same word-serial Montgomery reduction (64 wide MACs + 8 IMAD): 112 instead of 128 wide MACs.  Same inputs, same
// output (a * b * 2^-256 mod p in [0, p)) as fmul<P>.
#pragma once
#include "ff.cuh"

namespace zkb {
#ifdef __CUDACC__

// w[0..3] = {x0, x1} * b as two aligned 64-bit products
__device__ __forceinline__ void k_mul2(uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t x0, uint32_t x1, uint32_t b) {
    asm("mul.lo.u32 %0, %4, %6;\n\t" "mul.hi.u32 %1, %4, %6;\n\t" "mul.lo.u32 %2, %5, %6;\n\t" "mul.hi.u32 %3, %5, %6;"
        : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(x0), "r"(x1), "r"(b));
}
// w[0..3] += {x0, x1} * b, the carry rippling into u0, u1 (u1's carry out is dropped: the caller's bound)
__device__ __forceinline__ void k_mad2_r2(uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t &u0, uint32_t &u1,
                                          uint32_t x0, uint32_t x1, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %6, %8, %0;\n\t" "madc.hi.cc.u32 %1, %6, %8, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %8, %2;\n\t" "madc.hi.cc.u32 %3, %7, %8, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t" "addc.u32 %5, %5, 0;"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(u0), "+r"(u1) : "r"(x0), "r"(x1), "r"(b));
}
__device__ __forceinline__ void k_mad2_r3(uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t &u0, uint32_t &u1, uint32_t &u2,
                                          uint32_t x0, uint32_t x1, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %7, %9, %0;\n\t" "madc.hi.cc.u32 %1, %7, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %9, %2;\n\t" "madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t" "addc.cc.u32 %5, %5, 0;\n\t" "addc.u32 %6, %6, 0;"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(u0), "+r"(u1), "+r"(u2) : "r"(x0), "r"(x1), "r"(b));
}
__device__ __forceinline__ void k_mad2_r1(uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t &u0,
                                          uint32_t x0, uint32_t x1, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t" "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t" "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(u0) : "r"(x0), "r"(x1), "r"(b));
}

// r[0..7] = a[0..3] * b[0..3].  Products with i + j even accumulate in E (aligned pairs at limbs 0, 2, 4, 6), those with
// i + j odd in O (pairs at limbs 1, 3, 5; O[k] sits at limb k + 1), so that every 32 x 32 -> 64 product lands on an
// aligned register pair (one IMAD.WIDE each); r = E + (O << 32).
__device__ __forceinline__ void k_mul4x4(uint32_t (&r)[8], const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    uint32_t E[8], O[7];
    k_mul2(E[0], E[1], E[2], E[3], a[0], a[2], b[0]);                          // limbs 0..3
    k_mul2(E[4], E[5], E[6], E[7], a[1], a[3], b[3]);                          // limbs 4..7
    k_mad2_r2(E[2], E[3], E[4], E[5], E[6], E[7], a[1], a[3], b[1]);           // limbs 2..5
    k_mad2_r2(E[2], E[3], E[4], E[5], E[6], E[7], a[0], a[2], b[2]);           // limbs 2..5
    k_mul2(O[0], O[1], O[2], O[3], a[1], a[3], b[0]);                          // limbs 1..4
    O[4] = 0; O[5] = 0; O[6] = 0;
    k_mad2_r3(O[0], O[1], O[2], O[3], O[4], O[5], O[6], a[0], a[2], b[1]);     // limbs 1..4
    k_mad2_r1(O[2], O[3], O[4], O[5], O[6], a[1], a[3], b[2]);                 // limbs 3..6
    k_mad2_r1(O[2], O[3], O[4], O[5], O[6], a[0], a[2], b[3]);                 // limbs 3..6
    r[0] = E[0];
    asm("add.cc.u32 %0, %7, %14;\n\t"  "addc.cc.u32 %1, %8, %15;\n\t" "addc.cc.u32 %2, %9, %16;\n\t" "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t" "addc.cc.u32 %5, %12, %19;\n\t" "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
}

// s[0..3] = x[0..3] + y[0..3], returns the carry
__device__ __forceinline__ uint32_t k_add4(uint32_t (&s)[4], const uint32_t *x, const uint32_t *y) {
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\t" "addc.cc.u32 %1, %6, %10;\n\t" "addc.cc.u32 %2, %7, %11;\n\t" "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(c)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
    return c;
}

// T[0..15] = a * b (8 x 8 limbs) by one level of Karatsuba
__device__ __forceinline__ void k_mul8x8(uint32_t (&T)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t al[4] = {a[0], a[1], a[2], a[3]}, ah[4] = {a[4], a[5], a[6], a[7]};
    uint32_t bl[4] = {b[0], b[1], b[2], b[3]}, bh[4] = {b[4], b[5], b[6], b[7]};
    uint32_t sa[4], sb[4], P0[8], P2[8], M[9];
    const uint32_t ca = k_add4(sa, al, ah), cb = k_add4(sb, bl, bh);
    k_mul4x4(P0, al, bl);
    k_mul4x4(P2, ah, bh);
    {
        uint32_t P1[8];
        k_mul4x4(P1, sa, sb);
#pragma unroll
        for (int i = 0; i < 8; ++i) M[i] = P1[i];
    }
    // M = (sa + ca 2^128)(sb + cb 2^128) = P1 + (ca ? sb : 0) 2^128 + (cb ? sa : 0) 2^128 + (ca & cb) 2^256
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    M[8] = ca & cb;
    asm("add.cc.u32 %0, %0, %5;\n\t" "addc.cc.u32 %1, %1, %6;\n\t" "addc.cc.u32 %2, %2, %7;\n\t" "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
        : "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma));
    asm("add.cc.u32 %0, %0, %5;\n\t" "addc.cc.u32 %1, %1, %6;\n\t" "addc.cc.u32 %2, %2, %7;\n\t" "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
        : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb));
    // M -= P0 + P2  (the result a_lo b_hi + a_hi b_lo is non-negative and below 2^257)
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t *Q = pass ? P2 : P0;
        asm("sub.cc.u32 %0, %0, %9;\n\t"  "subc.cc.u32 %1, %1, %10;\n\t" "subc.cc.u32 %2, %2, %11;\n\t" "subc.cc.u32 %3, %3, %12;\n\t"
            "subc.cc.u32 %4, %4, %13;\n\t" "subc.cc.u32 %5, %5, %14;\n\t" "subc.cc.u32 %6, %6, %15;\n\t" "subc.cc.u32 %7, %7, %16;\n\t"
            "subc.u32 %8, %8, 0;"
            : "+r"(M[0]), "+r"(M[1]), "+r"(M[2]), "+r"(M[3]), "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
            : "r"(Q[0]), "r"(Q[1]), "r"(Q[2]), "r"(Q[3]), "r"(Q[4]), "r"(Q[5]), "r"(Q[6]), "r"(Q[7]));
    }
    // T = P0 + M 2^128 + P2 2^256
#pragma unroll
    for (int i = 0; i < 8; ++i) { T[i] = P0[i]; T[8 + i] = P2[i]; }
    asm("add.cc.u32 %0, %0, %12;\n\t"  "addc.cc.u32 %1, %1, %13;\n\t" "addc.cc.u32 %2, %2, %14;\n\t" "addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t" "addc.cc.u32 %5, %5, %17;\n\t" "addc.cc.u32 %6, %6, %18;\n\t" "addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t" "addc.cc.u32 %9, %9, 0;\n\t"   "addc.cc.u32 %10, %10, 0;\n\t" "addc.u32 %11, %11, 0;"
        : "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]),
          "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(M[0]), "r"(M[1]), "r"(M[2]), "r"(M[3]), "r"(M[4]), "r"(M[5]), "r"(M[6]), "r"(M[7]), "r"(M[8]));
}

// Montgomery product a * b * 2^-256 mod p: Karatsuba product, then the word-serial reduction of ff.cuh's CIOS loop run
// on the low half alone (its window never exceeds 8 limbs), then + the high half and one conditional subtraction.
template <class P>
__device__ __forceinline__ fe_t fmul_kara(const fe_t &a, const fe_t &b) {
    uint32_t T[16];
    k_mul8x8(T, a.v, b.v);
    uint32_t E[8], O[8], x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { E[k] = T[k]; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t m = (E[0] + x) * P::INV;
        row_mad_cin(E[0], x, O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);     // E[0] += x (carry into O), O += m * p_odd
        row_mad_cout(E, O[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);       // E += m * p_even: E[0] == 0 now
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
    fe_t r;
    asm("add.cc.u32 %0, %8, %16;\n\t"  "addc.cc.u32 %1, %9, %17;\n\t"  "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t" "addc.cc.u32 %4, %12, %20;\n\t" "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t" "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(x), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    uint32_t hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) hi[k] = T[8 + k];
    add8(r.v, r.v, hi);
    reduce_once<P>(r.v);
    return r;
}

#endif  // __CUDACC__
}  // namespace zkb
