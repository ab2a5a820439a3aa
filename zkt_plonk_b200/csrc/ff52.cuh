// ff52.cuh -- EXPERIMENTAL: Montgomery product on the FP64 pipe (52-bit limbs, R = 2^260).
//
// The integer product of ff.cuh saturates the fma pipe with IMAD.WIDE (257 issue slots per product); the FP64 pipe of
// B200 is as wide (measured: 16.8 T DFMA/s against 18.3 T IMAD/s, zkb_bench_int modes 3 / 0) and idle.  A 52 x 52-bit
// partial product is split exactly with two fused multiply-adds in round-toward-zero mode:
//     hi  = fma_rz(a, b, 2^104)                = 2^104 + floor(ab / 2^52) * 2^52        (ab < 2^104)
//     lo  = fma_rz(a, b, (2^104 + 2^52) - hi)  = 2^52 + (ab mod 2^52)
// so the mantissa fields of hi and lo ARE the two halves of the product; their bit patterns are accumulated as 64-bit
// integers (the exponent patterns are multiples of 2^52 and are pre-subtracted from the column accumulators).  Every
// floating-point operation is exact by construction; tools/dfma/proto.c checks the algorithm on the CPU with C `fma`.
// The reduction is digit-serial (five 52-bit digits of q), the result is a * b * 2^-260 mod p in [0, p).
#pragma once
#include <stdint.h>

#include "ff.cuh"

namespace zkb {

struct fe52_t { uint64_t v[5]; };            // value = sum v[i] 2^(52 i), v[i] < 2^52

#ifdef __CUDACC__
constexpr uint64_t MASK52 = (1ULL << 52) - 1;
constexpr uint64_t OFF_HI52 = 0x467ULL << 52, OFF_LO52 = 0x433ULL << 52;

template <class P>
struct P52 {                                  // 52-bit limbs of the modulus and -p^-1 mod 2^52, from the 32-bit limbs of ff.cuh
    static __host__ __device__ constexpr uint64_t word(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
    static __host__ __device__ constexpr uint64_t limb(int i) {
        return i == 0 ? word(0) & MASK52
             : i == 1 ? ((word(0) >> 52) | (word(1) << 12)) & MASK52
             : i == 2 ? ((word(1) >> 40) | (word(2) << 24)) & MASK52
             : i == 3 ? ((word(2) >> 28) | (word(3) << 36)) & MASK52
                      : word(3) >> 16;
    }
    static __host__ __device__ constexpr uint64_t pinv() {       // Newton iteration for p^-1 mod 2^64, negated, low 52 bits
        uint64_t p = word(0), x = p;                              // p * p = 1 mod 8 for odd p: 3 correct bits
        for (int k = 0; k < 6; ++k) x *= 2 - p * x;
        return (0 - x) & MASK52;
    }
};

__device__ __forceinline__ double d_from52(uint64_t x) {         // exact for x < 2^52
    return __longlong_as_double((long long)(x | OFF_LO52)) - 4503599627370496.0;
}

__device__ __forceinline__ constexpr uint64_t col_init52(int k) {   // minus the exponent patterns column k will receive
    int nlo = 0, nhi = 0;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) { if (i + j == k) ++nlo; if (i + j + 1 == k) ++nhi; }
    return 0 - 2 * ((uint64_t)nlo * OFF_LO52 + (uint64_t)nhi * OFF_HI52);
}

template <class P>
__device__ __forceinline__ fe52_t fmul52(const fe52_t &a, const fe52_t &b) {
    const double C1 = 20282409603651670423947251286016.0;        // 2^104
    const double C2 = 20282409603651674927546878656512.0;        // 2^104 + 2^52
    double ad[5], bd[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { ad[i] = d_from52(a.v[i]); bd[i] = d_from52(b.v[i]); }
    uint64_t col[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) col[k] = col_init52(k);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double hi = __fma_rz(ad[i], bd[j], C1);
            double lo = __fma_rz(ad[i], bd[j], C2 - hi);
            col[i + j + 1] += (uint64_t)__double_as_longlong(hi);
            col[i + j] += (uint64_t)__double_as_longlong(lo);
        }
    }
    const double pinv = d_from52(P52<P>::pinv());
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double td = d_from52(col[i] & MASK52);
        double h2 = __fma_rz(td, pinv, C1);
        double qd = __fma_rz(td, pinv, C2 - h2) - 4503599627370496.0;   // q = t * (-p^-1) mod 2^52
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const double pj = d_from52(P52<P>::limb(j));
            double hi = __fma_rz(qd, pj, C1);
            double lo = __fma_rz(qd, pj, C2 - hi);
            col[i + j + 1] += (uint64_t)__double_as_longlong(hi);
            col[i + j] += (uint64_t)__double_as_longlong(lo);
        }
        col[i + 1] += col[i] >> 52;                              // the low 52 bits of col[i] are zero now
    }
    uint64_t t[5], carry = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) { uint64_t v = col[5 + k] + carry; t[k] = v & MASK52; carry = v >> 52; }
    uint64_t d[5];
    long long borrow = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        long long v = (long long)t[k] - (long long)P52<P>::limb(k) + borrow;
        d[k] = (uint64_t)v & MASK52;
        borrow = v >> 52;
    }
    fe52_t r;
#pragma unroll
    for (int k = 0; k < 5; ++k) r.v[k] = borrow ? t[k] : d[k];
    return r;
}

// 8 x 32-bit limbs (ff.cuh) <-> 5 x 52-bit limbs; plain re-slicing of the same integer (no Montgomery factor change)
__device__ __forceinline__ fe52_t fe52_from_fe(const fe_t &x) {
    uint64_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (uint64_t)x.v[2 * i] | ((uint64_t)x.v[2 * i + 1] << 32);
    fe52_t r;
    r.v[0] = w[0] & MASK52;
    r.v[1] = ((w[0] >> 52) | (w[1] << 12)) & MASK52;
    r.v[2] = ((w[1] >> 40) | (w[2] << 24)) & MASK52;
    r.v[3] = ((w[2] >> 28) | (w[3] << 36)) & MASK52;
    r.v[4] = w[3] >> 16;
    return r;
}
__device__ __forceinline__ fe_t fe_from_fe52(const fe52_t &x) {
    uint64_t w[4];
    w[0] = x.v[0] | (x.v[1] << 52);
    w[1] = (x.v[1] >> 12) | (x.v[2] << 40);
    w[2] = (x.v[2] >> 24) | (x.v[3] << 28);
    w[3] = (x.v[3] >> 36) | (x.v[4] << 16);
    fe_t r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)w[i]; r.v[2 * i + 1] = (uint32_t)(w[i] >> 32); }
    return r;
}
#endif  // __CUDACC__
}  // namespace zkb
