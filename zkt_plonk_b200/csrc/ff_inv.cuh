// ff_inv.cuh -- modular inversion on the ALU pipe, for the batched-affine pair rounds (msm_pairs.cuh: an opt-in mode of the
// bucket accumulation, not the default product path).
//
// finv<P> (ff.cuh) is Fermat: ~380 dependent Montgomery products, all on the multiplier pipe, ~93 us as one chain.  A batch
// inversion needs ONE inverse per warp and batch, computed by one lane while the SM's other warps keep multiplying, so what
// matters is its latency and that it stays off the multiplier pipe: binary extended Euclid (HAC 14.61), ~530 shift /
// subtract steps on 8 x 32-bit limbs.  Bit-exact: every pair-round MSM test (tests/test_gpu_msm.py) goes through it.
#pragma once
#include "ff.cuh"

namespace zkb {
#ifdef __CUDACC__

// binary extended Euclid on 8 x 32-bit limbs.
// in: x in [1, p) (any representation); out: x^-1 mod p as a plain integer relation (out * x = 1 mod p).
__device__ __forceinline__ void shr1(uint32_t (&a)[8], uint32_t top) {
#pragma unroll
    for (int i = 0; i < 7; ++i) a[i] = __funnelshift_r(a[i], a[i + 1], 1);
    a[7] = __funnelshift_r(a[7], top, 1);
}
__device__ __forceinline__ bool is_one(const uint32_t (&a)[8]) {
    return a[0] == 1 && (a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0;
}
// halve a residue: a even -> a / 2, else (a + p) / 2   (a < p < 2^254: the sum fits 256 bits)
template <class P>
__device__ __forceinline__ void half_mod(uint32_t (&a)[8]) {
    if (a[0] & 1) {
        uint32_t m[8], s[8];
        load_mod<P>(m);
        uint32_t c = add8(s, a, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = s[i];
        shr1(a, c);
    } else {
        shr1(a, 0);
    }
}
template <class P>
__device__ __noinline__ fe_t inv_binary_euclid(const fe_t &x) {
    uint32_t u[8], v[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = x.v[i]; v[i] = P::mod(i); x1[i] = i == 0; x2[i] = 0; }
    while (!is_one(u) && !is_one(v)) {
        while (!(u[0] & 1)) { shr1(u, 0); half_mod<P>(x1); }
        while (!(v[0] & 1)) { shr1(v, 0); half_mod<P>(x2); }
        uint32_t d[8];
        uint32_t borrow = sub8(d, u, v);                         // u - v
        if (!borrow) {                                           // u >= v: u -= v, x1 -= x2 (mod p)
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = d[i];
            fe_t a, b;
#pragma unroll
            for (int i = 0; i < 8; ++i) { a.v[i] = x1[i]; b.v[i] = x2[i]; }
            a = fsub<P>(a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) x1[i] = a.v[i];
        } else {                                                 // v -= u, x2 -= x1 (mod p)
            sub8(d, v, u);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = d[i];
            fe_t a, b;
#pragma unroll
            for (int i = 0; i < 8; ++i) { a.v[i] = x2[i]; b.v[i] = x1[i]; }
            a = fsub<P>(a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) x2[i] = a.v[i];
        }
    }
    fe_t r;
    const bool first = is_one(u);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = first ? x1[i] : x2[i];
    return r;
}
// Montgomery inverse: for x = a R the Euclid result is a^-1 R^-1; one product by R^3 gives a^-1 R.
template <class P>
__device__ __forceinline__ fe_t finv_euclid(const fe_t &x) {
    fe_t r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.v[i] = P::r2(i);
    fe_t r3 = fmul<P>(r2, r2);                                   // R^2 * R^2 * R^-1 = R^3
    return fmul<P>(inv_binary_euclid<P>(x), r3);
}

#endif  // __CUDACC__
}  // namespace zkb
