// host_ff.h -- small host-side Fr/Fq helpers for the runtime around the kernels: domain constants (roots of unity, generator
// powers, n^-1), challenge-dependent scalars and the final XYZZ -> affine conversion of an MSM result.  O(1) work per call;
// all bulk arithmetic is on the device.  Generic over the limb count (4 x u64: every Fr and BN254's Fq; 6 x u64: the Fq of
// BLS12-381 / BLS12-377); the constants of the curve this build is for come from curve_params.h.
// Domain constants follow ark-poly 0.3 Radix2EvaluationDomain::new (group_gen = TWO_ADIC_ROOT_OF_UNITY squared
// TWO_ADICITY - log_n times; coset generator = Fr::multiplicative_generator(): 5 on BN254, 7 on BLS12-381, 22 on BLS12-377).
#pragma once
#include <stdint.h>
#include <string.h>

#include "curve_params.h"

namespace zkb {
namespace host {

typedef unsigned __int128 u128;
template <int L> struct FeT { uint64_t l[L]; };
typedef FeT<FR_L> Fe;                    // a scalar (and, on BN254, a base-field element)
typedef FeT<FQ_L> Fq;                    // a base-field element
typedef ParamsT<FR_L> Params;

template <int L> inline bool ge(const uint64_t *a, const uint64_t *b) {
    for (int i = L - 1; i >= 0; --i) if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
template <int L> inline void sub_raw(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t br = 0;
    for (int i = 0; i < L; ++i) {
        u128 t = (u128)a[i] - b[i] - br;
        o[i] = (uint64_t)t;
        br = (uint64_t)(t >> 64) & 1;
    }
}
// 4-limb forms used by callers that handle raw 256-bit scalars
inline bool ge(const uint64_t *a, const uint64_t *b) { return ge<4>(a, b); }
inline void sub_raw(uint64_t *o, const uint64_t *a, const uint64_t *b) { sub_raw<4>(o, a, b); }

template <int L> inline bool is_zero(const FeT<L> &a) {
    uint64_t d = 0;
    for (int i = 0; i < L; ++i) d |= a.l[i];
    return d == 0;
}
template <int L> inline FeT<L> one(const ParamsT<L> &P) { FeT<L> r; memcpy(r.l, P.one, 8 * L); return r; }

template <int L> inline FeT<L> add(const FeT<L> &a, const FeT<L> &b, const ParamsT<L> &P) {
    FeT<L> r; u128 c = 0;
    for (int i = 0; i < L; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || ge<L>(r.l, P.p)) sub_raw<L>(r.l, r.l, P.p);
    return r;
}
template <int L> inline FeT<L> sub(const FeT<L> &a, const FeT<L> &b, const ParamsT<L> &P) {
    FeT<L> r;
    if (ge<L>(a.l, b.l)) { sub_raw<L>(r.l, a.l, b.l); return r; }
    uint64_t t[L];
    sub_raw<L>(t, b.l, a.l);
    sub_raw<L>(r.l, P.p, t);
    return r;
}
template <int L> inline FeT<L> mul(const FeT<L> &a, const FeT<L> &b, const ParamsT<L> &P) {
    uint64_t t[L + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < L; ++i) {
        u128 c = 0;
        for (int j = 0; j < L; ++j) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[L]; t[L] = (uint64_t)c; t[L + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * P.inv;
        c = ((u128)m * P.p[0] + t[0]) >> 64;
        for (int j = 1; j < L; ++j) { c += (u128)m * P.p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[L]; t[L - 1] = (uint64_t)c; t[L] = t[L + 1] + (uint64_t)(c >> 64);
    }
    FeT<L> r;
    if (t[L] || ge<L>(t, P.p)) sub_raw<L>(r.l, t, P.p); else memcpy(r.l, t, 8 * L);
    return r;
}
template <int L> inline FeT<L> sqr(const FeT<L> &a, const ParamsT<L> &P) { return mul(a, a, P); }
template <int L> inline FeT<L> from_u64(uint64_t v, const ParamsT<L> &P) {
    FeT<L> t, r2;
    memset(t.l, 0, 8 * L);
    t.l[0] = v;
    memcpy(r2.l, P.r2, 8 * L);
    return mul(t, r2, P);
}
template <int L> inline FeT<L> pow(const FeT<L> &a, const uint64_t *e /* L limbs */, const ParamsT<L> &P) {
    FeT<L> acc = one(P);
    for (int i = 64 * L - 1; i >= 0; --i) {
        acc = sqr(acc, P);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a, P);
    }
    return acc;
}
template <int L> inline FeT<L> pow_u64(const FeT<L> &a, uint64_t e, const ParamsT<L> &P) {
    uint64_t ee[L];
    memset(ee, 0, sizeof ee);
    ee[0] = e;
    return pow(a, ee, P);
}
template <int L> inline FeT<L> inv(const FeT<L> &a, const ParamsT<L> &P) {           // a != 0
    uint64_t e[L], two[L];
    memset(two, 0, sizeof two);
    two[0] = 2;
    sub_raw<L>(e, P.p, two);
    return pow(a, e, P);
}

// Fr root of unity of order 2^log_n, as Radix2EvaluationDomain::new derives it: GENERATOR^T squared TWO_ADICITY - log_n times.
inline Fe fr_root_of_unity(unsigned log_n) {
    Fe w = pow(from_u64(FR_GENERATOR, FR), FR_T, FR);
    for (unsigned i = log_n; i < FR_TWO_ADICITY; ++i) w = sqr(w, FR);
    return w;
}

}  // namespace host
}  // namespace zkb
