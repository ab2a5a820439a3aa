// host_ff.h -- small host-side BN254 Fr/Fq helpers for the runtime around the kernels: domain
// constants (roots of unity, generator powers, n^-1), challenge-dependent scalars and the final
// XYZZ -> affine conversion of an MSM result.  O(1) work per call; all bulk arithmetic is on the device.
// Domain constants follow ark-poly 0.3 Radix2EvaluationDomain::new (group_gen = TWO_ADIC_ROOT squared
// 28 - log_n times; coset generator = Fr::multiplicative_generator() = 5).
#pragma once
#include <stdint.h>
#include <string.h>

namespace zkb {
namespace host {

typedef unsigned __int128 u128;
struct Fe { uint64_t l[4]; };

struct Params { uint64_t p[4], one[4], r2[4], inv; };

static const Params FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
    0xc2e1f593efffffffULL};
static const Params FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
    0x87d20782e4866389ULL};

inline bool ge(const uint64_t *a, const uint64_t *b) {
    for (int i = 3; i >= 0; --i) if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
inline void sub_raw(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - b[i] - br;
        o[i] = (uint64_t)t;
        br = (uint64_t)(t >> 64) & 1;
    }
}
inline bool is_zero(const Fe &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline Fe one(const Params &P) { Fe r; memcpy(r.l, P.one, 32); return r; }

inline Fe add(const Fe &a, const Fe &b, const Params &P) {
    Fe r; u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (ge(r.l, P.p)) sub_raw(r.l, r.l, P.p);
    return r;
}
inline Fe sub(const Fe &a, const Fe &b, const Params &P) {
    Fe r;
    if (ge(a.l, b.l)) { sub_raw(r.l, a.l, b.l); return r; }
    uint64_t t[4];
    sub_raw(t, b.l, a.l);
    sub_raw(r.l, P.p, t);
    return r;
}
inline Fe mul(const Fe &a, const Fe &b, const Params &P) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * P.inv;
        c = ((u128)m * P.p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; ++j) { c += (u128)m * P.p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fe r;
    if (t[4] || ge(t, P.p)) sub_raw(r.l, t, P.p); else memcpy(r.l, t, 32);
    return r;
}
inline Fe sqr(const Fe &a, const Params &P) { return mul(a, a, P); }
inline Fe from_u64(uint64_t v, const Params &P) {
    Fe t = {{v, 0, 0, 0}}, r2;
    memcpy(r2.l, P.r2, 32);
    return mul(t, r2, P);
}
inline Fe pow(const Fe &a, const uint64_t e[4], const Params &P) {
    Fe acc = one(P);
    for (int i = 255; i >= 0; --i) {
        acc = sqr(acc, P);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a, P);
    }
    return acc;
}
inline Fe pow_u64(const Fe &a, uint64_t e, const Params &P) {
    uint64_t ee[4] = {e, 0, 0, 0};
    return pow(a, ee, P);
}
inline Fe inv(const Fe &a, const Params &P) {           // a != 0
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub_raw(e, P.p, two);
    return pow(a, e, P);
}

// Fr root of unity of order 2^log_n, as Radix2EvaluationDomain::new derives it.
inline Fe fr_root_of_unity(unsigned log_n) {
    static const uint64_t T[4] = {0x9b9709143e1f593fULL, 0x181585d2833e8487ULL, 0x131a029b85045b68ULL,
                                  0x000000030644e72eULL};   // (r - 1) >> 28
    Fe w = pow(from_u64(5, FR), T, FR);
    for (unsigned i = log_n; i < 28; ++i) w = sqr(w, FR);
    return w;
}

}  // namespace host
}  // namespace zkb
