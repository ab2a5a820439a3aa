// ec.cuh -- BN254 G1 (y^2 = x^3 + 3) group law on the device.
//
// Replaces ark-ec 0.3 `short_weierstrass_jacobian::{GroupAffine, GroupProjective}` for the bucket
// method.  Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// a mixed addition costs 8M + 2S and needs no field inversion.  Only the affine image of the final
// sum is canonical, and that is what is compared bit-for-bit against arkworks' VariableBaseMSM result
// (plonk-core/src/commitment.rs:42), so the choice of projective system is free.
//
// Infinity: XYZZ with ZZ == 0; affine (x, y) == (0, 0) at the C boundary (not a curve point).
#pragma once
#include "ff.cuh"

namespace zkb {

struct g1a_t { fe_t x, y; };                 // affine, Montgomery Fq
struct g1x_t { fe_t x, y, zz, zzz; };        // XYZZ

#ifdef __CUDACC__
typedef FqP Q;

__device__ __forceinline__ bool g1a_is_inf(const g1a_t &p) { return fis_zero<Q>(p.x) && fis_zero<Q>(p.y); }
__device__ __forceinline__ bool g1x_is_inf(const g1x_t &p) { return fis_zero<Q>(p.zz); }

__device__ __forceinline__ g1x_t g1x_inf() {
    g1x_t r;
    r.x = fzero<Q>(); r.y = fzero<Q>(); r.zz = fzero<Q>(); r.zzz = fzero<Q>();
    return r;
}

__device__ __forceinline__ g1x_t g1x_from_affine(const g1a_t &p) {
    g1x_t r;
    if (g1a_is_inf(p)) return g1x_inf();
    r.x = p.x; r.y = p.y; r.zz = fone<Q>(); r.zzz = fone<Q>();
    return r;
}

// 2 * (affine p), p finite  (mdbl-2008-s-1)
static __device__ __noinline__ g1x_t g1x_double_affine(const g1a_t &p) {
    g1x_t r;
    fe_t u = fdbl<Q>(p.y);
    fe_t v = fsqr<Q>(u);
    fe_t w = fmul<Q>(u, v);
    fe_t s = fmul<Q>(p.x, v);
    fe_t xx = fsqr<Q>(p.x);
    fe_t m = fadd<Q>(fdbl<Q>(xx), xx);
    r.x = fsub<Q>(fsub<Q>(fsqr<Q>(m), s), s);
    r.y = fmsub2<Q>(m, fsub<Q>(s, r.x), w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * p  (dbl-2008-s-1)
static __device__ __noinline__ g1x_t g1x_double(const g1x_t &p) {
    if (g1x_is_inf(p)) return p;
    g1x_t r;
    fe_t u = fdbl<Q>(p.y);
    fe_t v = fsqr<Q>(u);
    fe_t w = fmul<Q>(u, v);
    fe_t s = fmul<Q>(p.x, v);
    fe_t xx = fsqr<Q>(p.x);
    fe_t m = fadd<Q>(fdbl<Q>(xx), xx);
    r.x = fsub<Q>(fsub<Q>(fsqr<Q>(m), s), s);
    r.y = fmsub2<Q>(m, fsub<Q>(s, r.x), w, p.y);
    r.zz = fmul<Q>(v, p.zz);
    r.zzz = fmul<Q>(w, p.zzz);
    return r;
}

// acc += q (affine), all special cases handled  (madd-2008-s)
__device__ __forceinline__ void g1x_add_mixed(g1x_t &acc, const g1a_t &q) {
    if (g1a_is_inf(q)) return;
    if (g1x_is_inf(acc)) { acc.x = q.x; acc.y = q.y; acc.zz = fone<Q>(); acc.zzz = fone<Q>(); return; }
    fe_t u2 = fmul<Q>(q.x, acc.zz);
    fe_t s2 = fmul<Q>(q.y, acc.zzz);
    fe_t p = fsub<Q>(u2, acc.x);
    fe_t r = fsub<Q>(s2, acc.y);
    if (fis_zero<Q>(p)) {                         // same x: doubling or cancellation (rare)
        if (fis_zero<Q>(r)) acc = g1x_double_affine(q);
        else acc = g1x_inf();
        return;
    }
    fe_t pp = fsqr<Q>(p);
    fe_t ppp = fmul<Q>(p, pp);
    fe_t qq = fmul<Q>(acc.x, pp);
    fe_t x3 = fsub<Q>(fsub<Q>(fsub<Q>(fsqr<Q>(r), ppp), qq), qq);
    fe_t y3 = fmsub2<Q>(r, fsub<Q>(qq, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = fmul<Q>(acc.zz, pp);
    acc.zzz = fmul<Q>(acc.zzz, ppp);
}

// acc += q (XYZZ)  (add-2008-s)
__device__ __forceinline__ void g1x_add(g1x_t &acc, const g1x_t &q) {
    if (g1x_is_inf(q)) return;
    if (g1x_is_inf(acc)) { acc = q; return; }
    fe_t u1 = fmul<Q>(acc.x, q.zz);
    fe_t u2 = fmul<Q>(q.x, acc.zz);
    fe_t s1 = fmul<Q>(acc.y, q.zzz);
    fe_t s2 = fmul<Q>(q.y, acc.zzz);
    fe_t p = fsub<Q>(u2, u1);
    fe_t r = fsub<Q>(s2, s1);
    if (fis_zero<Q>(p)) {
        if (fis_zero<Q>(r)) acc = g1x_double(acc);
        else acc = g1x_inf();
        return;
    }
    fe_t pp = fsqr<Q>(p);
    fe_t ppp = fmul<Q>(p, pp);
    fe_t qq = fmul<Q>(u1, pp);
    fe_t x3 = fsub<Q>(fsub<Q>(fsub<Q>(fsqr<Q>(r), ppp), qq), qq);
    fe_t y3 = fmsub2<Q>(r, fsub<Q>(qq, x3), s1, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = fmul<Q>(fmul<Q>(acc.zz, q.zz), pp);
    acc.zzz = fmul<Q>(fmul<Q>(acc.zzz, q.zzz), ppp);
}

// acc + q by FOUR cooperating lanes (consecutive, aligned to 4; sub = lane & 3; all four hold both operands and all four
// return the sum).  The 14 products of add-2008-s are independent in groups of four, so the four lanes run them as 4 rounds
// of one product each with the results exchanged by shuffles: ~2.7k cycles instead of ~7.7k for a lone thread.  For the
// latency-bound ends of the bucket method (binary tail of the window reduction, oversized-bucket combine), where a level is
// one dependent addition and nothing else can hide it.  Same field values as g1x_add (canonical residues), so the same point.
__device__ __forceinline__ fe_t fe_from_lane(const fe_t &v, int src, uint32_t mask) {
    fe_t r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __shfl_sync(mask, v.v[k], src, 4);
    return r;
}
__device__ __forceinline__ fe_t fe_pick(uint32_t sub, const fe_t &a0, const fe_t &a1, const fe_t &a2, const fe_t &a3) {
    fe_t r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = sub == 0 ? a0.v[k] : sub == 1 ? a1.v[k] : sub == 2 ? a2.v[k] : a3.v[k];
    return r;
}
static __device__ __noinline__ g1x_t g1x_add_coop4(const g1x_t &a, const g1x_t &b, uint32_t sub) {
    if (g1x_is_inf(b)) return a;                               // every branch is uniform inside the group of four
    if (g1x_is_inf(a)) return b;
    const uint32_t mask = 0xFu << ((threadIdx.x & 31u) & ~3u);
    // round 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    fe_t m = fmul<Q>(fe_pick(sub, a.x, b.x, a.y, b.y), fe_pick(sub, b.zz, a.zz, b.zzz, a.zzz));
    const fe_t u1 = fe_from_lane(m, 0, mask), u2 = fe_from_lane(m, 1, mask), s1 = fe_from_lane(m, 2, mask), s2 = fe_from_lane(m, 3, mask);
    const fe_t p = fsub<Q>(u2, u1), r = fsub<Q>(s2, s1);
    if (fis_zero<Q>(p)) {
        if (fis_zero<Q>(r)) return g1x_double(a);
        return g1x_inf();
    }
    // round 2: PP = P^2, RR = R^2, ZZ1 ZZ2, ZZZ1 ZZZ2
    m = fmul<Q>(fe_pick(sub, p, r, a.zz, a.zzz), fe_pick(sub, p, r, b.zz, b.zzz));
    const fe_t pp = fe_from_lane(m, 0, mask), rr = fe_from_lane(m, 1, mask), zz12 = fe_from_lane(m, 2, mask), zzz12 = fe_from_lane(m, 3, mask);
    // round 3: PPP = P PP, Q = U1 PP, ZZ3 = ZZ1 ZZ2 PP  (lane 3 repeats lane 2's product)
    m = fmul<Q>(fe_pick(sub, p, u1, zz12, zz12), pp);
    const fe_t ppp = fe_from_lane(m, 0, mask), q = fe_from_lane(m, 1, mask), zz3 = fe_from_lane(m, 2, mask);
    g1x_t o;
    o.x = fsub<Q>(fsub<Q>(fsub<Q>(rr, ppp), q), q);
    // round 4: R (Q - X3), S1 PPP, ZZZ3 = ZZZ1 ZZZ2 PPP
    m = fmul<Q>(fe_pick(sub, r, s1, zzz12, zzz12), fe_pick(sub, fsub<Q>(q, o.x), ppp, ppp, ppp));
    o.y = fsub<Q>(fe_from_lane(m, 0, mask), fe_from_lane(m, 1, mask));
    o.zz = zz3;
    o.zzz = fe_from_lane(m, 2, mask);
    return o;
}
// coordinate `sub` (0: x, 1: y, 2: zz, 3: zzz) of a point stored by lane `sub` of a group: one coalesced 128-byte store
__device__ __forceinline__ void g1x_store_coop4(void *p, const g1x_t &a, uint32_t sub) {
    fstore(reinterpret_cast<char *>(p) + 32 * sub, fe_pick(sub, a.x, a.y, a.zz, a.zzz));
}

__device__ __forceinline__ g1a_t g1a_neg(const g1a_t &p) {
    g1a_t r;
    r.x = p.x;
    r.y = fneg<Q>(p.y);
    return r;
}

__device__ __forceinline__ g1a_t g1a_load(const void *p) {
    g1a_t r;
    r.x = fload_ro(p);
    r.y = fload_ro(reinterpret_cast<const char *>(p) + 32);
    return r;
}
__device__ __forceinline__ g1x_t g1x_load(const void *p) {
    const char *c = reinterpret_cast<const char *>(p);
    g1x_t r;
    r.x = fload(c); r.y = fload(c + 32); r.zz = fload(c + 64); r.zzz = fload(c + 96);
    return r;
}
__device__ __forceinline__ void g1x_store(void *p, const g1x_t &a) {
    char *c = reinterpret_cast<char *>(p);
    fstore(c, a.x); fstore(c + 32, a.y); fstore(c + 64, a.zz); fstore(c + 96, a.zzz);
}

// XYZZ -> affine on the device (one field inversion; used off the critical path and in tests)
static __device__ __noinline__ g1a_t g1x_to_affine(const g1x_t &p) {
    g1a_t r;
    if (g1x_is_inf(p)) { r.x = fzero<Q>(); r.y = fzero<Q>(); return r; }
    fe_t zi = finv<Q>(p.zzz);                    // 1/ZZZ
    fe_t zz_i = fmul<Q>(zi, p.zz);               // ZZ/ZZZ = 1/Z
    zz_i = fsqr<Q>(zz_i);                        // 1/ZZ
    r.x = fmul<Q>(p.x, zz_i);
    r.y = fmul<Q>(p.y, zi);
    return r;
}
#endif  // __CUDACC__
}  // namespace zkb
