// ec.cuh -- G1 group law on the device for the curve of this build (y^2 = x^3 + b: BN254, BLS12-381, BLS12-377; the
// formulas below use a = 0 and never b).
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

struct g1a_t { fq_t x, y; };                 // affine, Montgomery Fq (64 bytes on BN254, 96 on the BLS12 curves)
struct g1x_t { fq_t x, y, zz, zzz; };        // XYZZ
constexpr int FQ_BYTES = 4 * FqP::N;         // bytes of one base-field element

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
    fq_t u = fdbl<Q>(p.y);
    fq_t v = fsqr<Q>(u);
    fq_t w = fmul<Q>(u, v);
    fq_t s = fmul<Q>(p.x, v);
    fq_t xx = fsqr<Q>(p.x);
    fq_t m = fadd<Q>(fdbl<Q>(xx), xx);
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
    fq_t u = fdbl<Q>(p.y);
    fq_t v = fsqr<Q>(u);
    fq_t w = fmul<Q>(u, v);
    fq_t s = fmul<Q>(p.x, v);
    fq_t xx = fsqr<Q>(p.x);
    fq_t m = fadd<Q>(fdbl<Q>(xx), xx);
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
    fq_t u2 = fmul<Q>(q.x, acc.zz);
    fq_t s2 = fmul<Q>(q.y, acc.zzz);
    fq_t p = fsub<Q>(u2, acc.x);
    fq_t r = fsub<Q>(s2, acc.y);
    if (fis_zero<Q>(p)) {                         // same x: doubling or cancellation (rare)
        if (fis_zero<Q>(r)) acc = g1x_double_affine(q);
        else acc = g1x_inf();
        return;
    }
    fq_t pp = fsqr<Q>(p);
    fq_t ppp = fmul<Q>(p, pp);
    fq_t qq = fmul<Q>(acc.x, pp);
    fq_t x3 = fsub<Q>(fsub<Q>(fsub<Q>(fsqr<Q>(r), ppp), qq), qq);
    fq_t y3 = fmsub2<Q>(r, fsub<Q>(qq, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = fmul<Q>(acc.zz, pp);
    acc.zzz = fmul<Q>(acc.zzz, ppp);
}

// acc += q (XYZZ)  (add-2008-s)
__device__ __forceinline__ void g1x_add(g1x_t &acc, const g1x_t &q) {
    if (g1x_is_inf(q)) return;
    if (g1x_is_inf(acc)) { acc = q; return; }
    fq_t u1 = fmul<Q>(acc.x, q.zz);
    fq_t u2 = fmul<Q>(q.x, acc.zz);
    fq_t s1 = fmul<Q>(acc.y, q.zzz);
    fq_t s2 = fmul<Q>(q.y, acc.zzz);
    fq_t p = fsub<Q>(u2, u1);
    fq_t r = fsub<Q>(s2, s1);
    if (fis_zero<Q>(p)) {
        if (fis_zero<Q>(r)) acc = g1x_double(acc);
        else acc = g1x_inf();
        return;
    }
    fq_t pp = fsqr<Q>(p);
    fq_t ppp = fmul<Q>(p, pp);
    fq_t qq = fmul<Q>(u1, pp);
    fq_t x3 = fsub<Q>(fsub<Q>(fsub<Q>(fsqr<Q>(r), ppp), qq), qq);
    fq_t y3 = fmsub2<Q>(r, fsub<Q>(qq, x3), s1, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = fmul<Q>(fmul<Q>(acc.zz, q.zz), pp);
    acc.zzz = fmul<Q>(fmul<Q>(acc.zzz, q.zzz), ppp);
}

__device__ __forceinline__ g1a_t g1a_neg(const g1a_t &p) {
    g1a_t r;
    r.x = p.x;
    r.y = fneg<Q>(p.y);
    return r;
}

__device__ __forceinline__ g1a_t g1a_load(const void *p) {
    g1a_t r;
    r.x = floadn_ro<FqP::N>(p);
    r.y = floadn_ro<FqP::N>(reinterpret_cast<const char *>(p) + FQ_BYTES);
    return r;
}
__device__ __forceinline__ g1x_t g1x_load(const void *p) {
    const char *c = reinterpret_cast<const char *>(p);
    g1x_t r;
    r.x = floadn<FqP::N>(c); r.y = floadn<FqP::N>(c + FQ_BYTES); r.zz = floadn<FqP::N>(c + 2 * FQ_BYTES); r.zzz = floadn<FqP::N>(c + 3 * FQ_BYTES);
    return r;
}
__device__ __forceinline__ void g1x_store(void *p, const g1x_t &a) {
    char *c = reinterpret_cast<char *>(p);
    fstore(c, a.x); fstore(c + FQ_BYTES, a.y); fstore(c + 2 * FQ_BYTES, a.zz); fstore(c + 3 * FQ_BYTES, a.zzz);
}

// XYZZ -> affine on the device (one field inversion; used off the critical path and in tests)
static __device__ __noinline__ g1a_t g1x_to_affine(const g1x_t &p) {
    g1a_t r;
    if (g1x_is_inf(p)) { r.x = fzero<Q>(); r.y = fzero<Q>(); return r; }
    fq_t zi = finv<Q>(p.zzz);                    // 1/ZZZ
    fq_t zz_i = fmul<Q>(zi, p.zz);               // ZZ/ZZZ = 1/Z
    zz_i = fsqr<Q>(zz_i);                        // 1/ZZ
    r.x = fmul<Q>(p.x, zz_i);
    r.y = fmul<Q>(p.y, zi);
    return r;
}
#endif  // __CUDACC__
}  // namespace zkb
