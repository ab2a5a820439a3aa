/*
 * zkb_oracle.c -- CPU restatement of the zkt-plonk prover hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED.  The arithmetic on this path lives in un-vendored crates.io dependencies of the
 * reference (ark-ff / ark-ec / ark-poly / ark-poly-commit / ark-bn254, all "0.3",
 * /root/reference/plonk-core/Cargo.toml:19-24).  No Rust toolchain and no copy of those crates exists
 * in this image, and the reference's tests hold no golden vector for MSM / NTT / commitment bytes
 * (SURVEY.md section 4), so this file restates the PUBLISHED algorithms of those crates and is checked
 * against (1) oracle/pyref.py, an independent Python big-int implementation of the mathematical
 * definitions, (2) the public BN254 (alt_bn128) known answers 2G / 3G, and (3) the algebraic identities
 * the reference's own unit tests use (permutation/mod.rs:328-392, lookup/mod.rs:101-164).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may load
 * this library.  The product (zkt_plonk_b200/, libzkb200.so) never links or calls it.
 *
 * What each function follows:
 *   fp_mul / fp_add / ...      ark-ff 0.3 Fp256 Montgomery arithmetic (R = 2^256, 4 x u64 LE limbs)
 *   g1_*                       ark-ec 0.3 short_weierstrass_jacobian::GroupProjective
 *                              (add-2007-bl, madd-2007-bl, dbl-2009-l for a = 0)
 *   zko_msm_g1                 ark-ec 0.3 msm::VariableBaseMSM::multi_scalar_mul, as called from
 *                              plonk-core/src/commitment.rs:42 and (through ark-poly-commit
 *                              kzg10::commit) prove.rs:134,179,250,307,374
 *   zko_ntt                    ark-poly 0.3 Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
 *                              as wrapped by plonk-core/src/util.rs:63-140
 *   zko_z1_evals               plonk-core/src/permutation/mod.rs:181-254
 *   zko_z2_evals               plonk-core/src/lookup/mod.rs:25-82
 *   zko_quotient_evals         plonk-core/src/proof_system/quotient_poly.rs:98-224 with
 *                              keys/arithmetic.rs:67-81, keys/permutation.rs:97-137, keys/lookup.rs:81-122
 *   zko_epk_free_tables        keys/mod.rs:109-119 (x_coset, zh_coset, l_1_coset)
 *   zko_poly_eval              ark-poly 0.3 DensePolynomial::evaluate (Horner), as used by
 *                              linearization_poly.rs:55-75
 *   zko_poly_lincomb           the axpy chains of linearization_poly.rs:77-111 and SonicKZG10::open's
 *                              combination with powers of the opening challenge
 *   zko_poly_divide_linear     ark-poly-commit 0.3 kzg10::compute_witness_polynomial: (p(X) - p(z)) / (X - z)
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct { u64 l[4]; } fe;           /* one scalar-field element, LE limbs (Fr has 4 limbs on every curve) */
#include "zko_curve_params.h"
#define NQ ZKO_FQ_L
typedef struct { u64 l[NQ]; } fq;          /* one base-field element: 4 limbs on BN254, 6 on BLS12-381 / BLS12-377 */

typedef struct {
    u64 p[4];     /* modulus */
    u64 r[4];     /* R mod p  (Montgomery one) */
    u64 r2[4];    /* R^2 mod p */
    u64 inv;      /* -p^-1 mod 2^64 */
} fparams;
typedef struct { u64 p[NQ], r[NQ], r2[NQ], inv; } fqparams;

static const fparams FR = ZKO_FR_INIT;
static const fqparams FQ = ZKO_FQ_INIT;

/* ------------------------------------------------------------------ field arithmetic: fp_* over Fr's 4 limbs, fq_* over Fq's NQ */
#define NL 4
#define FE fe
#define FPARAMS fparams
#define FN(name) fp_##name
#include "zko_field.inc"
#undef NL
#undef FE
#undef FPARAMS
#undef FN
#define NL NQ
#define FE fq
#define FPARAMS fqparams
#define FN(name) fq_##name
#include "zko_field.inc"
#undef NL
#undef FE
#undef FPARAMS
#undef FN
static inline int fe_is_zero(const fe *a) { return fp_is_zero(a); }
static inline int fe_eq(const fe *a, const fe *b) { return fp_eq(a, b); }
static inline int geq(const u64 *a, const u64 *b) { return fp_geq(a, b); }
static inline void sub_nored(u64 *o, const u64 *a, const u64 *b) { fp_sub_nored(o, a, b); }

/* ------------------------------------------------------------------ G1 Jacobian (Z == 0 is infinity) */
static const fq FQ_ZERO = {{0}};
typedef struct { fq x, y, z; } g1j;
typedef struct { fq x, y; } g1a;            /* (0,0) encodes infinity at the C boundary */

static inline int g1a_is_inf(const g1a *p) { return fq_is_zero(&p->x) && fq_is_zero(&p->y); }
static inline void g1j_set_inf(g1j *p) { memset(p, 0, sizeof *p); fq_one(&p->x, &FQ); fq_one(&p->y, &FQ); }
static inline int g1j_is_inf(const g1j *p) { return fq_is_zero(&p->z); }

static void g1j_double(g1j *r, const g1j *p) {           /* dbl-2009-l */
    if (g1j_is_inf(p)) { *r = *p; return; }
    fq a, b, c, d, e, f, t;
    fq_sqr(&a, &p->x, &FQ); fq_sqr(&b, &p->y, &FQ); fq_sqr(&c, &b, &FQ);
    fq_add(&d, &p->x, &b, &FQ); fq_sqr(&d, &d, &FQ); fq_sub(&d, &d, &a, &FQ); fq_sub(&d, &d, &c, &FQ); fq_dbl(&d, &d, &FQ);
    fq_dbl(&e, &a, &FQ); fq_add(&e, &e, &a, &FQ);
    fq_sqr(&f, &e, &FQ);
    fq z3; fq_mul(&z3, &p->y, &p->z, &FQ); fq_dbl(&z3, &z3, &FQ);
    fq x3; fq_dbl(&t, &d, &FQ); fq_sub(&x3, &f, &t, &FQ);
    fq y3; fq_sub(&t, &d, &x3, &FQ); fq_mul(&y3, &e, &t, &FQ);
    fq_dbl(&c, &c, &FQ); fq_dbl(&c, &c, &FQ); fq_dbl(&c, &c, &FQ); fq_sub(&y3, &y3, &c, &FQ);
    r->x = x3; r->y = y3; r->z = z3;
}

static void g1j_add_mixed(g1j *r, const g1j *p, const g1a *q) {   /* madd-2007-bl */
    if (g1a_is_inf(q)) { *r = *p; return; }
    if (g1j_is_inf(p)) { r->x = q->x; r->y = q->y; fq_one(&r->z, &FQ); return; }
    fq z1z1, u2, s2, h, hh, i, j, rr, v, t;
    fq_sqr(&z1z1, &p->z, &FQ);
    fq_mul(&u2, &q->x, &z1z1, &FQ);
    fq_mul(&s2, &q->y, &p->z, &FQ); fq_mul(&s2, &s2, &z1z1, &FQ);
    if (fq_eq(&u2, &p->x) && fq_eq(&s2, &p->y)) { g1j_double(r, p); return; }
    fq_sub(&h, &u2, &p->x, &FQ);
    fq_sqr(&hh, &h, &FQ);
    fq_dbl(&i, &hh, &FQ); fq_dbl(&i, &i, &FQ);
    fq_mul(&j, &h, &i, &FQ);
    fq_sub(&rr, &s2, &p->y, &FQ); fq_dbl(&rr, &rr, &FQ);
    fq_mul(&v, &p->x, &i, &FQ);
    fq x3, y3, z3;
    fq_sqr(&x3, &rr, &FQ); fq_sub(&x3, &x3, &j, &FQ); fq_sub(&x3, &x3, &v, &FQ); fq_sub(&x3, &x3, &v, &FQ);
    fq_sub(&t, &v, &x3, &FQ); fq_mul(&y3, &rr, &t, &FQ);
    fq_mul(&t, &p->y, &j, &FQ); fq_dbl(&t, &t, &FQ); fq_sub(&y3, &y3, &t, &FQ);
    fq_add(&z3, &p->z, &h, &FQ); fq_sqr(&z3, &z3, &FQ); fq_sub(&z3, &z3, &z1z1, &FQ); fq_sub(&z3, &z3, &hh, &FQ);
    r->x = x3; r->y = y3; r->z = z3;
}

static void g1j_add(g1j *r, const g1j *p, const g1j *q) {         /* add-2007-bl */
    if (g1j_is_inf(p)) { *r = *q; return; }
    if (g1j_is_inf(q)) { *r = *p; return; }
    fq z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fq_sqr(&z1z1, &p->z, &FQ); fq_sqr(&z2z2, &q->z, &FQ);
    fq_mul(&u1, &p->x, &z2z2, &FQ); fq_mul(&u2, &q->x, &z1z1, &FQ);
    fq_mul(&s1, &p->y, &q->z, &FQ); fq_mul(&s1, &s1, &z2z2, &FQ);
    fq_mul(&s2, &q->y, &p->z, &FQ); fq_mul(&s2, &s2, &z1z1, &FQ);
    if (fq_eq(&u1, &u2) && fq_eq(&s1, &s2)) { g1j_double(r, p); return; }
    fq_sub(&h, &u2, &u1, &FQ);
    fq_dbl(&i, &h, &FQ); fq_sqr(&i, &i, &FQ);
    fq_mul(&j, &h, &i, &FQ);
    fq_sub(&rr, &s2, &s1, &FQ); fq_dbl(&rr, &rr, &FQ);
    fq_mul(&v, &u1, &i, &FQ);
    fq x3, y3, z3;
    fq_sqr(&x3, &rr, &FQ); fq_sub(&x3, &x3, &j, &FQ); fq_sub(&x3, &x3, &v, &FQ); fq_sub(&x3, &x3, &v, &FQ);
    fq_sub(&t, &v, &x3, &FQ); fq_mul(&y3, &rr, &t, &FQ);
    fq_mul(&t, &s1, &j, &FQ); fq_dbl(&t, &t, &FQ); fq_sub(&y3, &y3, &t, &FQ);
    fq_add(&z3, &p->z, &q->z, &FQ); fq_sqr(&z3, &z3, &FQ); fq_sub(&z3, &z3, &z1z1, &FQ); fq_sub(&z3, &z3, &z2z2, &FQ);
    fq_mul(&z3, &z3, &h, &FQ);
    r->x = x3; r->y = y3; r->z = z3;
}

static void g1j_to_affine(g1a *r, const g1j *p) {
    if (g1j_is_inf(p)) { memset(r, 0, sizeof *r); return; }
    fq zi, zi2, zi3;
    fq_inv(&zi, &p->z, &FQ); fq_sqr(&zi2, &zi, &FQ); fq_mul(&zi3, &zi2, &zi, &FQ);
    fq_mul(&r->x, &p->x, &zi2, &FQ); fq_mul(&r->y, &p->y, &zi3, &FQ);
}

/* ------------------------------------------------------------------ exported helpers */
#define API __attribute__((visibility("default")))

/* which curve this oracle library was compiled for, and its base-field width */
API void zko_curve_info(int *curve_id, int *fq_words, int *fr_bits) {
    if (curve_id) *curve_id = ZKO_CURVE;
    if (fq_words) *fq_words = NQ;
    if (fr_bits) *fr_bits = ZKO_FR_BITS;
}
/* field 0: Fr (4 words per element), field 1: Fq (NQ words per element) */
API void zko_to_mont(int field, u64 *out, const u64 *in, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field) fq_to_mont((fq *)(out + NQ * i), (const fq *)(in + NQ * i), &FQ);
        else fp_to_mont((fe *)(out + 4 * i), (const fe *)(in + 4 * i), &FR);
    }
}
API void zko_from_mont(int field, u64 *out, const u64 *in, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field) fq_from_mont((fq *)(out + NQ * i), (const fq *)(in + NQ * i), &FQ);
        else fp_from_mont((fe *)(out + 4 * i), (const fe *)(in + 4 * i), &FR);
    }
}
/* Map arbitrary words into [0,p): clear the bits above the modulus' bit length (two on BN254), subtract p once if needed. */
static unsigned top_bits(const u64 *p, int nl) { unsigned b = 64; while (b && !((p[nl - 1] >> (b - 1)) & 1)) --b; return b; }
API void zko_normalize(int field, u64 *data, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field) {
            u64 *d = data + NQ * i; unsigned tb = top_bits(FQ.p, NQ);
            if (tb < 64) d[NQ - 1] &= (((u64)1 << tb) - 1);
            if (fq_geq(d, FQ.p)) fq_sub_nored(d, d, FQ.p);
        } else {
            u64 *d = data + 4 * i; unsigned tb = top_bits(FR.p, 4);
            if (tb < 64) d[3] &= (((u64)1 << tb) - 1);
            if (geq(d, FR.p)) sub_nored(d, d, FR.p);
        }
    }
}
/* op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 inv(a) (a != 0) -- all Montgomery in, Montgomery out */
API void zko_fp_binop(int field, int op, u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field) {
            const fq *x = (const fq *)(a + NQ * i), *y = (const fq *)(b + NQ * i); fq *o = (fq *)(out + NQ * i);
            switch (op) {
                case 0: fq_mul(o, x, y, &FQ); break;
                case 1: fq_add(o, x, y, &FQ); break;
                case 2: fq_sub(o, x, y, &FQ); break;
                case 3: fq_sqr(o, x, &FQ); break;
                default: fq_inv(o, x, &FQ); break;
            }
            continue;
        }
        const fparams *P = &FR;
        const fe *x = (const fe *)(a + 4 * i), *y = (const fe *)(b + 4 * i); fe *o = (fe *)(out + 4 * i);
        switch (op) {
            case 0: fp_mul(o, x, y, P); break;
            case 1: fp_add(o, x, y, P); break;
            case 2: fp_sub(o, x, y, P); break;
            case 3: fp_sqr(o, x, P); break;
            default: fp_inv(o, x, P); break;
        }
    }
}

/* ------------------------------------------------------------------ NTT (ark-poly Radix2EvaluationDomain) */
static void fr_root_of_unity(fe *w, unsigned log_n) {
    /* TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((r-1)/2^TWO_ADICITY) (5^((r-1)/2^28) on BN254), then squared TWO_ADICITY - log_n times
     * (Radix2EvaluationDomain::new) */
    static const u64 T[4] = ZKO_FR_T_INIT;
    fe g; fp_from_u64(&g, ZKO_FR_GENERATOR, &FR);
    fp_pow(w, &g, T, &FR);
    for (unsigned i = log_n; i < ZKO_FR_TWO_ADICITY; ++i) fp_sqr(w, w, &FR);
}

static inline size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* in place, natural order in and out, data holds 2^log_n Montgomery Fr elements.
 * inverse: uses w^-1 and multiplies by n^-1; coset: g = 5 powers before (forward) / g^-1 powers after (inverse). */
API int zko_ntt(u64 *data, unsigned log_n, int inverse, int coset, int threads) {
    if (log_n > ZKO_FR_TWO_ADICITY || log_n > 40) return -1;
    size_t n = (size_t)1 << log_n;
    fe *x = (fe *)data;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    fe w; fr_root_of_unity(&w, log_n);
    if (inverse) fp_inv(&w, &w, &FR);
    if (coset && !inverse) {                   /* distribute_powers(coeffs, g) */
        fe g, cur; fp_from_u64(&g, ZKO_FR_GENERATOR, &FR); fp_one(&cur, &FR);
        for (size_t i = 0; i < n; ++i) { fp_mul(&x[i], &x[i], &cur, &FR); fp_mul(&cur, &cur, &g, &FR); }
    }
    if (n > 1) {
        size_t half = n / 2;
        fe *roots = (fe *)malloc(half * sizeof(fe));
        fp_one(&roots[0], &FR);
        for (size_t i = 1; i < half; ++i) fp_mul(&roots[i], &roots[i - 1], &w, &FR);
        /* decimation in frequency: in-order input, bit-reversed output (ark-poly io_helper shape) */
        for (size_t gap = half; gap >= 1; gap >>= 1) {
            size_t step = half / gap;          /* root index stride for this layer */
#pragma omp parallel for schedule(static) if (n >= 4096)
            for (size_t k = 0; k < half; ++k) {
                size_t blk = k / gap, j = k % gap;
                size_t lo = blk * 2 * gap + j, hi = lo + gap;
                fe s, d;
                fp_add(&s, &x[lo], &x[hi], &FR);
                fp_sub(&d, &x[lo], &x[hi], &FR);
                x[lo] = s;
                if (j) fp_mul(&x[hi], &d, &roots[j * step], &FR); else x[hi] = d;
            }
        }
        free(roots);
        for (size_t i = 0; i < n; ++i) {       /* derange */
            size_t r = bitrev(i, log_n);
            if (i < r) { fe t = x[i]; x[i] = x[r]; x[r] = t; }
        }
    }
    if (inverse) {
        fe ninv; fp_from_u64(&ninv, (u64)n, &FR); fp_inv(&ninv, &ninv, &FR);
        if (coset) {                           /* ifft's size_inv, then distribute_powers(evals, g^-1) */
            fe gi, cur = ninv; fp_from_u64(&gi, ZKO_FR_GENERATOR, &FR); fp_inv(&gi, &gi, &FR);
            for (size_t i = 0; i < n; ++i) { fp_mul(&x[i], &x[i], &cur, &FR); fp_mul(&cur, &cur, &gi, &FR); }
        } else {
#pragma omp parallel for schedule(static) if (n >= 4096)
            for (size_t i = 0; i < n; ++i) fp_mul(&x[i], &x[i], &ninv, &FR);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ MSM (ark-ec VariableBaseMSM) */
static unsigned ceil_log2(size_t x) { unsigned l = 0; while (((size_t)1 << l) < x) ++l; return l; }

static inline u64 scalar_window(const u64 s[4], unsigned start, unsigned c) {
    /* (scalar >> start) mod 2^c on the 256-bit little-endian integer */
    unsigned limb = start >> 6, off = start & 63;
    u64 v = s[limb] >> off;
    if (off && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
    return c >= 64 ? v : (v & (((u64)1 << c) - 1));
}

/* points: n x (x,y) Montgomery Fq, (0,0) = infinity; scalars: n x canonical 256-bit integers (into_repr) */
API int zko_msm_g1(const u64 *points, const u64 *scalars, size_t n, u64 *out_xy /* 2 NQ words */, int *is_inf, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    const g1a *bases = (const g1a *)points;
    unsigned c = n < 32 ? 3 : (ceil_log2(n) * 69 / 100) + 2;      /* ln_without_floats(size) + 2 */
    const unsigned num_bits = ZKO_FR_BITS;                                   /* Fr MODULUS_BITS */
    unsigned nwin = (num_bits + c - 1) / c;
    g1j *wsum = (g1j *)malloc(nwin * sizeof(g1j));
    const u64 one[4] = {1, 0, 0, 0};
    size_t nb = ((size_t)1 << c) - 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (unsigned w = 0; w < nwin; ++w) {
        unsigned w_start = w * c;
        g1j res; g1j_set_inf(&res); res.z = FQ_ZERO;
        g1j *buckets = (g1j *)malloc(nb * sizeof(g1j));
        for (size_t b = 0; b < nb; ++b) { buckets[b].z = FQ_ZERO; }
        for (size_t i = 0; i < n; ++i) {
            const u64 *s = scalars + 4 * i;
            if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;       /* filter(|s| !s.is_zero()) */
            if (memcmp(s, one, 32) == 0) {
                if (w_start == 0) g1j_add_mixed(&res, &res, &bases[i]);
            } else {
                u64 d = scalar_window(s, w_start, c);
                if (d) g1j_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
            }
        }
        g1j running; running.z = FQ_ZERO;
        for (size_t b = nb; b-- > 0;) {
            g1j_add(&running, &running, &buckets[b]);
            g1j_add(&res, &res, &running);
        }
        free(buckets);
        wsum[w] = res;
    }
    g1j total; total.z = FQ_ZERO;
    for (unsigned w = nwin; w-- > 1;) {
        g1j_add(&total, &total, &wsum[w]);
        for (unsigned k = 0; k < c; ++k) g1j_double(&total, &total);
    }
    g1j_add(&total, &total, &wsum[0]);
    free(wsum);
    g1a aff; g1j_to_affine(&aff, &total);
    memcpy(out_xy, &aff, sizeof aff);
    if (is_inf) *is_inf = g1j_is_inf(&total);
    return 0;
}

/* out[i] = scalars[i] * base, affine; scalars canonical.  Plain double-and-add (definition). */
API void zko_g1_mul(const u64 *base_xy, const u64 *scalars, size_t n, u64 *out_xy) {
    const g1a *B = (const g1a *)base_xy;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; ++i) {
        const u64 *s = scalars + 4 * i;
        g1j acc; acc.z = FQ_ZERO;
        for (int b = 255; b >= 0; --b) {
            g1j_double(&acc, &acc);
            if ((s[b >> 6] >> (b & 63)) & 1) g1j_add_mixed(&acc, &acc, B);
        }
        g1j_to_affine((g1a *)(out_xy + 2 * NQ * i), &acc);
    }
}

/* out = sum_i points[i] (affine in, affine out) */
API void zko_g1_sum(const u64 *points, size_t n, u64 *out_xy /* 2 NQ words */) {
    g1j acc; acc.z = FQ_ZERO;
    for (size_t i = 0; i < n; ++i) g1j_add_mixed(&acc, &acc, (const g1a *)(points + 2 * NQ * i));
    g1j_to_affine((g1a *)out_xy, &acc);
}

/* out[i] = start + i * step (affine), i < n: cheap synthetic base points for CPU-only timing runs. */
API void zko_g1_walk(const u64 *start_xy, const u64 *step_xy, size_t n, u64 *out_xy) {
    const g1a *S = (const g1a *)start_xy, *D = (const g1a *)step_xy;
    const size_t CH = 4096;
    size_t nch = (n + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < nch; ++c) {
        /* chunk start = start + (c*CH) * step by double-and-add */
        g1j acc; acc.z = FQ_ZERO;
        size_t k = c * CH;
        for (int b = 63; b >= 0; --b) { g1j_double(&acc, &acc); if ((k >> b) & 1) g1j_add_mixed(&acc, &acc, D); }
        g1j_add_mixed(&acc, &acc, S);
        size_t hi = (c + 1) * CH < n ? (c + 1) * CH : n;
        for (size_t i = c * CH; i < hi; ++i) {
            g1j_to_affine((g1a *)(out_xy + 2 * NQ * i), &acc);
            g1j_add_mixed(&acc, &acc, D);
        }
    }
}

API int zko_g1_on_curve(const u64 *xy) {
    const g1a *p = (const g1a *)xy;
    if (g1a_is_inf(p)) return 1;
    fq l, r, three; fq_sqr(&l, &p->y, &FQ);
    fq_sqr(&r, &p->x, &FQ); fq_mul(&r, &r, &p->x, &FQ); fq_from_u64(&three, ZKO_G1_B, &FQ); fq_add(&r, &r, &three, &FQ);
    return fq_eq(&l, &r);
}

/* ------------------------------------------------------------------ grand products */
static void fr_domain_elements(fe *roots, unsigned log_n) {
    size_t n = (size_t)1 << log_n; fe w; fr_root_of_unity(&w, log_n);
    fp_one(&roots[0], &FR);
    for (size_t i = 1; i < n; ++i) fp_mul(&roots[i], &roots[i - 1], &w, &FR);
}

/* permutation/mod.rs:181-254.  All arrays n Montgomery Fr; out gets the n evaluations of z1. */
API void zko_z1_evals(unsigned log_n, const u64 *beta_, const u64 *gamma_, const u64 *a_, const u64 *b_, const u64 *c_,
                      const u64 *s1_, const u64 *s2_, const u64 *s3_, u64 *out_) {
    size_t n = (size_t)1 << log_n;
    const fe *a = (const fe *)a_, *b = (const fe *)b_, *c = (const fe *)c_;
    const fe *s1 = (const fe *)s1_, *s2 = (const fe *)s2_, *s3 = (const fe *)s3_;
    fe beta = *(const fe *)beta_, gamma = *(const fe *)gamma_, *out = (fe *)out_;
    fe *roots = (fe *)malloc(n * sizeof(fe)); fr_domain_elements(roots, log_n);
    fe k1, k2; fp_from_u64(&k1, 7, &FR); fp_from_u64(&k2, 13, &FR);
    fe *prod = (fe *)malloc(n * sizeof(fe));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n - 1; ++i) {
        fe br, t, num, den, u;
        fp_mul(&br, &beta, &roots[i], &FR);
        fp_add(&num, &br, &a[i], &FR); fp_add(&num, &num, &gamma, &FR);
        fp_mul(&t, &k1, &br, &FR); fp_add(&t, &t, &b[i], &FR); fp_add(&t, &t, &gamma, &FR); fp_mul(&num, &num, &t, &FR);
        fp_mul(&t, &k2, &br, &FR); fp_add(&t, &t, &c[i], &FR); fp_add(&t, &t, &gamma, &FR); fp_mul(&num, &num, &t, &FR);
        fp_mul(&den, &beta, &s1[i], &FR); fp_add(&den, &den, &a[i], &FR); fp_add(&den, &den, &gamma, &FR);
        fp_mul(&u, &beta, &s2[i], &FR); fp_add(&u, &u, &b[i], &FR); fp_add(&u, &u, &gamma, &FR); fp_mul(&den, &den, &u, &FR);
        fp_mul(&u, &beta, &s3[i], &FR); fp_add(&u, &u, &c[i], &FR); fp_add(&u, &u, &gamma, &FR); fp_mul(&den, &den, &u, &FR);
        fp_inv(&den, &den, &FR);
        fp_mul(&prod[i], &num, &den, &FR);
    }
    fe state; fp_one(&state, &FR); out[0] = state;
    for (size_t i = 0; i + 1 < n; ++i) { fp_mul(&state, &state, &prod[i], &FR); out[i + 1] = state; }
    free(prod); free(roots);
}

/* lookup/mod.rs:25-82 */
API void zko_z2_evals(unsigned log_n, const u64 *delta_, const u64 *eps_, const u64 *f_, const u64 *t_,
                      const u64 *h1_, const u64 *h2_, u64 *out_) {
    size_t n = (size_t)1 << log_n;
    const fe *f = (const fe *)f_, *t = (const fe *)t_, *h1 = (const fe *)h1_, *h2 = (const fe *)h2_;
    fe delta = *(const fe *)delta_, eps = *(const fe *)eps_, *out = (fe *)out_;
    fe one, opd, eopd; fp_one(&one, &FR); fp_add(&opd, &one, &delta, &FR); fp_mul(&eopd, &eps, &opd, &FR);
    fe *prod = (fe *)malloc(n * sizeof(fe));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n - 1; ++i) {
        fe num, den, u;
        fp_add(&num, &eps, &f[i], &FR); fp_mul(&num, &opd, &num, &FR);
        fp_mul(&u, &delta, &t[i + 1], &FR); fp_add(&u, &u, &eopd, &FR); fp_add(&u, &u, &t[i], &FR); fp_mul(&num, &num, &u, &FR);
        fp_mul(&den, &delta, &h2[i], &FR); fp_add(&den, &den, &eopd, &FR); fp_add(&den, &den, &h1[i], &FR);
        fp_mul(&u, &delta, &h1[i + 1], &FR); fp_add(&u, &u, &eopd, &FR); fp_add(&u, &u, &h2[i], &FR); fp_mul(&den, &den, &u, &FR);
        fp_inv(&den, &den, &FR);
        fp_mul(&prod[i], &num, &den, &FR);
    }
    fe state = one; out[0] = state;
    for (size_t i = 0; i + 1 < n; ++i) { fp_mul(&state, &state, &prod[i], &FR); out[i + 1] = state; }
    free(prod);
}

/* ------------------------------------------------------------------ quotient on the 4n coset */
/* keys/mod.rs:109-119: x_coset, zh_coset, l_1_coset as definitions on 5*<w_4n>; each 4n elements. */
API void zko_epk_free_tables(unsigned log_n, u64 *x_, u64 *zh_, u64 *l1_) {
    size_t n = (size_t)1 << log_n, n4 = 4 * n;
    fe *xs = (fe *)x_, *zh = (fe *)zh_, *l1 = (fe *)l1_;
    fe w, g, one, ninv; fr_root_of_unity(&w, log_n + 2); fp_from_u64(&g, ZKO_FR_GENERATOR, &FR); fp_one(&one, &FR);
    fp_from_u64(&ninv, (u64)n, &FR); fp_inv(&ninv, &ninv, &FR);
    xs[0] = g;
    for (size_t i = 1; i < n4; ++i) fp_mul(&xs[i], &xs[i - 1], &w, &FR);
    u64 e[4] = {(u64)n, 0, 0, 0};
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n4; ++i) {
        fe t, d; fp_pow(&t, &xs[i], e, &FR); fp_sub(&zh[i], &t, &one, &FR);
        fp_sub(&d, &xs[i], &one, &FR); fp_inv(&d, &d, &FR);
        fp_mul(&t, &zh[i], &ninv, &FR); fp_mul(&l1[i], &t, &d, &FR);
    }
}

/* challenges: alpha,beta,gamma,delta,epsilon (5 x 4 limbs).
 * wit: 9 pointers z1,z2,a,b,c,pi,t,h1,h2 ; epk: 13 pointers q_m,q_l,q_r,q_o,q_c,q_lookup,q_table,
 * sigma1,sigma2,sigma3,x,l1,zh ; all 4n Montgomery Fr.  out: 4n quotient coset evaluations. */
API void zko_quotient_evals(unsigned log_n, const u64 *ch, const u64 *const *wit, const u64 *const *epk, u64 *out_) {
    size_t n4 = (size_t)4 << log_n;
    const fe *C = (const fe *)ch;
    fe al = C[0], be = C[1], ga = C[2], de = C[3], ep = C[4];
    fe one, al2, al3, al4, al5, opd, eopd, k1, k2;
    fp_one(&one, &FR); fp_sqr(&al2, &al, &FR); fp_mul(&al3, &al2, &al, &FR); fp_mul(&al4, &al3, &al, &FR); fp_mul(&al5, &al4, &al, &FR);
    fp_add(&opd, &de, &one, &FR); fp_mul(&eopd, &ep, &opd, &FR);
    fp_from_u64(&k1, 7, &FR); fp_from_u64(&k2, 13, &FR);
    const fe *z1 = (const fe *)wit[0], *z2 = (const fe *)wit[1], *a = (const fe *)wit[2], *b = (const fe *)wit[3],
             *c = (const fe *)wit[4], *pi = (const fe *)wit[5], *t = (const fe *)wit[6], *h1 = (const fe *)wit[7], *h2 = (const fe *)wit[8];
    const fe *qm = (const fe *)epk[0], *ql = (const fe *)epk[1], *qr = (const fe *)epk[2], *qo = (const fe *)epk[3],
             *qc = (const fe *)epk[4], *qlk = (const fe *)epk[5], *qt = (const fe *)epk[6], *s1 = (const fe *)epk[7],
             *s2 = (const fe *)epk[8], *s3 = (const fe *)epk[9], *xc = (const fe *)epk[10], *l1 = (const fe *)epk[11],
             *zh = (const fe *)epk[12];
    fe *out = (fe *)out_;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n4; ++i) {
        size_t j = (i + 4) % n4;
        fe ar, u, v, w_, p1, p2, p3, pm, lk, l_1, l_2, l_3, l_4;
        /* arithmetic (keys/arithmetic.rs:67-81) */
        fp_mul(&ar, &a[i], &b[i], &FR); fp_mul(&ar, &ar, &qm[i], &FR);
        fp_mul(&u, &a[i], &ql[i], &FR); fp_add(&ar, &ar, &u, &FR);
        fp_mul(&u, &b[i], &qr[i], &FR); fp_add(&ar, &ar, &u, &FR);
        fp_mul(&u, &c[i], &qo[i], &FR); fp_add(&ar, &ar, &u, &FR);
        fp_add(&ar, &ar, &qc[i], &FR); fp_add(&ar, &ar, &pi[i], &FR);
        /* permutation (keys/permutation.rs:97-137) */
        fe bx; fp_mul(&bx, &be, &xc[i], &FR);
        fp_add(&u, &bx, &a[i], &FR); fp_add(&u, &u, &ga, &FR);
        fp_mul(&v, &bx, &k1, &FR); fp_add(&v, &v, &b[i], &FR); fp_add(&v, &v, &ga, &FR);
        fp_mul(&w_, &bx, &k2, &FR); fp_add(&w_, &w_, &c[i], &FR); fp_add(&w_, &w_, &ga, &FR);
        fp_mul(&p1, &al, &z1[i], &FR); fp_mul(&p1, &p1, &u, &FR); fp_mul(&p1, &p1, &v, &FR); fp_mul(&p1, &p1, &w_, &FR);
        fp_mul(&u, &be, &s1[i], &FR); fp_add(&u, &u, &a[i], &FR); fp_add(&u, &u, &ga, &FR);
        fp_mul(&v, &be, &s2[i], &FR); fp_add(&v, &v, &b[i], &FR); fp_add(&v, &v, &ga, &FR);
        fp_mul(&w_, &be, &s3[i], &FR); fp_add(&w_, &w_, &c[i], &FR); fp_add(&w_, &w_, &ga, &FR);
        fp_mul(&p2, &al, &z1[j], &FR); fp_mul(&p2, &p2, &u, &FR); fp_mul(&p2, &p2, &v, &FR); fp_mul(&p2, &p2, &w_, &FR);
        fp_sub(&p3, &z1[i], &one, &FR); fp_mul(&p3, &p3, &l1[i], &FR); fp_mul(&p3, &p3, &al2, &FR);
        fp_sub(&pm, &p1, &p2, &FR); fp_add(&pm, &pm, &p3, &FR);
        /* lookup (keys/lookup.rs:81-122) */
        fp_mul(&u, &qlk[i], &c[i], &FR); fp_add(&u, &u, &ep, &FR);
        fp_mul(&v, &de, &t[j], &FR); fp_add(&v, &v, &eopd, &FR); fp_add(&v, &v, &t[i], &FR);
        fp_mul(&l_1, &al3, &z2[i], &FR); fp_mul(&l_1, &l_1, &opd, &FR); fp_mul(&l_1, &l_1, &u, &FR); fp_mul(&l_1, &l_1, &v, &FR);
        fp_mul(&u, &de, &h2[i], &FR); fp_add(&u, &u, &eopd, &FR); fp_add(&u, &u, &h1[i], &FR);
        fp_mul(&v, &de, &h1[j], &FR); fp_add(&v, &v, &eopd, &FR); fp_add(&v, &v, &h2[i], &FR);
        fp_mul(&l_2, &al3, &z2[j], &FR); fp_mul(&l_2, &l_2, &u, &FR); fp_mul(&l_2, &l_2, &v, &FR);
        fp_sub(&l_3, &z2[i], &one, &FR); fp_mul(&l_3, &l_3, &al4, &FR); fp_mul(&l_3, &l_3, &l1[i], &FR);
        fp_mul(&l_4, &al5, &qt[i], &FR); fp_mul(&l_4, &l_4, &t[i], &FR);
        fp_sub(&lk, &l_1, &l_2, &FR); fp_add(&lk, &lk, &l_3, &FR); fp_add(&lk, &lk, &l_4, &FR);
        /* (arith + perm + lookup) * zh^-1  (quotient_poly.rs:220-224) */
        fe sum, zi; fp_add(&sum, &ar, &pm, &FR); fp_add(&sum, &sum, &lk, &FR);
        fp_inv(&zi, &zh[i], &FR);
        fp_mul(&out[i], &sum, &zi, &FR);
    }
}

/* out = sum_k coeffs[k] z^k (Montgomery in and out) */
API void zko_poly_eval(const u64 *coeffs_, size_t n, const u64 *z_, u64 *out_) {
    const fe *c = (const fe *)coeffs_;
    fe z, acc;
    memcpy(z.l, z_, 32);
    memset(&acc, 0, sizeof acc);
    for (size_t k = n; k-- > 0;) {
        fp_mul(&acc, &acc, &z, &FR);
        fp_add(&acc, &acc, &c[k], &FR);
    }
    memcpy(out_, acc.l, 32);
}

/* out[i] = sum_j scalars[j] * polys[j][i] for i < out_len (polynomial j has lens[j] coefficients) */
API void zko_poly_lincomb(size_t k, const u64 *const *polys, const size_t *lens, const u64 *scalars_, u64 *out_, size_t out_len) {
    fe *out = (fe *)out_;
    const fe *sc = (const fe *)scalars_;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)out_len; ++i) {
        fe acc, t;
        memset(&acc, 0, sizeof acc);
        for (size_t j = 0; j < k; ++j) {
            if ((size_t)i >= lens[j]) continue;
            fp_mul(&t, &sc[j], &((const fe *)polys[j])[i], &FR);
            fp_add(&acc, &acc, &t, &FR);
        }
        out[i] = acc;
    }
}

/* quot (n - 1 coefficients) = (p(X) - p(z)) / (X - z) by synthetic division from the top; eval = p(z) */
API void zko_poly_divide_linear(const u64 *coeffs_, size_t n, const u64 *z_, u64 *quot_, u64 *eval_) {
    const fe *c = (const fe *)coeffs_;
    fe *q = (fe *)quot_;
    fe z, carry, t;
    memcpy(z.l, z_, 32);
    memset(&carry, 0, sizeof carry);
    for (size_t k = n; k-- > 1;) {
        fp_mul(&t, &z, &carry, &FR);
        fp_add(&carry, &c[k], &t, &FR);
        q[k - 1] = carry;
    }
    fp_mul(&t, &z, &carry, &FR);
    if (n) fp_add(&t, &t, &c[0], &FR);
    memcpy(eval_, t.l, 32);
}

API int zko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
