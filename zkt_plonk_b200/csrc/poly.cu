// poly.cu -- grand products, quotient evaluation and polynomial utilities over BN254 Fr for sm_100a.
//
// Replaces the prover's serial / rayon CPU loops:
//   zkb_z1_evals_dev        compute_z1_poly           plonk-core/src/permutation/mod.rs:181-254 (before the iFFT)
//   zkb_z2_evals_dev        compute_z2_poly           plonk-core/src/lookup/mod.rs:25-82       (before the iFFT)
//   zkb_quotient_evals_dev  quotient_poly::compute    plonk-core/src/proof_system/quotient_poly.rs:98-224 with
//                           keys/arithmetic.rs:67-81, keys/permutation.rs:97-137, keys/lookup.rs:81-122
//   zkb_poly_eval_dev       DensePolynomial::evaluate linearization_poly.rs:55-75 (12 openings)
//   zkb_poly_lincomb_dev    poly * scalar + ...       linearization_poly.rs:77-111, SonicKZG10::open's combination
//   zkb_poly_divide_linear_dev  kzg10::open witness   (p(X) - p(z)) / (X - z), prove.rs:381-451
//   zkb_poly_add_blinders_dev   add_blinders_to_poly  prove.rs:472-483
//
// The reference inverts one field element per row (n - 1 Fermat/EEA inversions) and multiplies the ratios in a
// serial loop.  Here a grand product is z[i+1] = prefix_prod(num)[i] * suffix_prod(den)[i+1] / prod(den): two
// parallel multiplicative scans and ONE inversion.  All field results are exact, so the outputs are identical.
#include "ctx.h"
#include "ff.cuh"

using namespace zkb;
typedef FrP F;

namespace {

// ============================================================================================ scans over Fr
constexpr int SC_K = 8;                       // elements per thread
constexpr int SC_T = 256;                     // threads per CTA
constexpr int SC_TILE = SC_K * SC_T;          // 2048 elements per CTA

template <int OP> __device__ __forceinline__ fe_t op_id() { return OP == 0 ? fone<F>() : fzero<F>(); }
template <int OP> __device__ __forceinline__ fe_t op_apply(const fe_t &a, const fe_t &b) {
    return OP == 0 ? fmul<F>(a, b) : fadd<F>(a, b);
}

__device__ __forceinline__ fe_t shfl_up_fe(const fe_t &a, int d) {
    fe_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], d);
    return r;
}
__device__ __forceinline__ fe_t shfl_idx_fe(const fe_t &a, int lane) {
    fe_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], lane);
    return r;
}

// Inclusive scan of one value per thread over a 256-thread CTA.  Returns the inclusive result; *excl gets the
// exclusive one; *total the CTA total.  sm: 8 fe_t.
template <int OP>
__device__ __forceinline__ fe_t block_scan(const fe_t &v, fe_t *sm, fe_t *excl, fe_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    fe_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        fe_t y = shfl_up_fe(x, d);
        if (lane >= d) x = op_apply<OP>(y, x);
    }
    fe_t prev = shfl_up_fe(x, 1);             // exclusive inside the warp
    if (lane == 0) prev = op_id<OP>();
    if (lane == 31) sm[wid] = x;
    __syncthreads();
    fe_t woff = op_id<OP>(), tot = op_id<OP>();
#pragma unroll
    for (int w = 0; w < SC_T / 32; ++w) {     // 8 warp totals: every thread folds them (7 ops, no second barrier)
        fe_t t = sm[w];
        if (w < wid) woff = op_apply<OP>(woff, t);
        tot = op_apply<OP>(tot, t);
    }
    __syncthreads();
    *excl = op_apply<OP>(woff, prev);
    *total = tot;
    return op_apply<OP>(woff, x);
}

// memory index of scan position e (REV: the scan runs from the top of the array downwards)
template <int REV> __device__ __forceinline__ size_t scan_index(size_t e, size_t n) { return REV ? n - 1 - e : e; }

template <int OP, int REV>
__global__ void __launch_bounds__(SC_T) fr_scan_reduce_kernel(const uint4 *in, size_t n, uint4 *tile_tot) {
    __shared__ fe_t sm[SC_T / 32];
    size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_K;
    fe_t acc = op_id<OP>();
#pragma unroll
    for (int k = 0; k < SC_K; ++k) {
        size_t e = base + k;
        if (e < n) acc = op_apply<OP>(acc, fload(in + 2 * scan_index<REV>(e, n)));
    }
    fe_t ex, tot;
    block_scan<OP>(acc, sm, &ex, &tot);
    if (threadIdx.x == 0) fstore(tile_tot + 2 * (size_t)blockIdx.x, tot);
}

// single CTA: exclusive scan of the tile totals in place; grand total -> tile_tot[ntiles]
template <int OP>
__global__ void __launch_bounds__(SC_T) fr_scan_tiles_kernel(uint4 *tile_tot, uint32_t ntiles) {
    __shared__ fe_t sm[SC_T / 32];
    fe_t running = op_id<OP>();
    for (uint32_t base = 0; base < ntiles; base += SC_T) {
        uint32_t i = base + threadIdx.x;
        fe_t v = i < ntiles ? fload(tile_tot + 2 * (size_t)i) : op_id<OP>();
        fe_t ex, tot;
        block_scan<OP>(v, sm, &ex, &tot);
        if (i < ntiles) fstore(tile_tot + 2 * (size_t)i, op_apply<OP>(running, ex));
        running = op_apply<OP>(running, tot);
    }
    if (threadIdx.x == 0) fstore(tile_tot + 2 * (size_t)ntiles, running);
}

template <int OP, int REV>
__global__ void __launch_bounds__(SC_T) fr_scan_apply_kernel(const uint4 *in, size_t n, const uint4 *tile_tot, uint4 *out) {
    __shared__ fe_t sm[SC_T / 32];
    size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_K;
    fe_t v[SC_K];
    fe_t acc = op_id<OP>();
#pragma unroll
    for (int k = 0; k < SC_K; ++k) {
        size_t e = base + k;
        v[k] = e < n ? fload(in + 2 * scan_index<REV>(e, n)) : op_id<OP>();
        acc = op_apply<OP>(acc, v[k]);
        v[k] = acc;                              // inclusive inside the thread
    }
    fe_t ex, tot;
    block_scan<OP>(acc, sm, &ex, &tot);
    fe_t carry = op_apply<OP>(fload(tile_tot + 2 * (size_t)blockIdx.x), ex);
#pragma unroll
    for (int k = 0; k < SC_K; ++k) {
        size_t e = base + k;
        if (e < n) fstore(out + 2 * scan_index<REV>(e, n), op_apply<OP>(carry, v[k]));
    }
}

// inclusive scan (in place allowed); tile_tot needs ntiles + 1 elements; the grand total lands in tile_tot[ntiles]
template <int OP, int REV>
int fr_scan(zkb_ctx *ctx, const uint4 *in, uint4 *out, size_t n, uint4 *tile_tot) {
    uint32_t ntiles = (uint32_t)((n + SC_TILE - 1) / SC_TILE);
    fr_scan_reduce_kernel<OP, REV><<<ntiles, SC_T, 0, ctx->stream>>>(in, n, tile_tot);
    fr_scan_tiles_kernel<OP><<<1, SC_T, 0, ctx->stream>>>(tile_tot, ntiles);
    fr_scan_apply_kernel<OP, REV><<<ntiles, SC_T, 0, ctx->stream>>>(in, n, tile_tot, out);
    ctx->launches += 3;
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

__device__ __forceinline__ fe_t pow2lvl(const uint4 *tab, uint32_t s, unsigned long long e) {
    uint32_t lo = (uint32_t)(e & ((1ull << s) - 1));
    unsigned long long hi = e >> s;
    return fmul<F>(fload_ro(tab + 2 * (size_t)lo), fload_ro(tab + 2 * ((size_t)(1u << s) + hi)));
}

// ============================================================================================ grand products
struct Z1Args {
    const uint4 *a, *b, *c, *s1, *s2, *s3;
    const uint4 *wtab;        // two-level table of the domain generator w_n
    uint32_t wtab_s;
    uint32_t n;
    fe_t beta, gamma, k1beta, k2beta;
};

// rows i < n-1: num_i, den_i of permutation/mod.rs:232-243; row n-1: (1, 1)
__global__ void __launch_bounds__(256) z1_terms_kernel(Z1Args p, uint4 *num, uint4 *den) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    if (i == p.n - 1) { fstore(num + 2 * (size_t)i, fone<F>()); fstore(den + 2 * (size_t)i, fone<F>()); return; }
    fe_t root = pow2lvl(p.wtab, p.wtab_s, i);
    fe_t a = fload_ro(p.a + 2 * (size_t)i), b = fload_ro(p.b + 2 * (size_t)i), c = fload_ro(p.c + 2 * (size_t)i);
    fe_t ag = fadd<F>(a, p.gamma), bg = fadd<F>(b, p.gamma), cg = fadd<F>(c, p.gamma);
    fe_t n0 = fadd<F>(fmul<F>(p.beta, root), ag);
    fe_t n1 = fadd<F>(fmul<F>(p.k1beta, root), bg);
    fe_t n2 = fadd<F>(fmul<F>(p.k2beta, root), cg);
    fe_t d0 = fadd<F>(fmul<F>(p.beta, fload_ro(p.s1 + 2 * (size_t)i)), ag);
    fe_t d1 = fadd<F>(fmul<F>(p.beta, fload_ro(p.s2 + 2 * (size_t)i)), bg);
    fe_t d2 = fadd<F>(fmul<F>(p.beta, fload_ro(p.s3 + 2 * (size_t)i)), cg);
    fstore(num + 2 * (size_t)i, fmul<F>(fmul<F>(n0, n1), n2));
    fstore(den + 2 * (size_t)i, fmul<F>(fmul<F>(d0, d1), d2));
}

struct Z2Args {
    const uint4 *f, *t, *h1, *h2;
    uint32_t n;
    fe_t delta, eps, opd, eopd;    // delta, epsilon, 1 + delta, epsilon * (1 + delta)
};

// rows i < n-1: num_i, den_i of lookup/mod.rs:63-73; row n-1: (1, 1)
__global__ void __launch_bounds__(256) z2_terms_kernel(Z2Args p, uint4 *num, uint4 *den) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    if (i == p.n - 1) { fstore(num + 2 * (size_t)i, fone<F>()); fstore(den + 2 * (size_t)i, fone<F>()); return; }
    fe_t f = fload_ro(p.f + 2 * (size_t)i), t = fload_ro(p.t + 2 * (size_t)i), tn = fload_ro(p.t + 2 * (size_t)(i + 1));
    fe_t h1 = fload_ro(p.h1 + 2 * (size_t)i), h1n = fload_ro(p.h1 + 2 * (size_t)(i + 1)), h2 = fload_ro(p.h2 + 2 * (size_t)i);
    fe_t n0 = fmul<F>(p.opd, fadd<F>(p.eps, f));
    fe_t n1 = fadd<F>(fadd<F>(fmul<F>(p.delta, tn), p.eopd), t);
    fe_t d0 = fadd<F>(fadd<F>(fmul<F>(p.delta, h2), p.eopd), h1);
    fe_t d1 = fadd<F>(fadd<F>(fmul<F>(p.delta, h1n), p.eopd), h2);
    fstore(num + 2 * (size_t)i, fmul<F>(n0, n1));
    fstore(den + 2 * (size_t)i, fmul<F>(d0, d1));
}

// z[0] = 1; z[i+1] = prefix_num[i] * suffix_den[i+1] * total_den^-1
__global__ void __launch_bounds__(256) grand_product_combine_kernel(const uint4 *pnum, const uint4 *sden, const fe_t tinv,
                                                                    uint32_t n, uint4 *z) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) { fstore(z, fone<F>()); return; }
    fe_t v = fmul<F>(fload_ro(pnum + 2 * (size_t)(i - 1)), fload_ro(sden + 2 * (size_t)i));
    fstore(z + 2 * (size_t)i, fmul<F>(v, tinv));
}

struct GpWs { uint4 *num, *den, *tiles_a, *tiles_b; };

int gp_workspace(zkb_ctx *ctx, size_t n, GpWs &w) {
    size_t ntiles = (n + SC_TILE - 1) / SC_TILE + 2;
    size_t bytes = 2 * n * 32 + 2 * ntiles * 32;
    int rc = zkb_reserve(ctx, ctx->poly_ws, bytes);
    if (rc) return rc;
    char *p = (char *)ctx->poly_ws.p;
    w.num = (uint4 *)p; p += n * 32;
    w.den = (uint4 *)p; p += n * 32;
    w.tiles_a = (uint4 *)p; p += ntiles * 32;
    w.tiles_b = (uint4 *)p;
    return ZKB_OK;
}

fe_t dev_fe(const host::Fe &f);

// shared tail of z1 / z2 once num and den are materialised.  The one inversion of a grand product (of the total of the
// denominators) is done on the HOST: a lone GPU thread needs 151 us for the Fermat chain (profiles/r02t_trace_2^18.jsonl),
// the 32-byte round trip and the host's chain ~35 us, and the zero-denominator check comes with it.
int grand_product_finish(zkb_ctx *ctx, const GpWs &w, size_t n, uint4 *out) {
    uint32_t ntiles = (uint32_t)((n + SC_TILE - 1) / SC_TILE);
    int rc = fr_scan<0, 0>(ctx, w.num, w.num, n, w.tiles_a);         // inclusive prefix products of num
    if (rc) return rc;
    rc = fr_scan<0, 1>(ctx, w.den, w.den, n, w.tiles_b);             // inclusive suffix products of den
    if (rc) return rc;
    host::Fe total;
    ZKB_CUDA(ctx, cudaMemcpyAsync(total.l, w.tiles_b + 2 * (size_t)ntiles, 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->gp_failed = (total.l[0] | total.l[1] | total.l[2] | total.l[3]) == 0;   // the reference panics here (inverse().unwrap())
    const host::Fe tinv = ctx->gp_failed ? total : host::inv(total, host::FR);
    grand_product_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(w.num, w.den, dev_fe(tinv), (uint32_t)n, out);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

fe_t dev_fe(const host::Fe &f) { fe_t r; memcpy(r.v, f.l, 32); return r; }
host::Fe host_fe(const uint64_t *p) { host::Fe f; memcpy(f.l, p, 32); return f; }

// ============================================================================================ quotient
struct QuotArgs {
    const uint4 *z1, *z2, *a, *b, *c, *pi, *t, *h1, *h2;                               // witness cosets (4n)
    const uint4 *qm, *ql, *qr, *qo, *qc, *qlk, *qt, *s1, *s2, *s3, *l1;                // static epk cosets (4n)
    const uint4 *xtab;                                                                 // two-level table of w_4n
    uint32_t xtab_s, n4;
    uint32_t i_lo, i_hi;                                                               // elements [i_lo, i_hi) of the coset (a slice when several GPUs share the round)
    fe_t alpha, alpha2, alpha3, alpha4, alpha5, beta, gamma, delta, eps, opd, eopd, gen, k1, k2;
    fe_t a3opd;                                                                        // alpha^3 (1 + delta)
    fe_t zh_inv[4];                                                                    // 1 / zh on the coset, by i mod 4
};

// ZKB_QUOT_FUSED: the same values with fewer multiplier slots -- sums of products share ONE Montgomery reduction
// (fmaddn), the two L1 terms share a chain, K1 = 7 / K2 = 13 are addition chains on the ALU pipe, alpha^3 (1 + delta) and
// the coset generator come pre-multiplied from the host (the x table's high half is scaled by g).  Every value is the
// canonical residue of the same field expression, so the output is bit-identical to the plain kernel below.
// Measured on B200 (profiles/r01j): 2.71 -> 2.06 ms at n = 2^20 (4 Mi elements), same SHA-256 of the output, polynomial
// and prover suites bit-exact.  -DZKB_QUOT_FUSED=0 restores the plain kernel for A/B measurements.
#ifndef ZKB_QUOT_FUSED
#define ZKB_QUOT_FUSED 1
#endif
#if ZKB_QUOT_FUSED
__global__ void __launch_bounds__(128) quotient_kernel(const __grid_constant__ QuotArgs p, uint4 *out) {
    uint32_t i = p.i_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.i_hi) return;
    uint32_t j = i + 4 < p.n4 ? i + 4 : i + 4 - p.n4;           // "next": x * w_n  (quotient_poly.rs:53-94)
    const size_t I = 2 * (size_t)i, J = 2 * (size_t)j;
    fe_t a = fload_ro(p.a + I), b = fload_ro(p.b + I), c = fload_ro(p.c + I);
    // ---- arithmetic gate (keys/arithmetic.rs:67-81): (a b) q_m + a q_l + b q_r + c q_o under one reduction
    fe_t acc;
    {
        const fe_t u[4] = {fmul<F>(a, b), a, b, c};
        const fe_t v[4] = {fload_ro(p.qm + I), fload_ro(p.ql + I), fload_ro(p.qr + I), fload_ro(p.qo + I)};
        acc = fmaddn<F, 4>(u, v);
    }
    acc = fadd<F>(acc, fload_ro(p.qc + I));
    acc = fadd<F>(acc, fload_ro(p.pi + I));
    const fe_t l1 = fload_ro(p.l1 + I), z2 = fload_ro(p.z2 + I);
    // ---- permutation (keys/permutation.rs:97-137) + both L1 terms: alpha^2 L1 (z1 - 1) + alpha^4 L1 (z2 - 1)
    {
        fe_t bx = fmul<F>(p.beta, pow2lvl(p.xtab, p.xtab_s, i));          // beta * g * w_4n^i (g sits in the table)
        fe_t ag = fadd<F>(a, p.gamma), bg = fadd<F>(b, p.gamma), cg = fadd<F>(c, p.gamma);
        fe_t z1 = fload_ro(p.z1 + I), z1n = fload_ro(p.z1 + J);
        fe_t t0 = fmul<F>(p.alpha, z1);
        t0 = fmul<F>(t0, fadd<F>(bx, ag));
        t0 = fmul<F>(t0, fadd<F>(fmul_small<F>(bx, 7), bg));
        fe_t f0 = fadd<F>(fmul_small<F>(bx, 13), cg);
        fe_t t1 = fmul<F>(p.alpha, z1n);
        t1 = fmul<F>(t1, fadd<F>(fmul<F>(p.beta, fload_ro(p.s1 + I)), ag));
        t1 = fmul<F>(t1, fadd<F>(fmul<F>(p.beta, fload_ro(p.s2 + I)), bg));
        fe_t f1 = fadd<F>(fmul<F>(p.beta, fload_ro(p.s3 + I)), cg);
        fe_t inner = fmadd2<F>(fsub<F>(z1, fone<F>()), p.alpha2, fsub<F>(z2, fone<F>()), p.alpha4);
        const fe_t u[3] = {t0, fneg<F>(t1), inner};
        const fe_t v[3] = {f0, f1, l1};
        acc = fadd<F>(acc, fmaddn<F, 3>(u, v));
    }
    // ---- lookup (keys/lookup.rs:81-122)
    {
        fe_t t = fload_ro(p.t + I), tn = fload_ro(p.t + J), h1 = fload_ro(p.h1 + I), h1n = fload_ro(p.h1 + J);
        fe_t h2 = fload_ro(p.h2 + I), z2n = fload_ro(p.z2 + J);
        fe_t u0 = fadd<F>(fmul<F>(fload_ro(p.qlk + I), c), p.eps);
        fe_t v0 = fadd<F>(fadd<F>(fmul<F>(p.delta, tn), p.eopd), t);
        fe_t t0 = fmul<F>(fmul<F>(p.a3opd, z2), u0);
        fe_t u1 = fadd<F>(fadd<F>(fmul<F>(p.delta, h2), p.eopd), h1);
        fe_t v1 = fadd<F>(fadd<F>(fmul<F>(p.delta, h1n), p.eopd), h2);
        fe_t t1 = fmul<F>(fmul<F>(p.alpha3, z2n), u1);
        fe_t t3 = fmul<F>(p.alpha5, fload_ro(p.qt + I));
        const fe_t u[3] = {t0, fneg<F>(t1), t3};
        const fe_t v[3] = {v0, v1, t};
        acc = fadd<F>(acc, fmaddn<F, 3>(u, v));
    }
    // ---- divide by the vanishing polynomial: zh takes 4 values on the 4n coset (quotient_poly.rs:220-224)
    fstore(out + I, fmul<F>(acc, p.zh_inv[i & 3]));
}
#else
__global__ void __launch_bounds__(128) quotient_kernel(const __grid_constant__ QuotArgs p, uint4 *out) {
    uint32_t i = p.i_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.i_hi) return;
    uint32_t j = i + 4 < p.n4 ? i + 4 : i + 4 - p.n4;           // "next": x * w_n  (quotient_poly.rs:53-94)
    const size_t I = 2 * (size_t)i, J = 2 * (size_t)j;
    fe_t a = fload_ro(p.a + I), b = fload_ro(p.b + I), c = fload_ro(p.c + I);
    // ---- arithmetic gate (keys/arithmetic.rs:67-81)
    fe_t acc = fmul<F>(fmul<F>(a, b), fload_ro(p.qm + I));
    acc = fadd<F>(acc, fmul<F>(a, fload_ro(p.ql + I)));
    acc = fadd<F>(acc, fmul<F>(b, fload_ro(p.qr + I)));
    acc = fadd<F>(acc, fmul<F>(c, fload_ro(p.qo + I)));
    acc = fadd<F>(acc, fload_ro(p.qc + I));
    acc = fadd<F>(acc, fload_ro(p.pi + I));
    // ---- permutation (keys/permutation.rs:97-137); x_i = g * w_4n^i computed instead of streamed
    fe_t l1 = fload_ro(p.l1 + I);
    {
        fe_t x = fmul<F>(p.gen, pow2lvl(p.xtab, p.xtab_s, i));
        fe_t bx = fmul<F>(p.beta, x);
        fe_t ag = fadd<F>(a, p.gamma), bg = fadd<F>(b, p.gamma), cg = fadd<F>(c, p.gamma);
        fe_t z1 = fload_ro(p.z1 + I), z1n = fload_ro(p.z1 + J);
        fe_t t0 = fmul<F>(p.alpha, z1);
        t0 = fmul<F>(t0, fadd<F>(bx, ag));
        t0 = fmul<F>(t0, fadd<F>(fmul<F>(bx, p.k1), bg));
        t0 = fmul<F>(t0, fadd<F>(fmul<F>(bx, p.k2), cg));
        fe_t t1 = fmul<F>(p.alpha, z1n);
        t1 = fmul<F>(t1, fadd<F>(fmul<F>(p.beta, fload_ro(p.s1 + I)), ag));
        t1 = fmul<F>(t1, fadd<F>(fmul<F>(p.beta, fload_ro(p.s2 + I)), bg));
        t1 = fmul<F>(t1, fadd<F>(fmul<F>(p.beta, fload_ro(p.s3 + I)), cg));
        fe_t t2 = fmul<F>(fmul<F>(fsub<F>(z1, fone<F>()), l1), p.alpha2);
        acc = fadd<F>(acc, fadd<F>(fsub<F>(t0, t1), t2));
    }
    // ---- lookup (keys/lookup.rs:81-122)
    {
        fe_t t = fload_ro(p.t + I), tn = fload_ro(p.t + J), h1 = fload_ro(p.h1 + I), h1n = fload_ro(p.h1 + J);
        fe_t h2 = fload_ro(p.h2 + I), z2 = fload_ro(p.z2 + I), z2n = fload_ro(p.z2 + J);
        fe_t u = fadd<F>(fmul<F>(fload_ro(p.qlk + I), c), p.eps);
        fe_t v = fadd<F>(fadd<F>(fmul<F>(p.delta, tn), p.eopd), t);
        fe_t t0 = fmul<F>(fmul<F>(fmul<F>(fmul<F>(p.alpha3, z2), p.opd), u), v);
        u = fadd<F>(fadd<F>(fmul<F>(p.delta, h2), p.eopd), h1);
        v = fadd<F>(fadd<F>(fmul<F>(p.delta, h1n), p.eopd), h2);
        fe_t t1 = fmul<F>(fmul<F>(fmul<F>(p.alpha3, z2n), u), v);
        fe_t t2 = fmul<F>(fmul<F>(fsub<F>(z2, fone<F>()), p.alpha4), l1);
        fe_t t3 = fmul<F>(fmul<F>(p.alpha5, fload_ro(p.qt + I)), t);
        acc = fadd<F>(acc, fadd<F>(fadd<F>(fsub<F>(t0, t1), t2), t3));
    }
    // ---- divide by the vanishing polynomial: zh takes 4 values on the 4n coset (quotient_poly.rs:220-224)
    fstore(out + I, fmul<F>(acc, p.zh_inv[i & 3]));
}
#endif  // ZKB_QUOT_FUSED

// ============================================================================================ polynomial utilities
constexpr int EV_K = 16;                       // coefficients per thread in poly_eval

// partial[b] = sum over the CTA's chunk of c_k z^k
__global__ void __launch_bounds__(256) poly_eval_partial_kernel(const uint4 *coeffs, size_t n, fe_t z, const uint4 *ztab,
                                                                uint32_t ztab_s, uint4 *partial) {
    __shared__ fe_t sm[8];
    size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) * EV_K;
    fe_t acc = fzero<F>();
    if (base < n) {
        int cnt = (int)(n - base < (size_t)EV_K ? n - base : (size_t)EV_K);
        for (int k = cnt - 1; k >= 0; --k) acc = fadd<F>(fmul<F>(acc, z), fload_ro(coeffs + 2 * (base + k)));   // Horner
        acc = fmul<F>(acc, pow2lvl(ztab, ztab_s, base));
    }
    // CTA sum
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        fe_t o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], d);
        acc = fadd<F>(acc, o);
    }
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        fe_t s = sm[0];
        for (int w = 1; w < 8; ++w) s = fadd<F>(s, sm[w]);
        fstore(partial + 2 * (size_t)blockIdx.x, s);
    }
}

__global__ void __launch_bounds__(256) fr_sum_kernel(const uint4 *in, uint32_t n, uint4 *out) {
    __shared__ fe_t sm[8];
    fe_t acc = fzero<F>();
    for (uint32_t i = threadIdx.x; i < n; i += 256) acc = fadd<F>(acc, fload(in + 2 * (size_t)i));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        fe_t o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], d);
        acc = fadd<F>(acc, o);
    }
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        fe_t s = sm[0];
        for (int w = 1; w < 8; ++w) s = fadd<F>(s, sm[w]);
        fstore(out, s);
    }
}

constexpr int LC_MAX = 16;
struct LcArgs {
    const uint4 *p[LC_MAX];
    unsigned long long len[LC_MAX];
    fe_t s[LC_MAX];
    uint32_t k;
};

// Several polynomials at once (the 12 openings of round 5 share two points): blockIdx.y selects the polynomial, its point
// and its table; partial sums of polynomial k start at a.first[k].
struct EvArgs {
    const uint4 *p[LC_MAX];
    unsigned long long len[LC_MAX];
    fe_t z[LC_MAX];
    uint32_t tab[LC_MAX];      // offset (in field elements) of the point's two-level table
    uint32_t first[LC_MAX + 1];
    uint32_t s;
};

__global__ void __launch_bounds__(256) poly_eval_many_partial_kernel(const __grid_constant__ EvArgs a, const uint4 *tabs, uint4 *partial) {
    __shared__ fe_t sm[8];
    const uint32_t k = blockIdx.y;
    const size_t n = a.len[k];
    if ((size_t)blockIdx.x * 256 * EV_K >= n) return;                  // uniform over the CTA
    const uint4 *coeffs = a.p[k];
    const fe_t z = a.z[k];
    size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) * EV_K;
    fe_t acc = fzero<F>();
    if (base < n) {
        int cnt = (int)(n - base < (size_t)EV_K ? n - base : (size_t)EV_K);
        for (int c = cnt - 1; c >= 0; --c) acc = fadd<F>(fmul<F>(acc, z), fload_ro(coeffs + 2 * (base + c)));   // Horner
        acc = fmul<F>(acc, pow2lvl(tabs + 2 * (size_t)a.tab[k], a.s, base));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        fe_t o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], d);
        acc = fadd<F>(acc, o);
    }
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        fe_t t = sm[0];
        for (int w = 1; w < 8; ++w) t = fadd<F>(t, sm[w]);
        fstore(partial + 2 * ((size_t)a.first[k] + blockIdx.x), t);
    }
}

// out[k] = sum of partial[first[k] .. first[k+1])
__global__ void __launch_bounds__(256) fr_sum_many_kernel(const __grid_constant__ EvArgs a, const uint4 *partial, uint4 *out) {
    __shared__ fe_t sm[8];
    const uint32_t k = blockIdx.x, lo = a.first[k], hi = a.first[k + 1];
    fe_t acc = fzero<F>();
    for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) acc = fadd<F>(acc, fload(partial + 2 * (size_t)i));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        fe_t o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], d);
        acc = fadd<F>(acc, o);
    }
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        fe_t t = sm[0];
        for (int w = 1; w < 8; ++w) t = fadd<F>(t, sm[w]);
        fstore(out + 2 * (size_t)k, t);
    }
}

// out[i] = sum_k s_k * p_k[i]  (coefficients past a polynomial's length are zero)
__global__ void __launch_bounds__(256) poly_lincomb_kernel(const __grid_constant__ LcArgs a, uint4 *out, size_t out_len) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_len) return;
    fe_t acc = fzero<F>();
    for (uint32_t k = 0; k < a.k; ++k)
        if (i < a.len[k]) acc = fadd<F>(acc, fmul<F>(a.s[k], fload_ro(a.p[k] + 2 * i)));
    fstore(out + 2 * i, acc);
}

// t_k = p_k * z^k
__global__ void __launch_bounds__(256) scale_by_powers_kernel(const uint4 *in, size_t n, const uint4 *tab, uint32_t tab_s, uint4 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fstore(out + 2 * i, fmul<F>(fload_ro(in + 2 * i), pow2lvl(tab, tab_s, i)));
}

// quotient coefficient j-1 = z^-j * S_j for j = 1..n-1, where S is the inclusive suffix sum of p_k z^k
__global__ void __launch_bounds__(256) divide_finish_kernel(const uint4 *suffix, size_t n, const uint4 *itab, uint32_t itab_s, uint4 *quot) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (j >= n) return;
    fstore(quot + 2 * (j - 1), fmul<F>(fload_ro(suffix + 2 * j), pow2lvl(itab, itab_s, j)));
}

// quotient for z = 0: (p(X) - p(0)) / X is a shift
__global__ void __launch_bounds__(256) shift_down_kernel(const uint4 *in, size_t n, uint4 *quot) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (j >= n) return;
    fstore(quot + 2 * (j - 1), fload_ro(in + 2 * j));
}

// add_blinders_to_poly (prove.rs:472-483) for a polynomial whose buffer has room for len + k coefficients.  The reference
// FIRST extends the coefficient vector by the k blinders and THEN subtracts b_i from coeffs[i] for every i < k, so for
// len < k the subtraction reaches blinders that were just appended (len = 0: [b0-b0, b1-b1, b2-b2] = the zero polynomial;
// len = 1: [c0-b0, b0-b1, b1-b2, b2]).  One CTA: the barrier orders the two steps.
__global__ void add_blinders_kernel(uint4 *coeffs, size_t len, const __grid_constant__ LcArgs b) {
    uint32_t i = threadIdx.x;
    if (i < b.k) fstore(coeffs + 2 * (len + i), b.s[i]);
    __syncthreads();
    if (i < b.k) fstore(coeffs + 2 * (size_t)i, fsub<F>(fload(coeffs + 2 * (size_t)i), b.s[i]));
}

__global__ void fr_fill_kernel(uint4 *out, size_t n, fe_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fstore(out + 2 * i, v);
}

// out[0] = 1 + index of the highest non-zero coefficient (0 for the zero polynomial).  One atomic per CTA (and none
// when the CTA cannot raise the current maximum): a per-element atomicMax on one address serialises.
__global__ void __launch_bounds__(256) effective_len_kernel(const uint4 *coeffs, size_t n, unsigned long long *out) {
    __shared__ unsigned long long best[8];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (i < n) {
        uint4 a = coeffs[2 * i], b = coeffs[2 * i + 1];
        if (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) mine = (unsigned long long)(i + 1);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        unsigned long long o = __shfl_down_sync(0xffffffffu, mine, d);
        mine = o > mine ? o : mine;
    }
    if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = best[0];
        for (int w = 1; w < 8; ++w) m = best[w] > m ? best[w] : m;
        if (m) atomicMax(out, m);
    }
}

unsigned ceil_log2_sz(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) ++l; return l; }

}  // namespace

extern "C" {

int zkb_z1_evals_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t *a,
                     const uint64_t *b, const uint64_t *c, const uint64_t *sigma1, const uint64_t *sigma2,
                     const uint64_t *sigma3, uint64_t *out_dev) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!beta || !gamma || !a || !b || !c || !sigma1 || !sigma2 || !sigma3 || !out_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_z1_evals_dev: null argument");
    if (log_n > host::FR_TWO_ADICITY || log_n > 31) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_z1_evals_dev: log_n exceeds Fr TWO_ADICITY (28 on BN254) or 31");
    const size_t n = (size_t)1 << log_n;
    GpWs w;
    int rc = gp_workspace(ctx, n, w);
    if (rc) return rc;
    Z1Args p;
    p.a = (const uint4 *)a; p.b = (const uint4 *)b; p.c = (const uint4 *)c;
    p.s1 = (const uint4 *)sigma1; p.s2 = (const uint4 *)sigma2; p.s3 = (const uint4 *)sigma3;
    p.n = (uint32_t)n;
    host::Fe hb = host_fe(beta), hg = host_fe(gamma);
    p.beta = dev_fe(hb); p.gamma = dev_fe(hg);
    p.k1beta = dev_fe(host::mul(host::from_u64(7, host::FR), hb, host::FR));     // K1 = 7, K2 = 13 (permutation/constants.rs)
    p.k2beta = dev_fe(host::mul(host::from_u64(13, host::FR), hb, host::FR));
    const void *tab;
    rc = zkb_pow2lvl_cached(ctx, (2ull << 32) | (log_n << 1), log_n, host::fr_root_of_unity(log_n), host::one(host::FR), &tab, &p.wtab_s);
    if (rc) return rc;
    p.wtab = (const uint4 *)tab;
    z1_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, w.num, w.den);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return grand_product_finish(ctx, w, n, (uint4 *)out_dev);
}

int zkb_z2_evals_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t delta[4], const uint64_t epsilon[4], const uint64_t *f,
                     const uint64_t *t, const uint64_t *h1, const uint64_t *h2, uint64_t *out_dev) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!delta || !epsilon || !f || !t || !h1 || !h2 || !out_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_z2_evals_dev: null argument");
    if (log_n > host::FR_TWO_ADICITY || log_n > 31) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_z2_evals_dev: log_n exceeds Fr TWO_ADICITY (28 on BN254) or 31");
    const size_t n = (size_t)1 << log_n;
    GpWs w;
    int rc = gp_workspace(ctx, n, w);
    if (rc) return rc;
    Z2Args p;
    p.f = (const uint4 *)f; p.t = (const uint4 *)t; p.h1 = (const uint4 *)h1; p.h2 = (const uint4 *)h2;
    p.n = (uint32_t)n;
    host::Fe hd = host_fe(delta), he = host_fe(epsilon);
    host::Fe opd = host::add(host::one(host::FR), hd, host::FR);
    p.delta = dev_fe(hd); p.eps = dev_fe(he); p.opd = dev_fe(opd); p.eopd = dev_fe(host::mul(he, opd, host::FR));
    z2_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, w.num, w.den);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return grand_product_finish(ctx, w, n, (uint4 *)out_dev);
}

// 1 if the last grand product on this context met a zero denominator (the reference would have panicked)
int zkb_grand_product_failed(zkb_ctx *ctx) { return ctx ? ctx->gp_failed : 0; }

int zkb_quotient_evals_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t challenges[20], const uint64_t *const wit[9],
                           const uint64_t *const epk[11], uint64_t *out_dev) {
    return zkb_quotient_evals_range_dev(ctx, log_n, challenges, wit, epk, out_dev, 0, (size_t)4 << log_n);
}

// elements [lo, hi) only: the inputs must be valid on [lo, hi + 4) (cyclically) -- the slice a rank evaluates when the
// ranks of a box share round 4
int zkb_quotient_evals_range_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t challenges[20], const uint64_t *const wit[9],
                                 const uint64_t *const epk[11], uint64_t *out_dev, size_t lo, size_t hi) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!challenges || !wit || !epk || !out_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_quotient_evals_dev: null argument");
    if (log_n + 2 > host::FR_TWO_ADICITY || log_n + 2 > 31) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_quotient_evals_dev: 4n exceeds 2^28 (InvalidEvalDomainSize)");
    if (log_n < 3) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_quotient_evals_dev: n >= 8 required (quotient_poly.rs:44 asserts n >= 5)");
    for (int k = 0; k < 9; ++k) if (!wit[k]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_quotient_evals_dev: null witness table");
    for (int k = 0; k < 11; ++k) if (!epk[k]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_quotient_evals_dev: null key table");
    const size_t n = (size_t)1 << log_n, n4 = 4 * n;
    QuotArgs p;
    const uint4 **wp[9] = {&p.z1, &p.z2, &p.a, &p.b, &p.c, &p.pi, &p.t, &p.h1, &p.h2};
    for (int k = 0; k < 9; ++k) *wp[k] = (const uint4 *)wit[k];
    const uint4 **ep[11] = {&p.qm, &p.ql, &p.qr, &p.qo, &p.qc, &p.qlk, &p.qt, &p.s1, &p.s2, &p.s3, &p.l1};
    for (int k = 0; k < 11; ++k) *ep[k] = (const uint4 *)epk[k];
    p.n4 = (uint32_t)n4;
    if (lo > hi || hi > n4) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_quotient_evals_range_dev: bad element range");
    p.i_lo = (uint32_t)lo; p.i_hi = (uint32_t)hi;
    using namespace host;
    Fe al = host_fe(challenges), be = host_fe(challenges + 4), ga = host_fe(challenges + 8), de = host_fe(challenges + 12),
       ep_ = host_fe(challenges + 16);
    Fe al2 = sqr(al, FR), al3 = mul(al2, al, FR), al4 = mul(al3, al, FR), al5 = mul(al4, al, FR);
    Fe opd = add(one(FR), de, FR), eopd = mul(ep_, opd, FR), g = from_u64(host::FR_GENERATOR, FR);
    p.alpha = dev_fe(al); p.alpha2 = dev_fe(al2); p.alpha3 = dev_fe(al3); p.alpha4 = dev_fe(al4); p.alpha5 = dev_fe(al5);
    p.beta = dev_fe(be); p.gamma = dev_fe(ga); p.delta = dev_fe(de); p.eps = dev_fe(ep_); p.opd = dev_fe(opd); p.eopd = dev_fe(eopd);
    p.gen = dev_fe(g); p.k1 = dev_fe(from_u64(7, FR)); p.k2 = dev_fe(from_u64(13, FR));
    p.a3opd = dev_fe(mul(al3, opd, FR));
    // zh(x_i) = g^n * (w_4n^n)^(i mod 4) - 1 : four values (keys/mod.rs:114-116 evaluates x^n - 1 on the coset)
    Fe gn = pow_u64(g, (uint64_t)n, FR), w4 = pow_u64(fr_root_of_unity(log_n + 2), (uint64_t)n, FR), cur = gn;
    for (int k = 0; k < 4; ++k) {
        p.zh_inv[k] = dev_fe(inv(sub(cur, one(FR), FR), FR));
        cur = mul(cur, w4, FR);
    }
    const void *tab;
#if ZKB_QUOT_FUSED
    int rc = zkb_pow2lvl_cached(ctx, (9ull << 32) | ((log_n + 2) << 1), log_n + 2, fr_root_of_unity(log_n + 2), g, &tab, &p.xtab_s);   // g * w^e
#else
    int rc = zkb_pow2lvl_cached(ctx, (2ull << 32) | ((log_n + 2) << 1), log_n + 2, fr_root_of_unity(log_n + 2), one(FR), &tab, &p.xtab_s);
#endif
    if (rc) return rc;
    p.xtab = (const uint4 *)tab;
    if (hi > lo) quotient_kernel<<<(unsigned)((hi - lo + 127) / 128), 128, 0, ctx->stream>>>(p, (uint4 *)out_dev);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

// l_1_coset, the one epk table that needs an inversion per element: L1(x_i) = zh(x_i) / (n * (x_i - 1)) on the 4n coset.
// (keys/mod.rs:117-119 gets it as coset_fft(ifft(e_0)); same values.)  One multiplicative scan + one inversion.
int zkb_l1_coset_dev(zkb_ctx *ctx, unsigned log_n, uint64_t *out_dev) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!out_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_l1_coset_dev: null output");
    if (log_n + 2 > host::FR_TWO_ADICITY || log_n + 2 > 31) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_l1_coset_dev: 4n exceeds 2^28 (InvalidEvalDomainSize)");
    const size_t n = (size_t)1 << log_n;
    // L_1 = ifft(e_0) has all n coefficients equal to 1/n; its coset FFT over 4n is the table (keys/mod.rs:117-119)
    host::Fe ninv = host::inv(host::from_u64((uint64_t)n, host::FR), host::FR);
    fr_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint4 *)out_dev, n, dev_fe(ninv));
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return zkb_ntt_run(ctx, out_dev, n, log_n + 2, 0, 1);
}

int zkb_poly_eval_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, size_t n, const uint64_t z[4], uint64_t out[4]) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!z || !out || (!coeffs_dev && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_eval_dev: null argument");
    if (n == 0) { memset(out, 0, 32); return ZKB_OK; }
    unsigned lm = ceil_log2_sz(n);
    if (lm < 1) lm = 1;
    uint32_t nblocks = (uint32_t)((n + 256 * EV_K - 1) / (256 * EV_K));
    size_t tab_elems = ((size_t)1 << ((lm + 1) / 2)) + ((size_t)1 << (lm - (lm + 1) / 2));
    int rc = zkb_reserve(ctx, ctx->poly_ws, (tab_elems + nblocks + 2) * 32);
    if (rc) return rc;
    uint4 *tab = (uint4 *)ctx->poly_ws.p, *partial = tab + 2 * tab_elems, *res = partial + 2 * (size_t)nblocks;
    uint32_t s;
    host::Fe hz = host_fe(z);
    rc = zkb_pow2lvl_build(ctx, tab, lm, hz, host::one(host::FR), &s);
    if (rc) return rc;
    poly_eval_partial_kernel<<<nblocks, 256, 0, ctx->stream>>>((const uint4 *)coeffs_dev, n, dev_fe(hz), tab, s, partial);
    fr_sum_kernel<<<1, 256, 0, ctx->stream>>>(partial, nblocks, res);
    ctx->launches += 4;
    ZKB_CUDA(ctx, cudaGetLastError());
    ZKB_CUDA(ctx, cudaMemcpyAsync(out, res, 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

// k evaluations with one launch pair and ONE host round trip: round 5 evaluates 12 polynomials at xi or xi * w, and a round
// trip per evaluation (two table kernels, the partial sums, the total, a 32-byte download, a synchronise) cost 65 us each
// on a B200 (profiles/r02t_trace_2^18.jsonl) against 14 us of arithmetic.  Equal points share a table.
int zkb_poly_eval_many_dev(zkb_ctx *ctx, size_t k, const uint64_t *const *polys_dev, const size_t *lens, const uint64_t *points_host,
                           uint64_t *out_host) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (k == 0 || k > (size_t)LC_MAX) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_eval_many_dev: 1 <= k <= 16");
    if (!polys_dev || !lens || !points_host || !out_host) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_eval_many_dev: null argument");
    EvArgs a;
    size_t max_len = 0;
    uint32_t max_blocks = 0, total_blocks = 0;
    for (size_t i = 0; i < k; ++i) {
        if (!polys_dev[i] && lens[i]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_eval_many_dev: null polynomial");
        a.p[i] = (const uint4 *)polys_dev[i];
        a.len[i] = lens[i];
        memcpy(a.z[i].v, points_host + 4 * i, 32);
        const uint32_t nb = (uint32_t)((lens[i] + 256 * EV_K - 1) / (256 * EV_K));
        a.first[i] = total_blocks;
        total_blocks += nb;
        max_blocks = nb > max_blocks ? nb : max_blocks;
        max_len = lens[i] > max_len ? lens[i] : max_len;
    }
    a.first[k] = total_blocks;
    if (max_len == 0) { memset(out_host, 0, k * 32); return ZKB_OK; }
    unsigned lm = ceil_log2_sz(max_len);
    if (lm < 1) lm = 1;
    const size_t tab_elems = ((size_t)1 << ((lm + 1) / 2)) + ((size_t)1 << (lm - (lm + 1) / 2));
    size_t distinct[LC_MAX], n_distinct = 0;                            // index of the first polynomial with that point
    for (size_t i = 0; i < k; ++i) {
        size_t d = 0;
        while (d < n_distinct && memcmp(points_host + 4 * distinct[d], points_host + 4 * i, 32)) ++d;
        if (d == n_distinct) distinct[n_distinct++] = i;
        a.tab[i] = (uint32_t)(d * tab_elems);
    }
    int rc = zkb_reserve(ctx, ctx->poly_ws, (n_distinct * tab_elems + total_blocks + k + 2) * 32);
    if (rc) return rc;
    uint4 *tab = (uint4 *)ctx->poly_ws.p, *partial = tab + 2 * n_distinct * tab_elems, *res = partial + 2 * (size_t)total_blocks;
    for (size_t d = 0; d < n_distinct; ++d) {
        rc = zkb_pow2lvl_build(ctx, tab + 2 * d * tab_elems, lm, host_fe(points_host + 4 * distinct[d]), host::one(host::FR), &a.s);
        if (rc) return rc;
    }
    poly_eval_many_partial_kernel<<<dim3(max_blocks, (unsigned)k), 256, 0, ctx->stream>>>(a, tab, partial);
    fr_sum_many_kernel<<<(unsigned)k, 256, 0, ctx->stream>>>(a, partial, res);
    ctx->launches += 2 + 2 * n_distinct;
    ZKB_CUDA(ctx, cudaGetLastError());
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_host, res, k * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

int zkb_poly_lincomb_dev(zkb_ctx *ctx, size_t k, const uint64_t *const *polys_dev, const size_t *lens,
                         const uint64_t *scalars_host, uint64_t *out_dev, size_t out_len) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (k == 0 || k > (size_t)LC_MAX) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_lincomb_dev: 1 <= k <= 16");
    if (!polys_dev || !lens || !scalars_host || (!out_dev && out_len)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_lincomb_dev: null argument");
    LcArgs a;
    a.k = (uint32_t)k;
    for (size_t i = 0; i < k; ++i) {
        if (!polys_dev[i] && lens[i]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_lincomb_dev: null polynomial");
        a.p[i] = (const uint4 *)polys_dev[i];
        a.len[i] = lens[i];
        memcpy(a.s[i].v, scalars_host + 4 * i, 32);
    }
    if (out_len) poly_lincomb_kernel<<<(unsigned)((out_len + 255) / 256), 256, 0, ctx->stream>>>(a, (uint4 *)out_dev, out_len);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

int zkb_poly_divide_linear_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, size_t n, const uint64_t z[4], uint64_t *quot_dev,
                               uint64_t eval_out[4]) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!z || (!coeffs_dev && n) || (!quot_dev && n > 1)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_divide_linear_dev: null argument");
    if (n == 0) { if (eval_out) memset(eval_out, 0, 32); return ZKB_OK; }
    host::Fe hz = host_fe(z);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (host::is_zero(hz)) {
        if (n > 1) shift_down_kernel<<<blocks, 256, 0, ctx->stream>>>((const uint4 *)coeffs_dev, n, (uint4 *)quot_dev);
        ctx->launches += 1;
        ZKB_CUDA(ctx, cudaGetLastError());
        if (eval_out) ZKB_CUDA(ctx, cudaMemcpyAsync(eval_out, coeffs_dev, 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return ZKB_OK;
    }
    unsigned lm = ceil_log2_sz(n);
    if (lm < 1) lm = 1;
    size_t tab_elems = ((size_t)1 << ((lm + 1) / 2)) + ((size_t)1 << (lm - (lm + 1) / 2));
    size_t ntiles = (n + SC_TILE - 1) / SC_TILE + 2;
    int rc = zkb_reserve(ctx, ctx->poly_ws, (2 * tab_elems + n + ntiles) * 32);
    if (rc) return rc;
    uint4 *tab = (uint4 *)ctx->poly_ws.p, *itab = tab + 2 * tab_elems, *work = itab + 2 * tab_elems, *tiles = work + 2 * n;
    uint32_t s, si;
    rc = zkb_pow2lvl_build(ctx, tab, lm, hz, host::one(host::FR), &s);
    if (rc) return rc;
    rc = zkb_pow2lvl_build(ctx, itab, lm, host::inv(hz, host::FR), host::one(host::FR), &si);
    if (rc) return rc;
    scale_by_powers_kernel<<<blocks, 256, 0, ctx->stream>>>((const uint4 *)coeffs_dev, n, tab, s, work);
    rc = fr_scan<1, 1>(ctx, work, work, n, tiles);                    // inclusive suffix sums of p_k z^k
    if (rc) return rc;
    if (n > 1) divide_finish_kernel<<<blocks, 256, 0, ctx->stream>>>(work, n, itab, si, (uint4 *)quot_dev);
    ctx->launches += 6;
    ZKB_CUDA(ctx, cudaGetLastError());
    if (eval_out) ZKB_CUDA(ctx, cudaMemcpyAsync(eval_out, work, 32, cudaMemcpyDeviceToHost, ctx->stream));   // S_0 = p(z)
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

int zkb_poly_add_blinders_dev(zkb_ctx *ctx, uint64_t *coeffs_dev, size_t len, const uint64_t *blinders_host, size_t k) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!coeffs_dev || (!blinders_host && k)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_add_blinders_dev: null argument");
    if (k > (size_t)LC_MAX) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_add_blinders_dev: k <= 16");
    if (k == 0) return ZKB_OK;
    LcArgs b;
    b.k = (uint32_t)k;
    for (size_t i = 0; i < k; ++i) memcpy(b.s[i].v, blinders_host + 4 * i, 32);
    add_blinders_kernel<<<1, 32, 0, ctx->stream>>>((uint4 *)coeffs_dev, len, b);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

// DensePolynomial::from_coefficients_vec truncation: number of coefficients once trailing zeros are dropped.
int zkb_poly_effective_len_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, size_t n, size_t *out_len) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!out_len || (!coeffs_dev && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_poly_effective_len_dev: null argument");
    *out_len = 0;
    if (n == 0) return ZKB_OK;
    if (!ctx->len_slot) ZKB_CUDA(ctx, cudaMalloc((void **)&ctx->len_slot, 256));
    unsigned long long *slot = ctx->len_slot;
    ZKB_CUDA(ctx, cudaMemsetAsync(slot, 0, 8, ctx->stream));
    effective_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint4 *)coeffs_dev, n, slot);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    unsigned long long v = 0;
    ZKB_CUDA(ctx, cudaMemcpyAsync(&v, slot, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_len = (size_t)v;
    return ZKB_OK;
}

}  // extern "C"
