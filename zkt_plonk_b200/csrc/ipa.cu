// ipa.cu -- the folding rounds of the inner-product-argument commitment scheme on the device, for sm_100a.
//
// The reference's second `PC` (plonk-core/src/commitment.rs:49-86: `IPA<G, D>` = ark-poly-commit 0.3
// `ipa_pc::InnerProductArgPC`, run by every test of plonk-core/src/test.rs:73,84 on Bls12_381 / Bls12_377 G1) commits with the
// same MSM as KZG10 (over `ck.comm_key`) and opens with log2(d + 1) folding rounds.  One round of `open`, as the dependency
// runs it on the CPU over vectors of length n (coefficients c, powers of the point z, committer key G):
//     L = <c_r, G_l> + <c_r, z_l> h'      R = <c_l, G_r> + <c_l, z_r> h'          (two MSMs of n / 2, two inner products)
//     x = H(x_prev, L, R)                                                          (stays with the caller: a hash)
//     c_l += x^-1 c_r      z_l += x z_r      G_l += x G_r  (then batch-normalised to affine)
// zkb_ipa_round_lr_dev is the first line (the h' terms: two host scalar multiplications under the MSMs), and
// zkb_ipa_round_fold_dev the third.  All three vectors stay in HBM across the rounds; the caller sees two points and two field
// elements per round.  Results are group / field elements, so they equal the CPU's bit for bit in affine / Montgomery form.
//
// The key fold is the expensive half: n / 2 scalar multiplications by the SAME 255-bit challenge per round (n in total over an
// opening).  The challenge is recoded once on the host -- split through the curve's endomorphism into two 128-bit halves
// (x = k1 + lambda k2, glv_split), each in non-adjacent form -- and is uniform across the grid (no divergence); one thread per
// point runs 128 doublings with the halves' additions interleaved, in XYZZ coordinates, and normalises its own result
// (ipa_fold_key_glv_kernel; ipa_fold_key_kernel is the plain fold, kept as the fallback and for A/B).
// zkb_ipa_final_key_dev is the verifier's linear-time step: the check polynomial's coefficients expanded in HBM + one MSM.
#include "ctx.h"
#include "ec.cuh"

#include <stdlib.h>

using namespace zkb;
typedef FrP F;

namespace {

constexpr int AFF_W = 2 * host::FQ_L;

struct Bits256 { uint32_t w[8]; };

// partial[2 * cta] = sum c_r[i] z_l[i], partial[2 * cta + 1] = sum c_l[i] z_r[i] over the CTA's i < half
__global__ void __launch_bounds__(256) ipa_dots_partial_kernel(const uint4 *c, const uint4 *z, size_t half, uint4 *partial) {
    __shared__ fe_t sm[2][8];
    fe_t a0 = fzero<F>(), a1 = fzero<F>();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < half; i += (size_t)gridDim.x * 256) {
        const fe_t cl = fload_ro(c + 2 * i), cr = fload_ro(c + 2 * (half + i));
        const fe_t zl = fload_ro(z + 2 * i), zr = fload_ro(z + 2 * (half + i));
        a0 = fadd<F>(a0, fmul<F>(cr, zl));
        a1 = fadd<F>(a1, fmul<F>(cl, zr));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        fe_t o0, o1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o0.v[k] = __shfl_down_sync(0xffffffffu, a0.v[k], d);
            o1.v[k] = __shfl_down_sync(0xffffffffu, a1.v[k], d);
        }
        a0 = fadd<F>(a0, o0);
        a1 = fadd<F>(a1, o1);
    }
    if (lane == 0) { sm[0][wid] = a0; sm[1][wid] = a1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        fe_t s = sm[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) s = fadd<F>(s, sm[threadIdx.x][w]);
        fstore(partial + 2 * (2 * (size_t)blockIdx.x + threadIdx.x), s);
    }
}

// out[k] = sum over ctas of partial[2 * cta + k], k = 0, 1 (one CTA of 64 threads: a warp per sum)
__global__ void __launch_bounds__(64) ipa_dots_total_kernel(const uint4 *partial, uint32_t ctas, uint4 *out) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fe_t acc = fzero<F>();
    for (uint32_t i = lane; i < ctas; i += 32) acc = fadd<F>(acc, fload(partial + 2 * (2 * (size_t)i + k)));
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        fe_t o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = __shfl_down_sync(0xffffffffu, acc.v[j], d);
        acc = fadd<F>(acc, o);
    }
    if (lane == 0) fstore(out + 2 * k, acc);
}

// c_l += x^-1 c_r, z_l += x z_r
__global__ void __launch_bounds__(256) ipa_fold_scalars_kernel(uint4 *c, uint4 *z, size_t half, fe_t x, fe_t x_inv) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= half) return;
    fstore(c + 2 * i, fadd<F>(fload(c + 2 * i), fmul<F>(x_inv, fload_ro(c + 2 * (half + i)))));
    fstore(z + 2 * i, fadd<F>(fload(z + 2 * i), fmul<F>(x, fload_ro(z + 2 * (half + i)))));
}

// G_l[i] = G_l[i] + x * G_r[i], affine.  x arrives in non-adjacent form: digit b is +1 where `pos` has bit b, -1 where `neg` has
// it (never both, never two neighbours), `top` = index of the highest digit (always +1).  A third of the digits are non-zero
// against half of the bits, and -P is free on a short Weierstrass curve: ~85 mixed additions per point instead of ~127.
__global__ void __launch_bounds__(128) ipa_fold_key_kernel(g1a_t *key, uint32_t half, const __grid_constant__ Bits256 pos,
                                                           const __grid_constant__ Bits256 neg, int top) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const g1a_t r = g1a_load(key + half + i);
    const g1a_t rn = g1a_neg(r);
    g1x_t acc = g1x_from_affine(r);                               // the top digit
    for (int b = top - 1; b >= 0; --b) {
        acc = g1x_double(acc);
        const uint32_t m = 1u << (b & 31);
        if (pos.w[b >> 5] & m) g1x_add_mixed(acc, r);
        else if (neg.w[b >> 5] & m) g1x_add_mixed(acc, rn);
    }
    g1a_t l;                                                      // this thread overwrites the entry: not through the read-only path
    l.x = floadn<FqP::N>(&key[i].x);
    l.y = floadn<FqP::N>(&key[i].y);
    g1x_add_mixed(acc, l);
    const g1a_t a = g1x_to_affine(acc);
    fstore(&key[i].x, a.x);
    fstore(&key[i].y, a.y);
}

// Non-adjacent form of a canonical scalar below 2^255: with t = 3 x, digit b is +1 where bit b + 1 of t is set and that of x is
// not, -1 the other way round (Reitwiesner).  Returns the index of the top digit (-1 for x = 0).
int naf_of(const uint64_t x[4], Bits256 *pos, Bits256 *neg) {
    uint64_t xe[5] = {x[0], x[1], x[2], x[3], 0}, t[5], p[5], q[5];
    unsigned __int128 carry = 0;
    for (int k = 0; k < 5; ++k) {
        carry += (unsigned __int128)xe[k] * 3;
        t[k] = (uint64_t)carry;
        carry >>= 64;
    }
    for (int k = 0; k < 5; ++k) { p[k] = t[k] & ~xe[k]; q[k] = ~t[k] & xe[k]; }
    for (int k = 0; k < 4; ++k) {                                 // >> 1 across the five words; the result fits four
        const uint64_t pk = (p[k] >> 1) | (p[k + 1] << 63), qk = (q[k] >> 1) | (q[k + 1] << 63);
        pos->w[2 * k] = (uint32_t)pk; pos->w[2 * k + 1] = (uint32_t)(pk >> 32);
        neg->w[2 * k] = (uint32_t)qk; neg->w[2 * k + 1] = (uint32_t)(qk >> 32);
    }
    int top = 255;
    while (top >= 0 && !((pos->w[top >> 5] >> (top & 31)) & 1)) --top;
    return top;                                                   // -1 for x = 0
}

// The same fold through the curve's endomorphism: x = k1 + lambda k2 (mod r) with |k1|, |k2| < 2^130 and lambda (x, y) = (beta x, y), so
// x P = k1 P + k2 (beta P.x, P.y): half the doublings (and half the dependent chain, which is what a small round waits for).  Both
// halves arrive in non-adjacent form (GlvDigits); flip1 / flip2: the half is negative (its point is negated instead).
struct GlvDigits { uint32_t w[4][8]; };       // non-adjacent forms: [0] / [1] = +1 / -1 digits of k1, [2] / [3] of k2

static __device__ __noinline__ void g1x_add_xy(g1x_t &acc, const fq_t &x, const fq_t &y) {
    g1a_t q;
    q.x = x;
    q.y = y;
    g1x_add_mixed(acc, q);
}

// ONE call site each for the doubling and the addition (the loop body stays small: the first version, with four inlined additions,
// ran at an instruction-cache hit rate of 73 %, profiles/r02bh); MINB: resident CTAs per SM the register allocation is bounded for.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) ipa_fold_key_glv_kernel(g1a_t *key, uint32_t half, const __grid_constant__ GlvDigits dig, int top,
                                                                     int flip1, int flip2, const __grid_constant__ fq_t beta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const g1a_t r = g1a_load(key + half + i);
    g1x_t acc = g1x_inf();
    if (!g1a_is_inf(r)) {
        const fq_t bx = fmul<Q>(r.x, beta);
        const fq_t ny = fneg<Q>(r.y);
        for (int bit = top; bit >= 0; --bit) {
            acc = g1x_double(acc);
            const uint32_t m = 1u << (bit & 31), w = bit >> 5;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const bool plus = dig.w[2 * h][w] & m, minus = dig.w[2 * h + 1][w] & m;
                if (plus | minus) {
                    const bool negate = minus != (bool)(h ? flip2 : flip1);       // -digit, or a negative half: one of them flips y
                    g1x_add_xy(acc, h ? bx : r.x, negate ? ny : r.y);
                }
            }
        }
    }
    g1a_t l;
    l.x = floadn<FqP::N>(&key[i].x);
    l.y = floadn<FqP::N>(&key[i].y);
    g1x_add_mixed(acc, l);
    const g1a_t o = g1x_to_affine(acc);
    fstore(&key[i].x, o.x);
    fstore(&key[i].y, o.y);
}

// (k G) >> 384 for a 4-word k and a 5-word G: the rounding of the GLV split (three words at most; callers check the size)
void mulshift384(const uint64_t k[4], const uint64_t g[5], uint64_t out[4]) {
    uint64_t prod[9] = {0};
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 carry = 0;
        for (int j = 0; j < 5; ++j) {
            carry += (unsigned __int128)k[i] * g[j] + prod[i + j];
            prod[i + j] = (uint64_t)carry;
            carry >>= 64;
        }
        prod[i + 5] = (uint64_t)carry;
    }
    out[0] = prod[6]; out[1] = prod[7]; out[2] = prod[8]; out[3] = 0;
}

// |v| and sign of a residue read as the signed integer of smallest magnitude; false when that is not below 2^136
bool short_signed(const host::Fe &v, uint64_t mag[4], int *neg) {
    host::Fe zero, m;
    memset(zero.l, 0, 32);
    m = host::sub(zero, v, host::FR);                              // r - v (0 for v = 0)
    *neg = !host::ge<host::FR_L>(m.l, v.l);                        // r - v < v: the negative reading is the short one
    memcpy(mag, *neg ? m.l : v.l, 32);
    return mag[3] == 0 && mag[2] < 256;
}

// k (canonical, below r) = k1 + lambda k2 (mod r), the halves as magnitude and sign.  The result is verified; false: not short.
bool glv_split(const uint64_t k[4], uint64_t m1[4], int *neg1, uint64_t m2[4], int *neg2) {
    host::Fe c1, c2, kk, lam, mm1, mm2;
    mulshift384(k, host::GLV_G1, c1.l);
    mulshift384(k, host::GLV_G2, c2.l);
    memcpy(kk.l, k, 32);
    memcpy(lam.l, host::GLV_LAMBDA_MONT, 32);
    memcpy(mm1.l, host::GLV_M1_MONT, 32);
    memcpy(mm2.l, host::GLV_M2_MONT, 32);
    if (host::ge<host::FR_L>(c1.l, host::FR.p) || host::ge<host::FR_L>(c2.l, host::FR.p)) return false;
    // a canonical integer times a Montgomery form is the canonical product
    const host::Fe k2 = host::add(host::mul(c1, mm1, host::FR), host::mul(c2, mm2, host::FR), host::FR);
    const host::Fe k1 = host::sub(kk, host::mul(k2, lam, host::FR), host::FR);
    if (!short_signed(k1, m1, neg1) || !short_signed(k2, m2, neg2)) return false;
    // verify: (+-m1) + lambda (+-m2) == k
    host::Fe zero, a, b;
    memset(zero.l, 0, 32);
    memcpy(a.l, m1, 32);
    memcpy(b.l, m2, 32);
    if (*neg1) a = host::sub(zero, a, host::FR);
    if (*neg2) b = host::sub(zero, b, host::FR);
    const host::Fe back = host::add(a, host::mul(b, lam, host::FR), host::FR);
    return memcmp(back.l, k, 32) == 0;
}

// Coefficients of the verifier's check polynomial h(X) = prod_j (1 + x_j X^(2^(k-1-j))) (ipa_pc SuccinctCheckPolynomial::compute_coeffs):
// coefficient i is the product of the challenges whose bit (k - 1 - j) is set in i.  One thread per coefficient, <= k products.
struct CheckChallenges { fe_t x[30]; };
__global__ void __launch_bounds__(256) ipa_check_coeffs_kernel(uint4 *out, uint32_t k, const __grid_constant__ CheckChallenges ch) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >> k) return;
    fe_t acc = fone<F>();
    for (uint32_t j = 0; j < k; ++j)
        if ((i >> (k - 1 - j)) & 1) acc = fmul<F>(acc, ch.x[j]);
    fstore(out + 2 * i, acc);
}

bool pow2(size_t n) { return n >= 2 && (n & (n - 1)) == 0; }

}  // namespace

// Test hook (host only): the GLV split of a canonical scalar as the key fold uses it; returns 1 when it is short and verified.
int zkb_test_glv_split(const uint64_t k[4], uint64_t m1[4], int *neg1, uint64_t m2[4], int *neg2) {
    if (!k || !m1 || !neg1 || !m2 || !neg2 || host::ge<host::FR_L>(k, host::FR.p)) return ZKB_ERR_INVALID;
    return glv_split(k, m1, neg1, m2, neg2) ? 1 : 0;
}

// One round's cross terms over c, z (n Montgomery field elements each) and G (n affine points), all in HBM, n a power of two:
// ip_l = <c_r, z_l>, ip_r = <c_l, z_r> (Montgomery), l_xy = <c_r, G_l> + ip_l h', r_xy = <c_l, G_r> + ip_r h' (affine, Montgomery;
// identity = zeros and *_inf = 1).  h_prime_xy: the round-independent point h' = x_0 h in host memory, or NULL to get the two
// MSMs alone.  The h' terms are two host scalar multiplications run while the GPU is inside the MSMs.  The caller hashes.
int zkb_ipa_round_lr_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, const uint64_t *z_dev, const uint64_t *key_dev, size_t n,
                         const uint64_t *h_prime_xy, uint64_t *l_xy, int *l_inf, uint64_t *r_xy, int *r_inf, uint64_t ip_l[4],
                         uint64_t ip_r[4]) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!coeffs_dev || !z_dev || !key_dev || !l_xy || !r_xy || !ip_l || !ip_r) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ipa_round_lr_dev: null argument");
    if (!pow2(n) || n > ((size_t)1 << 30)) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_ipa_round_lr_dev: n must be a power of two, 2 <= n <= 2^30");
    const size_t half = n / 2;
    uint32_t ctas = (uint32_t)((half + 255) / 256);
    const uint32_t cap = (uint32_t)ctx->sm_count * 8;
    if (ctas > cap) ctas = cap;
    int rc = zkb_reserve(ctx, ctx->poly_ws, ((size_t)2 * ctas + 2) * 32);
    if (rc) return rc;
    uint4 *partial = (uint4 *)ctx->poly_ws.p, *total = partial + 2 * (size_t)2 * ctas;
    ipa_dots_partial_kernel<<<ctas, 256, 0, ctx->stream>>>((const uint4 *)coeffs_dev, (const uint4 *)z_dev, half, partial);
    ipa_dots_total_kernel<<<1, 64, 0, ctx->stream>>>(partial, ctas, total);
    ctx->launches += 2;
    ZKB_CUDA(ctx, cudaGetLastError());
    uint64_t ips[8];
    ZKB_CUDA(ctx, cudaMemcpyAsync(ips, total, 64, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    host::Fe one, canon[2];
    memset(one.l, 0, 32);
    one.l[0] = 1;
    for (int k = 0; k < 2; ++k) {                                  // canonical integers: the scalars of the h' terms
        host::Fe m;
        memcpy(m.l, ips + 4 * k, 32);
        canon[k] = host::mul(m, one, host::FR);
    }
    rc = zkb_msm_points_plus(ctx, key_dev, coeffs_dev + 4 * half, half, 1, h_prime_xy, canon[0].l, l_xy, l_inf);
    if (rc) return rc;
    rc = zkb_msm_points_plus(ctx, key_dev + (size_t)AFF_W * half, coeffs_dev, half, 1, h_prime_xy, canon[1].l, r_xy, r_inf);
    if (rc) return rc;
    memcpy(ip_l, ips, 32);
    memcpy(ip_r, ips + 4, 32);
    return ZKB_OK;
}

// The linear-time half of the verifier (ipa_pc check after succinct_check): <h-coefficients, G> over the whole committer key, to be
// compared with the proof's final_comm_key.  challenges_mont: the log2 n round challenges in proof order, Montgomery form.
int zkb_ipa_final_key_dev(zkb_ctx *ctx, const uint64_t *key_dev, size_t n, const uint64_t *challenges_mont, uint64_t *out_xy, int *is_inf) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!key_dev || !challenges_mont || !out_xy) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ipa_final_key_dev: null argument");
    if (!pow2(n) || n > ((size_t)1 << 30)) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_ipa_final_key_dev: n must be a power of two, 2 <= n <= 2^30");
    uint32_t k = 0;
    while (((size_t)1 << k) < n) ++k;
    CheckChallenges ch;
    memset(&ch, 0, sizeof ch);
    for (uint32_t j = 0; j < k; ++j) {
        if (host::ge<host::FR_L>(challenges_mont + 4 * j, host::FR.p)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ipa_final_key_dev: challenge not reduced");
        memcpy(ch.x[j].v, challenges_mont + 4 * j, 32);
    }
    int rc = zkb_reserve(ctx, ctx->poly_ws, n * 32);
    if (rc) return rc;
    ipa_check_coeffs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint4 *)ctx->poly_ws.p, k, ch);
    ctx->launches += 1;
    ZKB_CUDA(ctx, cudaGetLastError());
    return zkb_msm_points_plus(ctx, key_dev, (const uint64_t *)ctx->poly_ws.p, n, 1, nullptr, nullptr, out_xy, is_inf);
}

// The round's fold, in place: afterwards the first n / 2 entries of each vector are the next round's vectors.
// x, x_inv: the round challenge and its inverse, Montgomery form (x * x_inv == 1 is checked).
int zkb_ipa_round_fold_dev(zkb_ctx *ctx, uint64_t *coeffs_dev, uint64_t *z_dev, uint64_t *key_dev, size_t n, const uint64_t x[4],
                           const uint64_t x_inv[4]) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!coeffs_dev || !z_dev || !key_dev || !x || !x_inv) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ipa_round_fold_dev: null argument");
    if (!pow2(n) || n > ((size_t)1 << 30)) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_ipa_round_fold_dev: n must be a power of two, 2 <= n <= 2^30");
    host::Fe hx, hxi, one;
    memcpy(hx.l, x, 32);
    memcpy(hxi.l, x_inv, 32);
    memset(one.l, 0, 32);
    one.l[0] = 1;
    if (host::ge<host::FR_L>(hx.l, host::FR.p) || host::ge<host::FR_L>(hxi.l, host::FR.p)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ipa_round_fold_dev: challenge not reduced");
    const host::Fe prod = host::mul(host::mul(hx, hxi, host::FR), one, host::FR);      // canonical x * x_inv
    if (memcmp(prod.l, one.l, 32) != 0) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ipa_round_fold_dev: x * x_inv != 1");
    const host::Fe canon = host::mul(hx, one, host::FR);
    // key fold: through the endomorphism (half the doublings) unless ZKB_IPA_GLV=0 or the split is not short (never, for the
    // generated constants); ZKB_IPA_NAF=0 walks plain binary expansions instead of non-adjacent forms (A/B knobs, same results)
    const char *e_glv = getenv("ZKB_IPA_GLV"), *e_naf = getenv("ZKB_IPA_NAF");
    const bool use_naf = !(e_naf && e_naf[0] == '0');
    auto digits = [&](const uint64_t v[4], Bits256 *pos, Bits256 *neg) {
        int top = naf_of(v, pos, neg);
        if (!use_naf) {
            memcpy(pos->w, v, 32);
            memset(neg->w, 0, 32);
            top = 255;
            while (top >= 0 && !((pos->w[top >> 5] >> (top & 31)) & 1)) --top;
        }
        return top;
    };
    uint64_t m1[4], m2[4];
    int neg1 = 0, neg2 = 0;
    const bool use_glv = !(e_glv && e_glv[0] == '0') && glv_split(canon.l, m1, &neg1, m2, &neg2);
    Bits256 pos, neg, pos2, neg2b;
    int top = 0, top2 = -1;
    if (use_glv) {
        top = digits(m1, &pos, &neg);
        top2 = digits(m2, &pos2, &neg2b);
    } else {
        top = digits(canon.l, &pos, &neg);
    }
    fe_t dx, dxi;
    memcpy(dx.v, x, 32);
    memcpy(dxi.v, x_inv, 32);
    const size_t half = n / 2;
    ipa_fold_scalars_kernel<<<(unsigned)((half + 255) / 256), 256, 0, ctx->stream>>>((uint4 *)coeffs_dev, (uint4 *)z_dev, half, dx, dxi);
    if (use_glv) {
        fq_t beta;
        memcpy(beta.v, host::GLV_BETA, sizeof beta.v);
        GlvDigits dig;
        memcpy(dig.w[0], pos.w, 32);
        memcpy(dig.w[1], neg.w, 32);
        memcpy(dig.w[2], pos2.w, 32);
        memcpy(dig.w[3], neg2b.w, 32);
        const int t = top > top2 ? top : top2;
        ipa_fold_key_glv_kernel<FqP::N == 8 ? 4 : 2><<<(unsigned)((half + 127) / 128), 128, 0, ctx->stream>>>((g1a_t *)key_dev, (uint32_t)half, dig, t,
                                                                                                          neg1, neg2, beta);
    } else {
        ipa_fold_key_kernel<<<(unsigned)((half + 127) / 128), 128, 0, ctx->stream>>>((g1a_t *)key_dev, (uint32_t)half, pos, neg, top);
    }
    ctx->launches += 2;
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}
