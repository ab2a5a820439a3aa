// ntt.cu -- radix-2 NTT / INTT / coset variants over BN254 Fr for sm_100a.
//
// Replaces ark-poly 0.3 Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place as called through
// plonk-core/src/util.rs:63-140.  Natural order in and out, Montgomery form, 32-byte elements.
//
// Decomposition (multi-dimensional Cooley-Tukey, "four-step" generalised to <= 3 passes).  With
// N = T1*S1 (S1 = T2*S2 ...), a pass runs, for every inner offset c < S and outer block r, a size-T
// transform on the elements  base + j*S  (base = r*T*S + c), entirely in shared memory, then multiplies
// output k by the inter-pass twiddle w_{T*S}^{c*k} and writes it back to  base + k*S.  The last pass
// (S = 1) scatters output k of block r to  digit_reverse(r) + k*(N/T), which restores natural order.
// Because an Fr element is exactly one 32-byte sector, strided element access wastes no HBM bytes; the
// passes move 64*N bytes each.
//
// In a tile the transform is decimation-in-frequency with one __syncthreads per stage; elements sit in
// shared memory as two 16-byte planes (low / high half) so that a quarter-warp always covers all 32 banks.
// Coset scaling and the n^-1 factor are fused into the first load / last store.
#include <stdlib.h>

#include <algorithm>

#include "ctx.h"
#include "ff.cuh"
#include "host_ff.h"

using namespace zkb;

namespace {

enum : uint32_t { M_IN_COSET = 1, M_OUT_COSET = 2, M_OUT_CONST = 4, M_LAST = 8, M_FIRST = 16 };

constexpr unsigned NTT_MAX_BATCH = 16;   // transforms of one launch (blockIdx.y)

struct PassArgs {
    uint4 *data[NTT_MAX_BATCH];   // the transforms' own buffers: input of the first pass, output of the last
    uint4 *scr;                   // ping-pong scratch of the middle passes: transform y at scr + y * scr_stride
    unsigned long long scr_stride;
    const uint4 *tile_tw;   // w_T^j, j < T/2
    const uint4 *tw2;       // two-level table of w_M (M = T*S): lo[2^tw_s] then hi[M >> tw_s]
    const uint4 *cs2;       // two-level table of coset powers (g^e or g^-e / n): lo[2^cs_s] then hi
    const uint4 *tw_direct; // optional: w_M^(c*k) at [c * T + k]  (one product instead of two per element)
    const uint4 *cs_direct; // optional: coset power of every index e at [e]
    const uint4 *cs_tile;   // optional (first pass of a forward coset transform): (g^S)^j, j < T -- see the fold below
    unsigned long long len; // elements >= len read as zero (first pass only)
    uint32_t t, log_s, log_n;
    uint32_t tw_s, cs_s;
    uint32_t mode;
    uint32_t t1, t2;        // bits of pass 1 and of the middle pass (digit reversal in the last pass)
    fe_t scale;             // n^-1 for a plain inverse transform
};

// One spare 16-byte unit after every 8 elements: the register tail reads 8 consecutive elements per thread
// (stride 8 between threads), which would otherwise put a whole quarter-warp on one bank group.
__device__ __forceinline__ uint32_t sm_idx(uint32_t i) { return i + (i >> 3); }

__device__ __forceinline__ fe_t sm_load(const uint4 *lo, const uint4 *hi, uint32_t i) {
    i = sm_idx(i);
    uint4 a = lo[i], b = hi[i];
    fe_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sm_store(uint4 *lo, uint4 *hi, uint32_t i, const fe_t &r) {
    i = sm_idx(i);
    lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// base^e from a two-level table: lo[e & mask] * hi[e >> s]
__device__ __forceinline__ fe_t pow2lvl(const uint4 *tab, uint32_t s, unsigned long long e) {
    uint32_t lo = (uint32_t)(e & ((1ull << s) - 1));
    unsigned long long hi = e >> s;
    fe_t a = fload_ro(tab + 2 * (size_t)lo);
    fe_t b = fload_ro(tab + 2 * ((size_t)(1u << s) + hi));
    return fmul<FrP>(a, b);
}

// Factor of input element idx = c + j * S of the first pass of a forward coset transform.  g^idx = g^c * (g^S)^j, and the
// tile's transform is linear, so g^c can leave through the OUTPUT side: with cs_tile set the load multiplies by (g^S)^j (T
// entries: they stay in L1 / L2) and the inter-pass table of this pass carries w_M^(c k) * g^c.  Same number of products,
// one expanded table (N x 32 B) less to stream per transform.
__device__ __forceinline__ fe_t coset_in_factor(const PassArgs &a, unsigned long long idx, uint32_t j) {
    if (a.cs_tile) return fload_ro(a.cs_tile + 2 * (size_t)j);
    return a.cs_direct ? fload_ro(a.cs_direct + 2 * idx) : pow2lvl(a.cs2, a.cs_s, idx);
}

// inter-pass twiddle / output scaling and the store of output k of this tile (shared by both store paths)
__device__ __forceinline__ void emit_output(const PassArgs &a, uint4 *out, unsigned long long c, unsigned long long r,
                                            unsigned long long base, uint32_t k, fe_t v) {
    unsigned long long o;
    if (!(a.mode & M_LAST)) {
        v = fmul<FrP>(v, a.tw_direct ? fload_ro(a.tw_direct + 2 * ((c << a.t) + k)) : pow2lvl(a.tw2, a.tw_s, c * k));
        o = base + ((unsigned long long)k << a.log_s);
    } else {
        // r = k1 * 2^t2 + k2  ->  K = k1 + 2^t1 * k2 + 2^(t1+t2) * k
        unsigned long long k1 = r >> a.t2, k2 = r & ((1ull << a.t2) - 1);
        o = k1 + (k2 << a.t1) + ((unsigned long long)k << (a.log_n - a.t));
        if (a.mode & M_OUT_COSET) v = fmul<FrP>(v, a.cs_direct ? fload_ro(a.cs_direct + 2 * o) : pow2lvl(a.cs2, a.cs_s, o));
        else if (a.mode & M_OUT_CONST) v = fmul<FrP>(v, a.scale);
    }
    fstore(out + 2 * o, v);
}

__device__ __forceinline__ void bfly(fe_t &u, fe_t &v) {        // (u, v) <- (u + v, u - v)
    fe_t s = fadd<FrP>(u, v);
    v = fsub<FrP>(u, v);
    u = s;
}

__global__ void __launch_bounds__(256) ntt_pass_kernel(const __grid_constant__ PassArgs a) {
    extern __shared__ uint4 sm[];
    uint4 *const mine = a.data[blockIdx.y], *const scr = a.scr + blockIdx.y * a.scr_stride;
    const uint4 *const in = (a.mode & M_FIRST) ? mine : scr;
    uint4 *const out = (a.mode & M_LAST) ? mine : scr;
    const uint32_t T = 1u << a.t;
    uint4 *s_lo = sm, *s_hi = sm + T + (T >> 3) + 1;
    const uint32_t tid = threadIdx.x, NT = blockDim.x;

    const unsigned long long tile = blockIdx.x;
    const unsigned long long c = tile & ((1ull << a.log_s) - 1);
    const unsigned long long r = tile >> a.log_s;
    const unsigned long long base = (r << (a.t + a.log_s)) + c;

    // ---- load (zero padding + coset scaling fused)
    for (uint32_t j = tid; j < T; j += NT) {
        unsigned long long idx = base + ((unsigned long long)j << a.log_s);
        fe_t v;
        if (idx < a.len) {
            v = fload(in + 2 * idx);
            if (a.mode & M_IN_COSET) v = fmul<FrP>(v, coset_in_factor(a, idx, j));
        } else {
            v = fzero<FrP>();
        }
        sm_store(s_lo, s_hi, j, v);
    }

    // ---- decimation in frequency through shared memory, down to (not including) the last three stages
    const uint32_t tail = a.t >= 3 ? 3 : 0;           // stages done in registers
    for (uint32_t lh = a.t; lh-- > tail;) {           // h = 2^lh : butterfly half distance
        __syncthreads();
        const uint32_t h = 1u << lh;
        const uint32_t tw_shift = a.t - 1 - lh;       // twiddle index = j * (T / 2h)
        for (uint32_t b = tid; b < (T >> 1); b += NT) {
            uint32_t j = b & (h - 1);
            uint32_t i = ((b - j) << 1) + j;
            fe_t u = sm_load(s_lo, s_hi, i);
            fe_t v = sm_load(s_lo, s_hi, i + h);
            fe_t s = fadd<FrP>(u, v);
            fe_t d = fsub<FrP>(u, v);
            if (lh) d = fmul<FrP>(d, fload_ro(a.tile_tw + 2 * (size_t)(j << tw_shift)));
            sm_store(s_lo, s_hi, i, s);
            sm_store(s_lo, s_hi, i + h, d);
        }
    }
    __syncthreads();

    if (!tail) {                                       // tiny tiles: store straight from shared memory
        for (uint32_t i = tid; i < T; i += NT) {
            uint32_t k = a.t ? (__brev(i) >> (32 - a.t)) : 0;
            emit_output(a, out, c, r, base, k, sm_load(s_lo, s_hi, i));
        }
        return;
    }

    // ---- last three stages (h = 4, 2, 1) on 8 consecutive elements held in registers: 5 twiddle products
    // instead of 12 (w_8^0 and w_4^0 are 1, the last stage has no twiddle), no barriers, no bank conflicts,
    // and the results go to global memory without another trip through shared memory
    const fe_t w8_1 = fload_ro(a.tile_tw + 2 * (size_t)(T >> 3));
    const fe_t w8_2 = fload_ro(a.tile_tw + 2 * (size_t)(T >> 2));      // = w_4
    const fe_t w8_3 = fload_ro(a.tile_tw + 2 * (size_t)(3 * (T >> 3)));
    for (uint32_t q = tid; q < (T >> 3); q += NT) {
        fe_t x[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) x[p] = sm_load(s_lo, s_hi, 8 * q + p);
        bfly(x[0], x[4]); bfly(x[1], x[5]); bfly(x[2], x[6]); bfly(x[3], x[7]);
        x[5] = fmul<FrP>(x[5], w8_1); x[6] = fmul<FrP>(x[6], w8_2); x[7] = fmul<FrP>(x[7], w8_3);
        bfly(x[0], x[2]); bfly(x[1], x[3]); bfly(x[4], x[6]); bfly(x[5], x[7]);
        x[3] = fmul<FrP>(x[3], w8_2); x[7] = fmul<FrP>(x[7], w8_2);
        bfly(x[0], x[1]); bfly(x[2], x[3]); bfly(x[4], x[5]); bfly(x[6], x[7]);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            uint32_t k = __brev(8 * q + p) >> (32 - a.t);
            emit_output(a, out, c, r, base, k, x[p]);
        }
    }
}


// ------------------------------------------------------------------ compile-time tile size: radix-4 steps, 3 CTAs per SM
// Same pass as ntt_pass_kernel (same PassArgs, same tables, bit-identical results) restructured around what limited it
// (ncu, round 1: 2 CTAs of 256 threads per SM at 108 registers, 8 barriers per tile, twiddle loads stalling each product):
//   * LT is a template parameter, every loop is unrolled, so ptxas hoists the twiddle and shared-memory loads of a step above
//     its products;
//   * two butterfly stages per barrier (a radix-4 step on elements j, j + H/2, j + H, j + 3H/2: four products, three
//     twiddles w^j, w^(j + H/2), w^(2j)), the first stage of an odd-sized tile is done on the fly while loading (its
//     partners j and j + T/2 are loaded by the same thread), the last two stages run in registers on 4 consecutive
//     elements and go straight to global memory: 5 barriers instead of 8 for T = 2048;
//   * at most 4 elements live per thread (32 data registers), so the kernel fits 80 registers: 3 CTAs of 256 threads
//     (6 warps per scheduler instead of 4) to cover the fixed-latency carry chains and the load / store phases.
template <int LT>
__global__ void __launch_bounds__((1 << LT) / 8, 768 / ((1 << LT) / 8)) ntt_pass_kernel_r4(const __grid_constant__ PassArgs a) {
    constexpr uint32_t T = 1u << LT, NT = T / 8;
    constexpr bool ODD = (LT & 1) != 0;                            // odd tile bits: stage LT-1 is fused into the load
    extern __shared__ uint4 sm[];
    uint4 *const mine = a.data[blockIdx.y], *const scr = a.scr + blockIdx.y * a.scr_stride;
    const uint4 *const in = (a.mode & M_FIRST) ? mine : scr;
    uint4 *const out = (a.mode & M_LAST) ? mine : scr;
    uint4 *s_lo = sm, *s_hi = sm + T + (T >> 3) + 1;
    const uint32_t tid = threadIdx.x;
    const unsigned long long tile = blockIdx.x;
    const unsigned long long c = tile & ((1ull << a.log_s) - 1);
    const unsigned long long r = tile >> a.log_s;
    const unsigned long long base = (r << (LT + a.log_s)) + c;

    auto load_elem = [&](uint32_t j) -> fe_t {
        const unsigned long long idx = base + ((unsigned long long)j << a.log_s);
        if (idx >= a.len) return fzero<FrP>();
        fe_t v = fload(in + 2 * idx);
        if (a.mode & M_IN_COSET) v = fmul<FrP>(v, coset_in_factor(a, idx, j));
        return v;
    };

    // ---- load (zero padding + coset scaling fused); odd LT: butterfly with the partner T/2 away before the store
    if (ODD) {
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t j = tid + NT * q;                       // j < T/2
            fe_t u = load_elem(j), v = load_elem(j + T / 2);
            fe_t d = fsub<FrP>(u, v);
            u = fadd<FrP>(u, v);
            d = fmul<FrP>(d, fload_ro(a.tile_tw + 2 * (size_t)j));
            sm_store(s_lo, s_hi, j, u);
            sm_store(s_lo, s_hi, j + T / 2, d);
        }
    } else {
#pragma unroll
        for (uint32_t q = 0; q < 8; ++q) sm_store(s_lo, s_hi, tid + NT * q, load_elem(tid + NT * q));
    }

    // ---- radix-4 steps through shared memory: stages (lh, lh - 1), lh = LT-1-ODD, LT-3-ODD, ..., 3
    constexpr int LH0 = LT - 1 - (ODD ? 1 : 0);
#pragma unroll
    for (int lh = LH0; lh >= 3; lh -= 2) {
        __syncthreads();
        const uint32_t H = 1u << lh, s1 = 1u << (LT - 1 - lh);     // twiddle stride of stage lh; stage lh - 1 uses 2 * s1
#pragma unroll
        for (uint32_t q = 0; q < 2; ++q) {
            const uint32_t g = tid + NT * q;                       // group of 4, g < T/4
            const uint32_t j = g & (H / 2 - 1), blk = g >> (lh - 1);
            const uint32_t e0 = (blk << (lh + 1)) + j;
            fe_t x0 = sm_load(s_lo, s_hi, e0), x1 = sm_load(s_lo, s_hi, e0 + H / 2);
            fe_t x2 = sm_load(s_lo, s_hi, e0 + H), x3 = sm_load(s_lo, s_hi, e0 + H + H / 2);
            // stage lh: (x0, x2) with w^(j s1), (x1, x3) with w^((j + H/2) s1)
            fe_t d = fsub<FrP>(x0, x2);
            x0 = fadd<FrP>(x0, x2);
            x2 = fmul<FrP>(d, fload_ro(a.tile_tw + 2 * (size_t)(j * s1)));
            d = fsub<FrP>(x1, x3);
            x1 = fadd<FrP>(x1, x3);
            x3 = fmul<FrP>(d, fload_ro(a.tile_tw + 2 * (size_t)((j + H / 2) * s1)));
            // stage lh - 1: (x0, x1) and (x2, x3), both with w^(2 j s1)
            const fe_t w2 = fload_ro(a.tile_tw + 2 * (size_t)(2 * j * s1));
            d = fsub<FrP>(x0, x1);
            x0 = fadd<FrP>(x0, x1);
            x1 = fmul<FrP>(d, w2);
            d = fsub<FrP>(x2, x3);
            x2 = fadd<FrP>(x2, x3);
            x3 = fmul<FrP>(d, w2);
            sm_store(s_lo, s_hi, e0, x0);
            sm_store(s_lo, s_hi, e0 + H / 2, x1);
            sm_store(s_lo, s_hi, e0 + H, x2);
            sm_store(s_lo, s_hi, e0 + H + H / 2, x3);
        }
    }
    __syncthreads();

    // ---- last two stages (h = 2, 1) on 4 consecutive elements in registers: one product (w_4), then the outputs
    const fe_t w4 = fload_ro(a.tile_tw + 2 * (size_t)(T >> 2));
#pragma unroll
    for (uint32_t q = 0; q < 2; ++q) {
        const uint32_t g = tid + NT * q;
        fe_t x0 = sm_load(s_lo, s_hi, 4 * g), x1 = sm_load(s_lo, s_hi, 4 * g + 1);
        fe_t x2 = sm_load(s_lo, s_hi, 4 * g + 2), x3 = sm_load(s_lo, s_hi, 4 * g + 3);
        bfly(x0, x2);
        bfly(x1, x3);
        x3 = fmul<FrP>(x3, w4);
        bfly(x0, x1);
        bfly(x2, x3);
        emit_output(a, out, c, r, base, __brev(4 * g) >> (32 - LT), x0);
        emit_output(a, out, c, r, base, __brev(4 * g + 1) >> (32 - LT), x1);
        emit_output(a, out, c, r, base, __brev(4 * g + 2) >> (32 - LT), x2);
        emit_output(a, out, c, r, base, __brev(4 * g + 3) >> (32 - LT), x3);
    }
}

template <int LT>
int launch_r4(zkb_ctx *ctx, const PassArgs &a, size_t tiles, size_t count) {
    constexpr size_t T = (size_t)1 << LT;
    const size_t smem = 2 * (T + T / 8 + 1) * 16;
    static bool configured = false;
    if (!configured) {
        ZKB_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel_r4<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        configured = true;
    }
    ntt_pass_kernel_r4<LT><<<dim3((unsigned)tiles, (unsigned)count), (unsigned)(T / 8), smem, ctx->stream>>>(a);
    return ZKB_OK;
}

// out[j] = scale * base^(j << shift)
__global__ void pow_table_kernel(uint4 *out, fe_t base, fe_t scale, uint32_t count, uint32_t shift) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    unsigned long long e = (unsigned long long)j << shift;
    fe_t acc = scale, b = base;
    while (e) {
        if (e & 1) acc = fmul<FrP>(acc, b);
        b = fsqr<FrP>(b);
        e >>= 1;
    }
    fstore(out + 2 * (size_t)j, acc);
}

// out[c * T + k] = w_M^(c*k) for c < S, k < T (M = T*S), expanded from the two-level table
__global__ void expand_twiddle_kernel(uint4 *out, const uint4 *tw2, uint32_t tw_s, uint32_t t, unsigned long long count) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long c = i >> t, k = i & ((1ull << t) - 1);
    fstore(out + 2 * i, pow2lvl(tw2, tw_s, c * k));
}

// the same with the tile's coset factor folded in: out[c * T + k] = w_M^(c*k) * g^c
__global__ void expand_twiddle_coset_kernel(uint4 *out, const uint4 *tw2, uint32_t tw_s, const uint4 *cs2, uint32_t cs_s, uint32_t t,
                                            unsigned long long count) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long c = i >> t, k = i & ((1ull << t) - 1);
    fstore(out + 2 * i, fmul<FrP>(pow2lvl(tw2, tw_s, c * k), pow2lvl(cs2, cs_s, c)));
}

// out[e] = base^e (pre-scaled) for e < count, expanded from the two-level table
__global__ void expand_powers_kernel(uint4 *out, const uint4 *tab, uint32_t s, unsigned long long count) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fstore(out + 2 * i, pow2lvl(tab, s, i));
}

fe_t to_dev(const host::Fe &f) {
    fe_t r;
    memcpy(r.v, f.l, 32);
    return r;
}

int build_pow_table(zkb_ctx *ctx, uint4 *out, const host::Fe &base, const host::Fe &scale, uint32_t count, uint32_t shift) {
    pow_table_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(out, to_dev(base), to_dev(scale), count, shift);
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

// kind 1: tile table w_T^{+-j}, j < T/2
int get_tile_table(zkb_ctx *ctx, unsigned t, int inverse, const uint4 **out) {
    uint64_t key = (1ull << 32) | (t << 1) | (unsigned)inverse;
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        DevBuf b;
        uint32_t cnt = t ? (1u << (t - 1)) : 1;
        int rc = zkb_reserve(ctx, b, (size_t)cnt * 32);
        if (rc) return rc;
        host::Fe w = host::fr_root_of_unity(t);
        if (inverse) w = host::inv(w, host::FR);
        rc = build_pow_table(ctx, (uint4 *)b.p, w, host::one(host::FR), cnt, 0);
        if (rc) return rc;
        it = ctx->tables.emplace(key, b).first;
    }
    *out = (const uint4 *)it->second.p;
    return ZKB_OK;
}

// two-level table of base^e, e < 2^lm: lo[e & (2^s - 1)] * hi[e >> s]; hi optionally pre-scaled
int get_2lvl_table(zkb_ctx *ctx, uint64_t key, unsigned lm, const host::Fe &base, const host::Fe &hi_scale,
                   const uint4 **out, uint32_t *s_out) {
    uint32_t s = (lm + 1) / 2;
    *s_out = s;
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        DevBuf b;
        uint32_t nlo = 1u << s, nhi = 1u << (lm - s);
        int rc = zkb_reserve(ctx, b, (size_t)(nlo + nhi) * 32);
        if (rc) return rc;
        rc = build_pow_table(ctx, (uint4 *)b.p, base, host::one(host::FR), nlo, 0);
        if (rc) return rc;
        rc = build_pow_table(ctx, (uint4 *)b.p + 2 * (size_t)nlo, base, hi_scale, nhi, s);
        if (rc) return rc;
        it = ctx->tables.emplace(key, b).first;
    }
    *out = (const uint4 *)it->second.p;
    return ZKB_OK;
}

// Direct (fully expanded) tables trade HBM for one field product per element per use; kept for transforms up to
// 2^DIRECT_MAX_LOG elements (N x 32 B each), cached like the small tables.
constexpr unsigned DIRECT_MAX_LOG = 26;

int get_direct_table(zkb_ctx *ctx, uint64_t key, unsigned long long count, const uint4 *src, uint32_t s, int twiddle_t,
                     const uint4 **out) {
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        DevBuf b;
        int rc = zkb_reserve(ctx, b, (size_t)count * 32);
        if (rc) return rc;
        unsigned blocks = (unsigned)((count + 255) / 256);
        if (twiddle_t >= 0) expand_twiddle_kernel<<<blocks, 256, 0, ctx->stream>>>((uint4 *)b.p, src, s, (uint32_t)twiddle_t, count);
        else expand_powers_kernel<<<blocks, 256, 0, ctx->stream>>>((uint4 *)b.p, src, s, count);
        ZKB_CUDA(ctx, cudaGetLastError());
        it = ctx->tables.emplace(key, b).first;
    }
    *out = (const uint4 *)it->second.p;
    return ZKB_OK;
}

}  // namespace

// shared with poly.cu: cached / caller-owned two-level power tables (lo[2^s] then hi[2^(lm-s)], s = (lm+1)/2)
int zkb_pow2lvl_cached(zkb_ctx *ctx, uint64_t key, unsigned lm, const zkb::host::Fe &base, const zkb::host::Fe &hi_scale,
                       const void **out, uint32_t *s_out) {
    const uint4 *t = nullptr;
    int rc = get_2lvl_table(ctx, key, lm, base, hi_scale, &t, s_out);
    *out = t;
    return rc;
}

int zkb_pow2lvl_build(zkb_ctx *ctx, void *out, unsigned lm, const zkb::host::Fe &base, const zkb::host::Fe &hi_scale,
                      uint32_t *s_out) {
    uint32_t s = (lm + 1) / 2, nlo = 1u << s, nhi = 1u << (lm - s);
    *s_out = s;
    int rc = build_pow_table(ctx, (uint4 *)out, base, host::one(host::FR), nlo, 0);
    if (rc) return rc;
    return build_pow_table(ctx, (uint4 *)out + 2 * (size_t)nlo, base, hi_scale, nhi, s);
}

int zkb_ntt_run(zkb_ctx *ctx, uint64_t *d_data, size_t len, unsigned log_n, int inverse, int coset) {
    return zkb_ntt_run_batch(ctx, &d_data, 1, len, log_n, inverse, coset);
}

// `count` transforms of the same shape in one launch per pass (blockIdx.y selects the transform): a single 2^20 pass is
// 512 tiles = 1.15 waves of the 444 resident CTAs, nine of them are 10.4
int zkb_ntt_run_batch(zkb_ctx *ctx, uint64_t *const *d_ptrs, size_t count, size_t len, unsigned log_n, int inverse, int coset) {
    if (!count) return ZKB_OK;
    if (!d_ptrs) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ntt: null pointer table");
    for (size_t k = 0; k < count; ++k) if (!d_ptrs[k]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ntt: null data pointer");
    if (count > NTT_MAX_BATCH) {
        for (size_t k = 0; k < count; k += NTT_MAX_BATCH) {
            int rc = zkb_ntt_run_batch(ctx, d_ptrs + k, std::min<size_t>(NTT_MAX_BATCH, count - k), len, log_n, inverse, coset);
            if (rc) return rc;
        }
        return ZKB_OK;
    }
    if (log_n > host::FR_TWO_ADICITY || log_n > 31) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_ntt: log_n exceeds Fr TWO_ADICITY (28 on BN254) or 31");
    const size_t n = (size_t)1 << log_n;
    if (len > n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ntt: len > 2^log_n");
    inverse = inverse ? 1 : 0;
    coset = coset ? 1 : 0;

    // split log_n into <= 3 passes of <= 11 bits
    unsigned bits[3] = {0, 0, 0}, m;
    if (log_n <= 11) { m = 1; bits[0] = log_n; }
    else if (log_n <= 22) { m = 2; bits[0] = (log_n + 1) / 2; bits[1] = log_n - bits[0]; }
    else { m = 3; bits[0] = (log_n + 2) / 3; bits[1] = (log_n - bits[0] + 1) / 2; bits[2] = log_n - bits[0] - bits[1]; }

    if (m > 1) {
        int rc = zkb_reserve(ctx, ctx->ntt_scratch, count * n * 32);
        if (rc) return rc;
    }

    ZKB_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));

    const uint4 *cs2 = nullptr;
    uint32_t cs_s = 0;
    host::Fe ninv = host::inv(host::from_u64((uint64_t)n, host::FR), host::FR);
    if (coset) {
        host::Fe g = host::from_u64(host::FR_GENERATOR, host::FR);
        uint64_t key = (3ull << 32) | (log_n << 1) | (unsigned)inverse;
        int rc = inverse ? get_2lvl_table(ctx, key, log_n, host::inv(g, host::FR), ninv, &cs2, &cs_s)
                         : get_2lvl_table(ctx, key, log_n, g, host::one(host::FR), &cs2, &cs_s);
        if (rc) return rc;
    }
    const bool direct = log_n >= 12 && log_n <= DIRECT_MAX_LOG && !ctx->ntt_no_direct;
    // forward coset transforms of two or more passes fold the coset factor (coset_in_factor): no expanded coset table then
    const bool fold_coset = direct && coset && !inverse && m >= 2 && !ctx->ntt_no_fold;
    const uint4 *cs_direct = nullptr;
    if (coset && direct && !fold_coset) {
        uint64_t key = (5ull << 32) | (log_n << 1) | (unsigned)inverse;
        int rc = get_direct_table(ctx, key, n, cs2, cs_s, -1, &cs_direct);
        if (rc) return rc;
    }

    unsigned log_s = log_n;
    for (unsigned p = 0; p < m; ++p) {
        const unsigned t = bits[p];
        log_s -= t;
        PassArgs a;
        memset(&a, 0, sizeof a);
        const bool first = p == 0, last = p == m - 1;
        for (size_t k = 0; k < count; ++k) a.data[k] = (uint4 *)d_ptrs[k];
        a.scr = (uint4 *)ctx->ntt_scratch.p;
        a.scr_stride = 2 * (unsigned long long)n;                 // in uint4 units
        a.len = first ? len : n;
        a.t = t; a.log_s = log_s; a.log_n = log_n;
        a.t1 = m >= 2 ? bits[0] : 0;
        a.t2 = m == 3 ? bits[1] : 0;
        int rc = get_tile_table(ctx, t, inverse, &a.tile_tw);
        if (rc) return rc;
        if (!last) {
            unsigned lm = t + log_s;             // enclosing sub-transform size M = T*S
            host::Fe w = host::fr_root_of_unity(lm);
            if (inverse) w = host::inv(w, host::FR);
            uint64_t key = (2ull << 32) | (lm << 1) | (unsigned)inverse;
            rc = get_2lvl_table(ctx, key, lm, w, host::one(host::FR), &a.tw2, &a.tw_s);
            if (rc) return rc;
            if (direct && first && coset && !inverse && fold_coset) {
                // forward coset transform, first pass: (g^S)^j at the load, g^c in the inter-pass table (coset_in_factor)
                uint64_t ckey = (6ull << 32) | ((uint64_t)t << 16) | (log_n << 1);
                auto it = ctx->tables.find(ckey);
                if (it == ctx->tables.end()) {
                    DevBuf b;
                    rc = zkb_reserve(ctx, b, ((size_t)1 << t) * 32);
                    if (rc) return rc;
                    rc = build_pow_table(ctx, (uint4 *)b.p, host::from_u64(host::FR_GENERATOR, host::FR), host::one(host::FR), 1u << t, log_s);
                    if (rc) return rc;
                    it = ctx->tables.emplace(ckey, b).first;
                }
                a.cs_tile = (const uint4 *)it->second.p;
                uint64_t dkey = (7ull << 32) | ((uint64_t)t << 16) | (lm << 1);
                it = ctx->tables.find(dkey);
                if (it == ctx->tables.end()) {
                    DevBuf b;
                    rc = zkb_reserve(ctx, b, ((size_t)1 << lm) * 32);
                    if (rc) return rc;
                    const unsigned long long cnt = 1ull << lm;
                    expand_twiddle_coset_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>((uint4 *)b.p, a.tw2, a.tw_s, cs2, cs_s, t, cnt);
                    ZKB_CUDA(ctx, cudaGetLastError());
                    it = ctx->tables.emplace(dkey, b).first;
                }
                a.tw_direct = (const uint4 *)it->second.p;
            } else if (direct) {
                uint64_t dkey = (4ull << 32) | ((uint64_t)t << 16) | (lm << 1) | (unsigned)inverse;
                rc = get_direct_table(ctx, dkey, 1ull << lm, a.tw2, a.tw_s, (int)t, &a.tw_direct);
                if (rc) return rc;
            }
        }
        a.cs_direct = cs_direct;
        a.cs2 = cs2; a.cs_s = cs_s;
        a.mode = (last ? M_LAST : 0) | (first ? M_FIRST : 0);
        if (first && coset && !inverse) a.mode |= M_IN_COSET;
        if (last && inverse) a.mode |= coset ? M_OUT_COSET : M_OUT_CONST;
        a.scale = to_dev(ninv);
        const size_t T = (size_t)1 << t;
        // one thread per 8 elements: every thread is busy in the register tail and runs 4 butterflies per stage
        unsigned threads = (unsigned)(T / 8 < 32 ? 32 : (T / 8 > 256 ? 256 : T / 8));
        size_t tiles = n >> t;
        size_t smem = 2 * (T + T / 8 + 1) * 16;
        // measured (profiles/r02d_ntt_*.jsonl): the radix-4 kernel wins on 2048-element tiles (2^22: 0.94 vs 1.05 ms, nine at
        // once 7.1 vs 8.4 ms) and loses on smaller ones (2^20: 0.30 vs 0.26 ms) -- also when it is built for the generic
        // kernel's occupancy (128 registers, 512 threads per SM, 84 KB of L1 left: profiles/r02o_ntt_kernel3.jsonl), so it is
        // not the L1 that its 216 KB of shared memory per SM take away
        const bool r4_all = ctx->ntt_kernel == 2;
        rc = ZKB_OK;
        if (t == 11 && ctx->ntt_kernel != 1) rc = launch_r4<11>(ctx, a, tiles, count);
        else if (r4_all && t == 10) rc = launch_r4<10>(ctx, a, tiles, count);
        else if (r4_all && t == 9) rc = launch_r4<9>(ctx, a, tiles, count);
        else if (r4_all && t == 8) rc = launch_r4<8>(ctx, a, tiles, count);
        else ntt_pass_kernel<<<dim3((unsigned)tiles, (unsigned)count), threads, smem, ctx->stream>>>(a);
        if (rc) return rc;
        ctx->launches += 1;
        ZKB_CUDA(ctx, cudaGetLastError());
    }
    return ZKB_OK;
}
