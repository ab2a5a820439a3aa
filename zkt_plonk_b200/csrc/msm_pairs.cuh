// msm_pairs.cuh -- batched-affine PAIR ROUNDS in front of the XYZZ bucket accumulation (included by msm.cu only).
//
// The bucket method's inner work is N * W point additions.  In XYZZ coordinates one mixed addition costs 8M + 2S (1232
// multiply-adds as implemented); in affine coordinates it costs 2M + 1S plus ONE field inversion, and Montgomery's trick
// shares an inversion among a batch at 3M per member: 5M + 1S = 788 multiply-adds per addition if the inversion itself is
// amortised away and kept off the multiplier pipe.  What makes that possible here:
//
//   * after the counting sort the entries of one bucket are contiguous, so "entry 2k + entry 2k+1 of every bucket" is a set
//     of fully INDEPENDENT additions: a round halves every bucket (an odd entry is carried over by reference), and a
//     thread can take any run of m consecutive pairs of the round -- no per-bucket accumulator, no conflicts, no locks;
//   * a thread owns m ~ 32..128 pairs: prefix products of the m denominators x2 - x1 go to HBM (32 B per pair; the B200 has
//     the bandwidth and the capacity to spare: this path trades ~320 B of traffic per addition for 36 % of the multiplier
//     work), ONE inversion per thread and batch, then the additions in reverse order;
//   * the inversion is a binary extended Euclid on the ALU pipe (ff_inv.cuh), which the multiplier-bound additions leave
//     idle: it costs issue slots of other warps, no IMAD slots;
//   * R rounds remove 1 - 2^-R of the additions (R = 3: 87 %); what is left per bucket (a few points, now in a pool of
//     intermediate sums) goes through the unchanged XYZZ accumulation, window reduction and fold.
//
// References ("refs") are 32-bit: bit 31 = negate (signed digit), bit 30 = the point lives in the pool of intermediate
// sums rather than in the base / table array, bits 0..29 = index.  Round r reads refs_r / counts_r / starts_r and writes
// refs_{r+1} / counts_{r+1} / starts_{r+1}; the sums of all rounds stay in the pool (a carried-over odd entry of round r
// may be read many rounds later).  Special pairs (an operand at infinity, P = Q, P = -Q) are recognised from the x
// coordinates alone -- the same predicate in both phases -- kept out of the batch, and resolved by the complete XYZZ
// formulas; affine outputs are canonical, so the results are bit-identical to the XYZZ-only path.
#pragma once
#include "ec.cuh"
#if ZKB_CURVE == ZKB_CURVE_BN254
#include "ff_inv.cuh"
#endif

namespace zkb {
namespace pairs {

constexpr uint32_t SIGN = 0x80000000u, POOL = 0x40000000u, ID_MASK = 0x3fffffffu;
constexpr int THREADS = 128;
constexpr int MAX_ROUNDS = 6;
constexpr uint32_t M_MIN = 32, M_MAX = 128;
constexpr uint32_t PF = 4;               // L2 prefetch distance, in pairs

__device__ __forceinline__ const g1a_t *ref_ptr(const g1a_t *points, const g1a_t *pool, uint32_t ref) {
    return ((ref & POOL) ? pool : points) + (ref & ID_MASK);
}
// `prefetch.global.L2` pulls the whole 128-byte line: exact for a pair of adjacent pool entries (rounds >= 1), but for 64-byte
// table points gathered at random (round 0) it doubles the DRAM traffic (measured: 6.4 GB read per 6.7 M pairs).  The sized
// alternative, cp.async.bulk.prefetch.L2, takes a warp-uniform address: ptxas serialises it over the 32 lanes (a waterfall
// loop of ~7 instructions per lane), which costs more than it saves.
__device__ __forceinline__ void prefetch_x(const g1a_t *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_xy(const g1a_t *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p) + 32));
}
__device__ __forceinline__ void prefetch_32(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The pair rounds (an opt-in mode, off by default) exist for the 8-limb base field only: ff_inv.cuh's Euclid and the register
// budget of pair_add_kernel are written for BN254.  On the BLS12 builds a plan never has rounds (msm.cu pick_pair_rounds).
#if ZKB_CURVE == ZKB_CURVE_BN254
#define ZKB_HAVE_PAIR_ROUNDS 1
// ------------------------------------------------------------------ per-round bookkeeping
// pk[b] = pairs of bucket b | entries of bucket b after the round << 32, b < nb; pk[nb] = 0 (the scan's total lands there)
__global__ void __launch_bounds__(256) pack_kernel(const uint32_t *__restrict__ counts, uint32_t nb, unsigned long long *__restrict__ pk) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    uint32_t c = b < nb ? counts[b] : 0;
    pk[b] = (unsigned long long)(c >> 1) | ((unsigned long long)((c + 1) >> 1) << 32);
}

constexpr uint32_t SCAN_TILE64 = 2048;   // 256 threads x 8
__device__ __forceinline__ unsigned long long block_exclusive_scan64(unsigned long long v, unsigned long long *sm, unsigned long long *total) {
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) sm[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned long long s = lane < 8 ? sm[lane] : 0;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        if (lane < 8) sm[8 + lane] = s;
    }
    __syncthreads();
    unsigned long long warp_off = wid ? sm[8 + wid - 1] : 0;
    *total = sm[15];
    unsigned long long r = warp_off + x - v;
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(256) scan64_reduce_kernel(const unsigned long long *in, uint32_t n, unsigned long long *tile_sums) {
    __shared__ unsigned long long sm[16];
    uint32_t base = blockIdx.x * SCAN_TILE64 + threadIdx.x * 8;
    unsigned long long s = 0, total;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (base + k < n) s += in[base + k];
    block_exclusive_scan64(s, sm, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) scan64_sums_kernel(unsigned long long *tile_sums, uint32_t ntiles) {
    __shared__ unsigned long long sm[16];
    unsigned long long running = 0;
    for (uint32_t base = 0; base < ntiles; base += 256) {
        uint32_t i = base + threadIdx.x;
        unsigned long long v = i < ntiles ? tile_sums[i] : 0, total;
        unsigned long long ex = block_exclusive_scan64(v, sm, &total);
        if (i < ntiles) tile_sums[i] = running + ex;
        running += total;
    }
}
__global__ void __launch_bounds__(256) scan64_apply_kernel(unsigned long long *data, uint32_t n, const unsigned long long *tile_sums) {
    __shared__ unsigned long long sm[16];
    uint32_t base = blockIdx.x * SCAN_TILE64 + threadIdx.x * 8;
    unsigned long long v[8], s = 0, total;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = base + k < n ? data[base + k] : 0; s += v[k]; }
    unsigned long long ex = block_exclusive_scan64(s, sm, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (base + k < n) data[base + k] = ex; ex += v[k]; }
}

// after the scan: sc[b] = (pairs before bucket b, entries of the next round before bucket b).  Next round's counts and
// starts, and the carried-over odd entry of every bucket.
__global__ void __launch_bounds__(256) finish_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ starts,
                                                     const uint32_t *__restrict__ refs, const uint2 *__restrict__ sc, uint32_t nb,
                                                     uint32_t *__restrict__ ncounts, uint32_t *__restrict__ nstarts, uint32_t *__restrict__ nrefs) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t c = counts[b], ns = sc[b].y;
    ncounts[b] = (c + 1) >> 1;
    nstarts[b] = ns;
    if (c & 1) nrefs[ns + (c >> 1)] = refs[starts[b] + c - 1];
}

// ------------------------------------------------------------------ the pair additions
// complete addition of two affine points (any special case), affine result: rare path
static __device__ __noinline__ g1a_t add_slow(const g1a_t &p, const g1a_t &q) {
    g1x_t acc = g1x_from_affine(p);
    g1x_add_mixed(acc, q);
    return g1x_to_affine(acc);
}
__device__ __forceinline__ bool special_pair(const fe_t &x1, const fe_t &x2, const fe_t &d) {
    // an operand at infinity is (0, 0); x = 0 with y != 0 would be a genuine point and is merely sent down the slow path
    return fis_zero<Q>(x1) || fis_zero<Q>(x2) || fis_zero<Q>(d);
}

// Thread t of the grid owns pairs [t * m, t * m + m) of this round.  Pair j of the round is entries (2k, 2k + 1) of bucket
// b where sc[b].x <= j < sc[b + 1].x and k = j - sc[b].x; its sum is pool[out_base + j] and becomes entry k of bucket b
// in the next round.
__global__ void __launch_bounds__(THREADS, 4) pair_add_kernel(const g1a_t *__restrict__ points, g1a_t *__restrict__ pool,
                                                              const uint32_t *__restrict__ refs, const uint32_t *__restrict__ starts,
                                                              const uint2 *__restrict__ sc, uint32_t nb, uint32_t m, uint32_t out_base,
                                                              uint32_t *__restrict__ nrefs, uint2 *__restrict__ pairrefs,
                                                              fe_t *__restrict__ prefix) {
    const uint32_t T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const uint32_t total = sc[nb].x;
    if ((unsigned long long)(t - lane) * m >= total) return;      // the whole warp is past the end (warp-uniform)
    const unsigned long long j0l = (unsigned long long)t * m;
    const uint32_t j0 = j0l < total ? (uint32_t)j0l : total, cnt = min(m, total - j0);   // cnt == 0: only helps with the inversion
    uint2 *my_refs = pairrefs + t;                               // element i at my_refs[i * T]
    fe_t *my_pre = prefix + t;

    // ---- phase 1a: which entries form my pairs (metadata only)
    if (cnt) {
        uint32_t lo = 0, hi = nb;                                // largest b with sc[b].x <= j0
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sc[mid].x <= j0) lo = mid; else hi = mid;
        }
        uint32_t b = lo;
        uint2 cur = sc[b], nxt = sc[b + 1];
        uint32_t k = j0 - cur.x, pb = nxt.x - cur.x, sb = starts[b];
        uint32_t i = 0;
        while (i < cnt) {
            if (k == pb) {                                       // next bucket that has pairs
                ++b;
                cur = nxt;
                nxt = sc[b + 1];
                pb = nxt.x - cur.x;
                k = 0;
                if (pb) sb = starts[b];
                continue;
            }
            const uint32_t len = min(pb - k, cnt - i);
            const uint32_t *src = refs + sb + 2 * k;
            uint32_t *dst = nrefs + cur.y + k;
#pragma unroll 4
            for (uint32_t u = 0; u < len; ++u) {
                my_refs[(size_t)(i + u) * T] = make_uint2(src[2 * u], src[2 * u + 1]);
                dst[u] = POOL | (out_base + j0 + i + u);
            }
            i += len;
            k += len;
        }
    }

    // Odd warps cut their pairs into two batches (one third, two thirds), even warps run one: every batch ends phase 1 with an
    // inversion that keeps its warp off the multiplier for tens of microseconds, and warps that started together would all
    // invert together (measured: 12 % of a round's stall samples).  Staggered, one half of an SM's warps multiplies while the
    // other inverts.
    const fe_t one = fone<Q>();
    const uint32_t nbatch = ((threadIdx.x >> 5) & 1) ? 2 : 1;
    for (uint32_t bt = 0; bt < nbatch; ++bt) {
        const uint32_t lo = nbatch == 1 ? 0 : (bt == 0 ? 0 : cnt / 3), hi = nbatch == 1 ? cnt : (bt == 0 ? cnt / 3 : cnt);
        const uint32_t len = hi - lo;

        // ---- phase 1b: running product of the denominators of pairs [lo, hi).  Loads run two pairs ahead in registers (x
        // coordinates only) and the references three ahead; no L2 prefetch instruction (it pulls 128-byte lines).
        fe_t run = one;
        if (len) {
            uint2 r1 = my_refs[(size_t)lo * T], r2 = len > 1 ? my_refs[(size_t)(lo + 1) * T] : r1, r3 = len > 2 ? my_refs[(size_t)(lo + 2) * T] : r1;
            fe_t xa1 = fload_ro(&ref_ptr(points, pool, r1.x)->x), xa2 = fload_ro(&ref_ptr(points, pool, r1.y)->x);
            fe_t xb1 = fload_ro(&ref_ptr(points, pool, r2.x)->x), xb2 = fload_ro(&ref_ptr(points, pool, r2.y)->x);
            for (uint32_t i = lo; i < hi; ++i) {
                const uint2 r4 = i + 3 < hi ? my_refs[(size_t)(i + 3) * T] : r3;
                fe_t xc1 = xb1, xc2 = xb2;
                if (i + 2 < hi) {
                    xc1 = fload_ro(&ref_ptr(points, pool, r3.x)->x);
                    xc2 = fload_ro(&ref_ptr(points, pool, r3.y)->x);
                }
                const fe_t d = fsub<Q>(xa2, xa1);
                if (!special_pair(xa1, xa2, d)) run = fmul<Q>(run, d);
                fstore(&my_pre[(size_t)i * T], run);
                xa1 = xb1; xa2 = xb2; xb1 = xc1; xb2 = xc2;
                r3 = r4;
            }
        }

        // ---- ONE inversion per WARP and batch: the 32 lane totals are multiplied up by two shuffle scans (prefix and suffix
        // products, 5 steps each), lane 0 inverts the warp total (binary extended Euclid, ALU pipe: the other warps of the SM
        // keep the multiplier busy), and 1 / T_l = (1 / T_0..31) * T_0..l-1 * T_l+1..31.  A lock-step inversion in all 32
        // lanes would cost the same issue slots per lane and diverge in its data-dependent loops (measured: 0.5 ms per round).
        fe_t inv;
        {
            fe_t pfx = run, sfx = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                fe_t up, dn;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    up.v[k] = __shfl_up_sync(0xffffffffu, pfx.v[k], d);
                    dn.v[k] = __shfl_down_sync(0xffffffffu, sfx.v[k], d);
                }
                if (lane >= (uint32_t)d) pfx = fmul<Q>(pfx, up);
                if (lane + d < 32) sfx = fmul<Q>(sfx, dn);
            }
            fe_t tot;                                            // T_0 * ... * T_31 sits in lane 31's prefix
#pragma unroll
            for (int k = 0; k < 8; ++k) tot.v[k] = __shfl_sync(0xffffffffu, pfx.v[k], 31);
            if (lane == 0) tot = finv_euclid<Q>(tot);
            fe_t before, after;                                  // exclusive prefix / suffix products of my lane
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                tot.v[k] = __shfl_sync(0xffffffffu, tot.v[k], 0);
                before.v[k] = __shfl_up_sync(0xffffffffu, pfx.v[k], 1);
                after.v[k] = __shfl_down_sync(0xffffffffu, sfx.v[k], 1);
            }
            inv = tot;
            if (lane > 0) inv = fmul<Q>(inv, before);
            if (lane < 31) inv = fmul<Q>(inv, after);
        }

        // ---- phase 2: the additions of pairs [lo, hi), last pair first.  The next pair's points and prefix product are
        // loaded (into registers) in the middle of the current addition, once lambda is known and Py, Qy are dead.
        if (len) {
            uint2 pr = my_refs[(size_t)(hi - 1) * T];
            g1a_t p = g1a_load(ref_ptr(points, pool, pr.x)), q = g1a_load(ref_ptr(points, pool, pr.y));
            fe_t pre = hi - 1 > lo ? fload(&my_pre[(size_t)(hi - 2) * T]) : one;
            uint2 npr = hi - 1 > lo ? my_refs[(size_t)(hi - 2) * T] : pr;
            for (uint32_t i = hi; i-- > lo;) {
                // (pool sums were written by earlier rounds = earlier launches, table points never change: read-only path)
                if (pr.x & SIGN) p.y = fneg<Q>(p.y);
                if (pr.y & SIGN) q.y = fneg<Q>(q.y);
                const fe_t d = fsub<Q>(q.x, p.x);
                const bool more = i > lo;
                const uint2 nnpr = i > lo + 1 ? my_refs[(size_t)(i - 2) * T] : npr;
                g1a_t r, np = p, nq = q;
                fe_t npre = one;
                if (special_pair(p.x, q.x, d)) {
                    r = add_slow(p, q);
                    if (more) {
                        np = g1a_load(ref_ptr(points, pool, npr.x));
                        nq = g1a_load(ref_ptr(points, pool, npr.y));
                        if (i > lo + 1) npre = fload(&my_pre[(size_t)(i - 2) * T]);
                    }
                } else {
                    const fe_t inv_d = fmul<Q>(inv, pre);
                    inv = fmul<Q>(inv, d);
                    const fe_t lam = fmul<Q>(fsub<Q>(q.y, p.y), inv_d);
                    const fe_t px = p.x, py = p.y, qx = q.x;
                    if (more) {                                  // the next pair is in flight during the two products below
                        np = g1a_load(ref_ptr(points, pool, npr.x));
                        nq = g1a_load(ref_ptr(points, pool, npr.y));
                        if (i > lo + 1) npre = fload(&my_pre[(size_t)(i - 2) * T]);
                    }
                    r.x = fsub<Q>(fsub<Q>(fsqr<Q>(lam), px), qx);
                    r.y = fsub<Q>(fmul<Q>(lam, fsub<Q>(px, r.x)), py);
                }
                fstore(&pool[out_base + j0 + i].x, r.x);
                fstore(&pool[out_base + j0 + i].y, r.y);
                p = np; q = nq; pre = npre; pr = npr; npr = nnpr;
            }
        }
    }
}

#else
#define ZKB_HAVE_PAIR_ROUNDS 0
#endif

// ------------------------------------------------------------------ host side
struct Plan {
    uint32_t rounds = 0;
    uint64_t e_ub[MAX_ROUNDS + 1];       // upper bound of the entries going into round r
    uint64_t p_ub[MAX_ROUNDS];           // upper bound of the pairs of round r
    uint64_t pool_base[MAX_ROUNDS + 1];  // static offset of round r's sums in the pool
    uint32_t m[MAX_ROUNDS], grid[MAX_ROUNDS];
    uint64_t scratch_elems = 0;          // pairrefs / prefix capacity (max over rounds of m * grid * THREADS)
};

// rounds: pair rounds wanted (<= MAX_ROUNDS); entries = n * W upper bound; nb = buckets; resident = threads of one full wave
inline Plan make_plan(uint32_t rounds, uint64_t entries, uint64_t nb, uint64_t resident) {
    Plan pl;
    pl.rounds = rounds;
    pl.e_ub[0] = entries;
    pl.pool_base[0] = 0;
    for (uint32_t r = 0; r < rounds; ++r) {
        pl.p_ub[r] = pl.e_ub[r] / 2;
        const uint64_t nonempty = pl.e_ub[r] < nb ? pl.e_ub[r] : nb;
        pl.e_ub[r + 1] = (pl.e_ub[r] + nonempty + 1) / 2;        // sum of ceil(c / 2)
        if (pl.e_ub[r + 1] > pl.e_ub[r]) pl.e_ub[r + 1] = pl.e_ub[r];
        pl.pool_base[r + 1] = pl.pool_base[r] + pl.p_ub[r];
        uint64_t m = (pl.p_ub[r] + resident - 1) / resident;     // one full wave if possible
        if (m < M_MIN) m = M_MIN;
        if (m > M_MAX) m = M_MAX;
        pl.m[r] = (uint32_t)m;
        pl.grid[r] = (uint32_t)((pl.p_ub[r] + m * THREADS - 1) / (m * THREADS));
        if (pl.grid[r] == 0) pl.grid[r] = 1;
        const uint64_t cap = (uint64_t)pl.m[r] * pl.grid[r] * THREADS;
        if (cap > pl.scratch_elems) pl.scratch_elems = cap;
    }
    return pl;
}

struct Ws {
    uint32_t *refs[2], *counts[2], *starts[2];
    unsigned long long *pk, *scan_tmp;
    uint2 *pairrefs;
    fq_t *prefix;
    g1a_t *pool;
};

}  // namespace pairs
}  // namespace zkb
