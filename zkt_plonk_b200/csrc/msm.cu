// msm.cu -- Pippenger bucket-method G1 multi-scalar multiplication for sm_100a.
//
// Replaces ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` (called from plonk-core/src/commitment.rs:42 and,
// through ark-poly-commit's kzg10::commit / open, from prove.rs:134,179,250,307,374,381,427).  The reference
// uses unsigned c-bit windows, one rayon task per window and serial bucket loops; results are only canonical
// as affine points, so this implementation is free to differ in everything but the final (x, y):
//
//   1. signed-digit windows (digits in [-2^(c-1), 2^(c-1)], half the buckets), c chosen by a cost model;
//   2. counting sort of (bucket, point|sign) pairs with global atomics (order inside a bucket is irrelevant
//      because the group is commutative);
//   3. buckets cut into tasks of <= SEG points; tasks are ordered by size so that every warp runs 32
//      equally long loops; one thread accumulates one task in XYZZ coordinates (8M + 2S per point, the
//      next point's gather is issued before the current addition);
//   4. tasks of oversized buckets (skewed scalars: zeros/ones/small values) are combined warp-cooperatively;
//   5. per-group weighted bucket sums sum_j (j+1) * bucket_j over a dense bucket array, by index-digit recursion:
//      with j = R*t + b,  sum_j j*X[j] = sum_t A[t] + R * sum_t t*S[t]  where S[t] = sum_b X[R*t+b] and
//      A[t] = sum_b b*X[R*t+b].  Level l turns X_l into X_{l+1} = S (R times shorter) and a new plain stream A_l,
//      and shortens the older plain streams A_0..A_{l-1} by plain R-to-1 sums.  Every level is one launch of fully
//      independent short chains (R = 8 while the arrays are long, R = 2 -- depth one -- for the tail), there is
//      no scalar multiplication and no serial fold on the device;
//   6. the level totals (<= W x ~20 x 128 B) go to the host, which finishes with a Horner fold (log2 R doublings per
//      level, then c doublings per window) and converts to affine -- a few hundred group operations, cheaper there
//      than on one GPU thread.
//
// Fixed-base mode (zkb_srs_precompute): for the resident SRS the table T[w][i] = 2^(c*w) * P_i is built once
// (W x N x 64 B of HBM).  Then every window feeds ONE shared set of 2^(c-1) buckets (entry id = w*N + i), a much
// larger c pays off (c = 20 at 2^20: 13 windows instead of 16, i.e. 19 % fewer bucket insertions), and the
// host-side Horner fold disappears.  This trades HBM capacity (180 GB) for integer work.
//
// The integer pipe (IMAD.WIDE) bounds step 3; everything else is a few percent of the time.
#include "ctx.h"
#include "ec.cuh"
#include "host_ff.h"
#include "msm_pairs.cuh"

#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <array>
#include <vector>

using namespace zkb;

namespace {

constexpr uint32_t SEG_MAX = 1024;       // upper bound of the per-MSM task length `seg` (points per accumulation task)
constexpr uint32_t RED_THREADS = 128;    // threads per CTA in the window reduction
constexpr uint32_t RED_MAX_LEVELS = 26;  // levels of the weighted-sum recursion
constexpr uint32_t RED_CTAS_PER_SM = FqP::N == 8 ? 3 : 2;   // resident CTAs of the wide level (<= 170 registers per thread; 255 with 12-limb coordinates)
constexpr uint32_t ACC_CTAS_PER_SM = FqP::N == 8 ? 4 : 2;   // bucket accumulation: 128 registers per thread; 12-limb coordinates: 218 at
                                                            // 2 CTAs per SM -- measured against 168 registers (76 bytes spilled) at 3 CTAs
                                                            // per SM: 5.95 vs 6.18 ms at 2^20 on BLS12-381 (more warps do not pay for the spills)
constexpr uint32_t SIGN_BIT = 0x80000000u;
// widths at the C boundary, in 64-bit words: an affine point (x, y) and an un-normalised XYZZ partial sum (8 / 16 on BN254,
// 12 / 24 on the BLS12 curves: zkb_curve_info reports them)
constexpr int AFF_W = 2 * host::FQ_L, XYZZ_W = 4 * host::FQ_L;
constexpr size_t AFF_BYTES = 8 * AFF_W, XYZZ_BYTES = 8 * XYZZ_W;
constexpr uint32_t SCALAR_BITS = FrP::BITS;   // 254 (BN254), 255 (BLS12-381), 253 (BLS12-377)

struct MsmPlan {
    uint32_t c, W, B;                    // window bits, windows, buckets per group (2^(c-1))
    uint32_t G;                          // bucket groups: W (one per window) or 1 (fixed-base tables)
    uint32_t wide;                       // windows [0, wide) are c bits wide, the others c - 1 (wide = W: uniform)
    uint64_t nbuckets;                   // G * B
    uint32_t red_levels;                 // weighted-sum recursion: levels, radix and input length of each
    uint32_t red_r[RED_MAX_LEVELS];
    uint32_t red_m[RED_MAX_LEVELS + 1];
    uint64_t red_buf_elems[2];           // ping-pong buffers of the recursion (levels alternate)
    uint32_t seg;                        // max points per accumulation task
    uint32_t id_base, id_stride;         // fixed-base: entry id = id_base + w * id_stride + i
    pairs::Plan pp;                      // pair rounds in front of the XYZZ accumulation (pp.rounds == 0: none)
};

struct MsmWs {                           // carved out of ctx->msm_ws
    uint32_t *counts, *starts, *cursor, *ntasks, *task_base, *sorted;
    uint32_t *scan_tmp;                  // block sums for the scans
    uint32_t *size_hist, *size_cursor;   // SEG + 1 bins (the cursors are the histogram after its in-place scan)
    size_t zero_bytes;                   // counts, size_hist and misc are adjacent: bytes cleared per MSM, from counts on
    uint32_t *misc;                      // [0] heavy buckets, [1] total tasks, [2] chunk items, [3] chunk_out slots, [4] multi-chunk buckets
    uint32_t *heavy_list;                // buckets cut into more than one task
    uint32_t *heavy_slot;                // per heavy bucket: first slot in chunk_out, or ~0 when it has a single chunk
    uint32_t *multi_list;                // heavy-list indices of the buckets with more than one chunk
    uint2 *chunk_items;                  // (heavy-list index, chunk): one warp folds up to HEAVY_CHUNK task results
    g1x_t *chunk_out;                    // chunk sums of multi-chunk buckets
    uint2 *task_order;
    g1x_t *task_out;                     // partial sums of the tasks of multi-task buckets
    g1x_t *bucket_val;                   // dense: one XYZZ value per bucket (zero = empty)
    g1x_t *red_buf[2];                   // level outputs, [stream][group][t]
    pairs::Ws pw;                        // batched-affine pair rounds (msm_pairs.cuh); unused when the plan has no rounds
};

struct MsmSlot {                         // one MSM in flight: its own workspace, result buffer and completion event
    DevBuf ws;
    void *pinned = nullptr;              // pinned host buffer for the group sums
    size_t pinned_bytes = 0;
    cudaEvent_t acc_done = nullptr, tail_done = nullptr;
};

struct MsmState {
    MsmSlot slot[2];                     // slot 0: single MSMs; slots 0/1 alternate in pipelined batches
    cudaStream_t tail_stream = nullptr;  // high-priority stream for the latency-bound tail of a pipelined MSM
    cudaStream_t copy_stream = nullptr;  // uploads of host scalars, overlapped with the MSM of the previous part
    cudaEvent_t part_uploaded[8] = {};
    static constexpr int MAX_PARTS = 8;
    cudaEvent_t ev[MAX_PARTS][5] = {};   // phase boundaries of the last MSM, per part (a plain MSM is one part)
    int ev_parts = 1;                    // parts of the last MSM
    cudaStream_t sort_stream = nullptr;  // multi-part MSMs: the sort phase of part k + 1 runs here under the accumulation of part k
    cudaEvent_t part_sorted[MAX_PARTS] = {}, part_acc[MAX_PARTS] = {};
    DevBuf shared_buckets;               // multi-part MSMs: the dense bucket array all parts accumulate into
    cudaEvent_t ev_pairs[2] = {nullptr, nullptr};                        // around the pair rounds of the last MSM
    uint32_t last_rounds = 0;
    bool ev_valid = false;
    uint64_t last_entries = 0;           // n * W upper bound of bucket insertions of the last MSM
    uint32_t last_c = 0, last_W = 0;
    void *fixed_base = nullptr;          // FixedBase* (fixed-base window tables of the resident SRS)
    // open batch of pipelined commitments (zkb_commit_push / zkb_commit_finish)
    std::vector<std::array<uint64_t, XYZZ_W>> pipe_partial;   // folded XYZZ partial sum of every pushed commitment
    std::vector<char> pipe_pending;      // 1: enqueued, result still in its slot's pinned buffer
    size_t pipe_expected = 1;            // commitments the open batch will hold (zkb_commit_expect): decides the multi-GPU layout
    MsmPlan pipe_plan[2];
};

// ------------------------------------------------------------------ digits
// Signed c-bit digits of a canonical 254-bit scalar.  Calls f(w, bucket_in_window, negative) for non-zero digits.
// Window w covers bits [start_w, start_w + width_w): the first `wide` windows are c bits wide, the rest c - 1
// (wide = W: uniform c-bit windows).  Balanced widths keep every window's digit spread over the whole bucket range.
template <class F>
__device__ __forceinline__ void for_each_digit(const uint32_t (&s)[8], uint32_t c, uint32_t W, uint32_t wide, F f, uint32_t w0 = 0) {
    uint32_t carry = 0, bit = 0;
    for (uint32_t w = 0; w < W; ++w) {
        const uint32_t width = w < wide ? c : c - 1;
        const uint32_t half = 1u << (width - 1), full = 1u << width;
        uint32_t limb = bit >> 5, off = bit & 31;
        uint32_t lo = limb < 8 ? s[limb] : 0, hi = limb + 1 < 8 ? s[limb + 1] : 0;
        uint32_t raw = (uint32_t)((((uint64_t)hi << 32) | lo) >> off) & (full - 1);
        bit += width;
        raw += carry;
        if (raw > half) {
            carry = 1;                            // digit = raw - 2^c <= 0
            if (raw != full && w >= w0) f(w, full - raw - 1, true);
        } else {
            carry = 0;
            if (raw && w >= w0) f(w, raw - 1, false);
        }
    }
}

// digit of window 0 alone (no incoming carry): returns false when it is zero
__device__ __forceinline__ bool first_digit(const uint32_t (&s)[8], uint32_t c, uint32_t wide, uint32_t *bucket, bool *neg) {
    const uint32_t width = wide > 0 ? c : c - 1;
    const uint32_t half = 1u << (width - 1), full = 1u << width;
    const uint32_t raw = s[0] & (full - 1);                       // width <= 24 bits: inside the first limb
    if (raw > half) { *bucket = full - raw - 1; *neg = true; return true; }   // raw < full always (no carry in)
    *bucket = raw - 1;
    *neg = false;
    return raw != 0;
}

__device__ __forceinline__ void load_scalar(const uint4 *p, size_t i, uint32_t (&s)[8]) {
    uint4 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// Window 0 is where skewed scalars collide: every scalar equal to one (and every small one) has its only digit there, so
// a witness-like vector sends 20 % of its entries to ONE counter.  Lanes of a warp that hit the same bucket are found
// with match.any and served by one atomic (the leader adds the group's size and hands out ranks); the other windows'
// digits are spread and go one atomic each.
__device__ __forceinline__ uint32_t warp_aggregated_add(uint32_t *counters, uint32_t bucket, bool has, uint32_t *rank) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t peers = __match_any_sync(0xffffffffu, has ? bucket : 0xffffffffu);     // lanes without a digit group together
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (has && lane == leader) base = atomicAdd(&counters[bucket], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    *rank = __popc(peers & ((1u << lane) - 1));
    return base;
}

// gstride = B when every window has its own bucket group, 0 when all windows share one (fixed-base tables)
__global__ void msm_count_kernel(const uint4 *scalars, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, uint32_t gstride, uint32_t *counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (i < n) load_scalar(scalars, i, s);                        // lanes past the end keep s = 0: no digits
    uint32_t b0 = 0, rank;
    bool neg0 = false;
    const bool has0 = first_digit(s, c, wide, &b0, &neg0);
    warp_aggregated_add(counts, b0, has0, &rank);
    for_each_digit(s, c, W, wide, [&](uint32_t w, uint32_t b, bool) { atomicAdd(&counts[w * gstride + b], 1u); }, 1);
}

// entry id = id_base + w * id_stride + i  (id_stride = 0: plain bases; = SRS size: fixed-base table rows)
__global__ void msm_scatter_kernel(const uint4 *scalars, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, uint32_t gstride,
                                   uint32_t id_base, uint32_t id_stride, uint32_t *cursor, uint32_t *sorted) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (i < n) load_scalar(scalars, i, s);
    uint32_t b0 = 0, rank;
    bool neg0 = false;
    const bool has0 = first_digit(s, c, wide, &b0, &neg0);
    const uint32_t base0 = warp_aggregated_add(cursor, b0, has0, &rank);
    if (has0) sorted[base0 + rank] = (id_base + i) | (neg0 ? SIGN_BIT : 0u);
    for_each_digit(s, c, W, wide, [&](uint32_t w, uint32_t b, bool neg) {
        uint32_t slot = atomicAdd(&cursor[w * gstride + b], 1u);
        sorted[slot] = (id_base + w * id_stride + i) | (neg ? SIGN_BIT : 0u);
    }, 1);
}

// ------------------------------------------------------------------ exclusive scan of uint32 (three phases)
constexpr uint32_t SCAN_TILE = 2048;     // elements per CTA (256 threads x 8)

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *sm, uint32_t *total) {
    // 256 threads; returns exclusive prefix of v, *total = sum over the block
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) sm[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < 8 ? sm[lane] : 0;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        if (lane < 8) sm[8 + lane] = s;   // inclusive warp totals
    }
    __syncthreads();
    uint32_t warp_off = wid ? sm[8 + wid - 1] : 0;
    *total = sm[15];
    uint32_t r = warp_off + x - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) scan_reduce_kernel(const uint32_t *in, uint32_t n, uint32_t *tile_sums) {
    __shared__ uint32_t sm[16];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * 8, s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan_256(s, sm, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of tile_sums in place; total written to *total_out (may be null)
__global__ void __launch_bounds__(256) scan_sums_kernel(uint32_t *tile_sums, uint32_t ntiles, uint32_t *total_out) {
    __shared__ uint32_t sm[16];
    uint32_t running = 0;
    for (uint32_t base = 0; base < ntiles; base += 256) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0, total;
        uint32_t ex = block_exclusive_scan_256(v, sm, &total);
        if (i < ntiles) tile_sums[i] = running + ex;
        running += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = running;
}

// out2 (may be null): a second copy of the result (the sort's write cursors start as a copy of the bucket offsets)
__global__ void __launch_bounds__(256) scan_apply_kernel(const uint32_t *in, uint32_t n, const uint32_t *tile_sums, uint32_t *out, uint32_t *out2) {
    __shared__ uint32_t sm[16];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * 8, v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    uint32_t total;
    uint32_t ex = block_exclusive_scan_256(s, sm, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (base + k < n) { out[base + k] = ex; if (out2) out2[base + k] = ex; }
        ex += v[k];
    }
}

int exclusive_scan(zkb_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tmp, uint32_t *total_out, uint32_t *out2 = nullptr) {
    uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_reduce_kernel<<<ntiles, 256, 0, ctx->stream>>>(in, n, tmp);
    scan_sums_kernel<<<1, 256, 0, ctx->stream>>>(tmp, ntiles, total_out);
    scan_apply_kernel<<<ntiles, 256, 0, ctx->stream>>>(in, n, tmp, out, out2);
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

// ------------------------------------------------------------------ tasks
constexpr uint32_t HEAVY_CHUNK = 64;     // task results folded by one warp (two per lane, then a shuffle tree)

__global__ void msm_ntasks_kernel(const uint32_t *counts, uint32_t nb, uint32_t SEG, uint32_t *ntasks, uint32_t *size_hist,
                                  uint32_t *misc, uint32_t *heavy_list, uint32_t *heavy_slot, uint32_t *multi_list, uint2 *chunk_items) {
    __shared__ uint32_t h[SEG_MAX + 1];
    for (uint32_t k = threadIdx.x; k <= SEG; k += blockDim.x) h[k] = 0;
    __syncthreads();
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) {
        uint32_t cnt = counts[b];
        uint32_t nt = (cnt + SEG - 1) / SEG;
        ntasks[b] = nt;
        if (nt) {
            uint32_t full = cnt / SEG, rem = cnt - full * SEG;
            if (full) atomicAdd(&h[0], full);               // bin = SEG - size (descending size order)
            if (rem) atomicAdd(&h[SEG - rem], 1u);
            if (nt > 1) {                                   // oversized bucket: its task results are folded in chunks of HEAVY_CHUNK
                const uint32_t h = atomicAdd(&misc[0], 1u), chunks = (nt + HEAVY_CHUNK - 1) / HEAVY_CHUNK;
                heavy_list[h] = b;
                const uint32_t c0 = atomicAdd(&misc[2], chunks);
                for (uint32_t k = 0; k < chunks; ++k) chunk_items[c0 + k] = make_uint2(h, k);
                if (chunks > 1) {
                    heavy_slot[h] = atomicAdd(&misc[3], chunks);
                    multi_list[atomicAdd(&misc[4], 1u)] = h;
                } else {
                    heavy_slot[h] = 0xffffffffu;
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= SEG; k += blockDim.x) if (h[k]) atomicAdd(&size_hist[k], h[k]);
}

// Tasks ordered by descending size.  A CTA ranks its tasks per size bin in shared memory and reserves one global
// range per (CTA, bin), so the global atomics are few and spread (a per-task atomic on ~40 hot bins cost 100 us).
__global__ void __launch_bounds__(256) msm_task_scatter_kernel(const uint32_t *counts, const uint32_t *ntasks, uint32_t nb, uint32_t SEG,
                                                               uint32_t *size_cursor, uint2 *task_order) {
    __shared__ uint32_t h[SEG_MAX + 1], base[SEG_MAX + 1];
    for (uint32_t k = threadIdx.x; k <= SEG; k += blockDim.x) h[k] = 0;
    __syncthreads();
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nt = b < nb ? ntasks[b] : 0, cnt = nt ? counts[b] : 0;
    uint32_t full = cnt / SEG, rem = cnt - full * SEG;
    uint32_t rank_full = 0, rank_rem = 0;
    if (full) rank_full = atomicAdd(&h[0], full);             // bin = SEG - size
    if (rem) rank_rem = atomicAdd(&h[SEG - rem], 1u);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= SEG; k += blockDim.x) base[k] = h[k] ? atomicAdd(&size_cursor[k], h[k]) : 0;
    __syncthreads();
    for (uint32_t s = 0; s < full; ++s) task_order[base[0] + rank_full + s] = make_uint2(b, s);
    if (rem) task_order[base[SEG - rem] + rank_rem] = make_uint2(b, full);
}

// ------------------------------------------------------------------ bucket accumulation (the hot kernel)
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p) + 32));
}

__global__ void __launch_bounds__(128, ACC_CTAS_PER_SM) msm_accumulate_kernel(const g1a_t *__restrict__ points, const g1a_t *__restrict__ pool,
                                                             const uint32_t *__restrict__ sorted,
                                                             const uint32_t *__restrict__ counts, const uint32_t *__restrict__ starts,
                                                             const uint32_t *__restrict__ ntasks, const uint32_t *__restrict__ task_base,
                                                             const uint2 *__restrict__ task_order, const uint32_t *__restrict__ misc,
                                                             uint32_t SEG, g1x_t *__restrict__ task_out, g1x_t *__restrict__ bucket_val,
                                                             uint32_t PF, uint32_t accum) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= misc[1]) return;
    uint2 task = task_order[t];
    uint32_t b = task.x, s = task.y;
    uint32_t cnt = min(SEG, counts[b] - s * SEG);
    const uint32_t *idx = sorted + starts[b] + s * SEG;

    // accum: a later part of a multi-part MSM (msm_run_parts) -- the bucket already holds the sum of the earlier parts
    const bool single = ntasks[b] == 1;
    g1x_t acc = g1x_inf();
    if (accum && single) acc = g1x_load(bucket_val + b);
    // Gather pipeline: register load one point ahead, add the current point.  One addition of a warp takes ~10 us with four
    // warps per scheduler, ten times an HBM access, so the load issued one iteration ahead always arrives in time; the
    // optional L2 prefetch PF points ahead (round 1's default, PF = 6) buys nothing and `prefetch.global.L2` pulls whole
    // 128-byte lines for 64-byte points: measured 1.894 ms without it against 1.905 with it at 2^20
    // (profiles/r02k_msm_prefetch_distance.log), and half the DRAM traffic.
    // (PF = L2 prefetch distance in points, 0 = none: a kernel argument so that its effect on time and DRAM traffic can be measured)
    // an entry is a point of the base / table array or (bit 30, after pair rounds) a sum in the pool of intermediate results
    for (uint32_t k = 1; k < PF && k < cnt; ++k) prefetch_l2(pairs::ref_ptr(points, pool, idx[k]));
    uint32_t v = idx[0];
    g1a_t p = g1a_load(pairs::ref_ptr(points, pool, v));
    for (uint32_t k = 0; k < cnt; ++k) {
        uint32_t vn = 0;
        g1a_t pn;
        if (k + PF < cnt) prefetch_l2(pairs::ref_ptr(points, pool, idx[k + PF]));
        if (k + 1 < cnt) {                                   // issue the next gather before the addition
            vn = idx[k + 1];
            pn = g1a_load(pairs::ref_ptr(points, pool, vn));
        }
        if (v & SIGN_BIT) p.y = fneg<FqP>(p.y);
        g1x_add_mixed(acc, p);
        v = vn;
        p = pn;
    }
    g1x_store(single ? bucket_val + b : task_out + task_base[b] + s, acc);
}

// one warp per oversized bucket: lanes stride over the bucket's task results, shuffle tree at the end
__device__ __forceinline__ g1x_t shfl_down_g1x(const g1x_t &p, int d) {
    g1x_t r;
#pragma unroll
    for (int i = 0; i < FqP::N; ++i) {
        r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], d);
        r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], d);
        r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], d);
        r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], d);
    }
    return r;
}

// Oversized buckets (skewed scalars: ones / small values pile up on a few digits), two stages of one warp per work item.
// Stage 1: a warp folds one chunk of <= HEAVY_CHUNK task results of a bucket into the bucket's value (single chunk) or into a
// chunk sum.  Stage 2: a warp folds the chunk sums of a multi-chunk bucket.  A 200k-entry bucket cut into 64-entry tasks is 3300 tasks -> 52 chunks -> one value, each
// stage ~30 us, instead of one CTA walking all task results.
// `len` points at base[0 .. len) folded by one warp: lanes stride over them, then a shuffle tree.  (Four cooperating lanes
// per addition -- the 14 products of an XYZZ addition as 4 rounds of one product per lane -- were measured here and in the
// binary tail of the window reduction and dropped: a lone warp is bound by its instruction issue latency, and the exchange
// by shuffles costs what the shorter product chain saves; profiles/r02f, r02g.)
__device__ __forceinline__ void warp_fold_store(const g1x_t *base, uint32_t len, g1x_t *dst, bool add_dst = false) {
    const uint32_t lane = threadIdx.x & 31;
    g1x_t acc = g1x_inf();
    for (uint32_t k = lane; k < len; k += 32) g1x_add(acc, g1x_load(base + k));
    for (int d = 16; d >= 1; d >>= 1) {
        if ((uint32_t)d >= len) continue;                          // uniform across the warp
        g1x_t o = shfl_down_g1x(acc, d);
        if (lane < (uint32_t)d) g1x_add(acc, o);
    }
    __syncwarp();
    if (lane == 0) {
        if (add_dst) g1x_add(acc, g1x_load(dst));                  // multi-part MSM: the bucket's sum over the earlier parts
        g1x_store(dst, acc);
    }
}

__global__ void __launch_bounds__(128) msm_combine_chunks_kernel(const uint32_t *misc, const uint32_t *heavy_list, const uint32_t *heavy_slot,
                                                                 const uint2 *chunk_items, const uint32_t *ntasks, const uint32_t *task_base,
                                                                 const g1x_t *task_out, g1x_t *chunk_out, g1x_t *bucket_val, uint32_t accum) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nitems = misc[2];
    for (uint32_t it = warp; it < nitems; it += nwarps) {
        const uint2 item = chunk_items[it];
        const uint32_t b = heavy_list[item.x], nt = ntasks[b], lo = item.y * HEAVY_CHUNK, len = min(HEAVY_CHUNK, nt - lo);
        const uint32_t slot = heavy_slot[item.x];
        warp_fold_store(task_out + task_base[b] + lo, len, slot == 0xffffffffu ? bucket_val + b : chunk_out + slot + item.y,
                        accum && slot == 0xffffffffu);
    }
}

__global__ void __launch_bounds__(128) msm_combine_final_kernel(const uint32_t *misc, const uint32_t *heavy_list, const uint32_t *heavy_slot,
                                                                const uint32_t *multi_list, const uint32_t *ntasks, const g1x_t *chunk_out,
                                                                g1x_t *bucket_val, uint32_t accum) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nmulti = misc[4];
    for (uint32_t it = warp; it < nmulti; it += nwarps) {
        const uint32_t h = multi_list[it], b = heavy_list[h], chunks = (ntasks[b] + HEAVY_CHUNK - 1) / HEAVY_CHUNK;
        warp_fold_store(chunk_out + heavy_slot[h], chunks, bucket_val + b, accum != 0);
    }
}

// ------------------------------------------------------------------ per-group weighted bucket sums (index-digit recursion)
// One level.  in: [n_plain + 1 streams][G groups][M_in], out: [n_plain + 2][G][M_out], M_out = ceil(M_in / R).
//   stream 0 (X):            out[0][g][t] = sum_b X[R*t+b]   and   out[n_plain+1][g][t] = sum_b b * X[R*t+b]  (the new A stream)
//   stream s >= 1 (plain A): out[s][g][t] = sum_b in[s][g][R*t+b]
// Then sum_j j*X[j] = sum_t A[t] + R * sum_t t*out[0][t]: the next level continues on out[0].  Elements past M_in read
// as the identity.  grid = (ceil(M_out / 128), n_plain + 1, G).
template <bool WIDE>
__global__ void __launch_bounds__(RED_THREADS, WIDE ? RED_CTAS_PER_SM : 1) msm_wsum_level_kernel(const g1x_t *__restrict__ in, uint32_t M_in,
                                                                                                uint32_t M_out, uint32_t R, uint32_t n_plain,
                                                                                                g1x_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * RED_THREADS + threadIdx.x;
    if (t >= M_out) return;
    const uint32_t strm = blockIdx.y, g = blockIdx.z, G = gridDim.z;
    const g1x_t *src = in + ((size_t)strm * G + g) * M_in + (size_t)R * t;
    const uint32_t avail = min(R, M_in - R * t);                   // >= 1
    g1x_t *dst = out + ((size_t)strm * G + g) * M_out + t;
    if (!WIDE) {                                                  // R == 2: chains of depth one
        g1x_t x0 = g1x_load(src);
        if (avail > 1) {
            g1x_t x1 = g1x_load(src + 1);
            if (strm == 0) g1x_store(out + ((size_t)(n_plain + 1) * G + g) * M_out + t, x1);
            g1x_add(x0, x1);
        } else if (strm == 0) {
            g1x_store(out + ((size_t)(n_plain + 1) * G + g) * M_out + t, g1x_inf());
        }
        g1x_store(dst, x0);
        return;
    }
    if (strm != 0) {                                              // plain R-to-1 sum
        g1x_t acc = g1x_load(src), cur = acc;
        if (avail > 1) cur = g1x_load(src + 1);
        for (uint32_t b = 1; b < avail; ++b) {
            g1x_t nxt = cur;
            if (b + 1 < avail) nxt = g1x_load(src + b + 1);        // next element in flight during the addition
            g1x_add(acc, cur);
            cur = nxt;
        }
        g1x_store(dst, acc);
        return;
    }
    // weighted: run = X[b..], acc = sum_{k >= b} (k - b + 1) X[k], walking b from the top down to 1
    g1x_t run = g1x_inf(), acc = g1x_inf();
    g1x_t cur = g1x_load(src + avail - 1);
    for (uint32_t b = avail - 1; b >= 1; --b) {
        g1x_t nxt = g1x_load(src + b - 1);
        g1x_add(run, cur);
        g1x_add(acc, run);
        cur = nxt;
    }
    g1x_add(run, cur);                                            // + X[0]: the plain sum
    g1x_store(dst, run);
    g1x_store(out + ((size_t)(n_plain + 1) * G + g) * M_out + t, acc);
}

// A binary level with FOUR lanes per addition.  A level's additions are independent but each is one dependent chain of 14
// field products -- ~9 us in a warp of its own -- and from the fifth level on a level is a single CTA or less, so the
// sixteen levels cost their additions' LATENCY (0.18 of the 0.43 ms of a 2^19-bucket reduction).  The products of an
// XYZZ addition form four rounds of at most four independent ones; lane l of a group computes one product per round, the
// results move by shuffles (9 field elements per addition), and lane l stores coordinate l of the sum:
//   round 1   u1 = X1 ZZ2      u2 = X2 ZZ1     s1 = Y1 ZZZ2     s2 = Y2 ZZZ1        p = u2 - u1, r = s2 - s1
//   round 2   pp = p^2         zz = ZZ1 ZZ2    rr = r^2         zzz = ZZZ1 ZZZ2
//   round 3   ppp = p pp       q = u1 pp       ZZ3 = zz pp      --                   X3 = rr - ppp - 2 q
//   round 4   --               r (q - X3)      s1 ppp           ZZZ3 = zzz ppp       Y3 = r (q - X3) - s1 ppp
// 16 lane-products instead of 14, a quarter of the depth.  Infinity and p == 0 (doubling / cancellation) are uniform over a
// group and take the one-lane formulas.  Same sums as msm_wsum_level_kernel<false> (canonical field elements either way).
__device__ __forceinline__ fq_t shfl_fe(unsigned mask, const fq_t &a, int src) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < FqP::N; ++i) r.v[i] = __shfl_sync(mask, a.v[i], src);
    return r;
}
__device__ __forceinline__ fq_t shfl_xor_fe(unsigned mask, const fq_t &a, int m) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < FqP::N; ++i) r.v[i] = __shfl_xor_sync(mask, a.v[i], m);
    return r;
}
__device__ __forceinline__ fq_t sel_fe(bool c, const fq_t &a, const fq_t &b) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < FqP::N; ++i) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

__global__ void __launch_bounds__(RED_THREADS) msm_wsum_level_coop_kernel(const g1x_t *__restrict__ in, uint32_t M_in, uint32_t M_out,
                                                                         uint32_t n_plain, g1x_t *__restrict__ out) {
    const uint32_t t = (blockIdx.x * RED_THREADS + threadIdx.x) >> 2, l = threadIdx.x & 3, base = (threadIdx.x & 31) & ~3u;
    const uint32_t strm = blockIdx.y, g = blockIdx.z, G = gridDim.z;
    const unsigned m0 = __ballot_sync(0xffffffffu, t < M_out);    // whole groups leave; the masks name the lanes that stay
    if (t >= M_out) return;
    const g1x_t *src = in + ((size_t)strm * G + g) * M_in + (size_t)2 * t;
    const char *p0 = reinterpret_cast<const char *>(src), *p1 = p0 + sizeof(g1x_t);
    char *dst = reinterpret_cast<char *>(out + ((size_t)strm * G + g) * M_out + t);
    const bool pair = M_in - 2 * t > 1;
    if (strm == 0)                                                // the new A stream: X[2t + 1], one coordinate per lane
        fstore(reinterpret_cast<char *>(out + ((size_t)(n_plain + 1) * G + g) * M_out + t) + FQ_BYTES * l, pair ? floadn<FqP::N>(p1 + FQ_BYTES * l) : fzero<Q>());
    // coordinates: 0 x, 1 y, 2 zz, 3 zzz.  lane 0: X1, ZZ2; lane 1: X2, ZZ1; lane 2: Y1, ZZZ2; lane 3: Y2, ZZZ1
    const fq_t a = pair || !(l & 1) ? floadn<FqP::N>(((l & 1) ? p1 : p0) + FQ_BYTES * (l >> 1)) : fzero<Q>();
    const fq_t b = pair ? floadn<FqP::N>(((l & 1) ? p0 : p1) + 2 * FQ_BYTES + FQ_BYTES * (l >> 1)) : fzero<Q>();
    // infinity on either side (ZZ == 0; a missing partner counts as infinity): the sum is the other point
    const bool inf2 = __shfl_sync(m0, (int)fis_zero<Q>(b), base) != 0;          // ZZ2 sits on lane 0
    const bool inf1 = __shfl_sync(m0, (int)fis_zero<Q>(b), base + 1) != 0;      // ZZ1 on lane 1
    const unsigned m1 = __ballot_sync(m0, !(inf1 || inf2));
    if (inf1 || inf2) {
        fstore(dst + FQ_BYTES * l, inf2 ? floadn<FqP::N>(p0 + FQ_BYTES * l) : floadn<FqP::N>(p1 + FQ_BYTES * l));
        return;
    }
    fq_t prod = fmul<Q>(a, b);                                                   // u1 | u2 | s1 | s2
    const fq_t other = shfl_xor_fe(m1, prod, 1);
    const fq_t lo = sel_fe(l & 1, other, prod);                                  // u1 (lanes 0, 1), s1 (lanes 2, 3)
    const fq_t d = fsub<Q>(sel_fe(l & 1, prod, other), lo);                      // p  (lanes 0, 1), r  (lanes 2, 3)
    const bool pz = __shfl_sync(m1, (int)fis_zero<Q>(d), base) != 0;
    const unsigned m2 = __ballot_sync(m1, !pz);
    if (pz) {                                                                    // same x: every lane runs the complete formulas
        g1x_t x0 = g1x_load(p0);
        g1x_add(x0, g1x_load(p1));
        fstore(dst + FQ_BYTES * l, l == 0 ? x0.x : l == 1 ? x0.y : l == 2 ? x0.zz : x0.zzz);
        return;
    }
    const fq_t bp = shfl_xor_fe(m2, b, 1);                                       // lane 1: ZZ2, lane 3: ZZZ2
    const fq_t prod2 = fmul<Q>(sel_fe(l & 1, b, d), sel_fe(l & 1, bp, d));       // pp | zz | rr | zzz
    const fq_t pp = shfl_fe(m2, prod2, base), zz = shfl_fe(m2, prod2, base + 1);
    const fq_t prod3 = fmul<Q>(l == 0 ? d : l == 1 ? lo : l == 2 ? zz : pp, pp); // ppp | q | ZZ3 | (unused)
    const fq_t ppp = shfl_fe(m2, prod3, base), q = shfl_fe(m2, prod3, base + 1), rr = shfl_fe(m2, prod2, base + 2);
    const fq_t r = shfl_fe(m2, d, base + 2);
    const fq_t x3 = fsub<Q>(fsub<Q>(fsub<Q>(rr, ppp), q), q);
    const fq_t prod4 = fmul<Q>(l == 1 ? r : l == 2 ? lo : l == 3 ? prod2 : ppp, l == 1 ? fsub<Q>(q, x3) : ppp);   // -- | r (q - X3) | s1 ppp | ZZZ3
    const fq_t t1 = shfl_fe(m2, prod4, base + 2);
    fstore(dst + FQ_BYTES * l, l == 0 ? x3 : l == 1 ? fsub<Q>(prod4, t1) : l == 2 ? prod3 : prod4);
}

// Fixed-base table: rows[w][i] = 2^(c*w) * P_i (affine), w < W.  One thread per point walks the windows.
__global__ void __launch_bounds__(128) msm_precompute_kernel(const g1a_t *__restrict__ points, uint32_t n, uint32_t c, uint32_t W,
                                                             uint32_t wide, g1a_t *__restrict__ rows) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1a_t p = g1a_load(points + i);
    fstore(&rows[i].x, p.x);
    fstore(&rows[i].y, p.y);
    g1x_t acc = g1x_from_affine(p);
    for (uint32_t w = 1; w < W; ++w) {
        const uint32_t width = (w - 1) < wide ? c : c - 1;      // width of the window below this row
        for (uint32_t k = 0; k < width; ++k) acc = g1x_double(acc);
        g1a_t a = g1x_to_affine(acc);
        fstore(&rows[(size_t)w * n + i].x, a.x);
        fstore(&rows[(size_t)w * n + i].y, a.y);
    }
}

// ------------------------------------------------------------------ helpers: SRS generation, scalar conversion
// out[i] = scalars[i] * base (affine), plain double-and-add.  Mirrors what PC::setup does once per SRS
// (powers_of_g = [tau^i] G); used to build synthetic SRS / test points directly in HBM.
__global__ void __launch_bounds__(128) g1_fixed_base_mul_kernel(g1a_t base, const uint4 *scalars, uint32_t n, g1a_t *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(scalars, i, s);
    g1x_t acc = g1x_inf();
    for (int bit = 255; bit >= 0; --bit) {
        acc = g1x_double(acc);
        if ((s[bit >> 5] >> (bit & 31)) & 1) g1x_add_mixed(acc, base);
    }
    g1a_t a = g1x_to_affine(acc);
    fstore(&out[i].x, a.x);
    fstore(&out[i].y, a.y);
}

// Fr Montgomery -> canonical (what into_repr() does before kzg10::commit hands scalars to the MSM)
__global__ void fr_from_mont_kernel(const uint4 *in, uint4 *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fstore(out + 2 * (size_t)i, ffrom_mont<FrP>(fload_ro(in + 2 * (size_t)i)));
}

// ------------------------------------------------------------------ host-side group law (window fold + affine)
namespace hec {
typedef host::Fq Fe;                     // base-field element (4 or 6 x u64)
using host::FQ;
struct Pt { Fe x, y, zz, zzz; };
static_assert(sizeof(Pt) == XYZZ_BYTES && sizeof(g1x_t) == XYZZ_BYTES && sizeof(g1a_t) == AFF_BYTES, "point layouts");
inline bool is_inf(const Pt &p) { return host::is_zero(p.zz); }
inline Pt inf() { Pt p; memset(&p, 0, sizeof p); return p; }
inline Pt dbl(const Pt &p) {
    if (is_inf(p)) return p;
    Pt r;
    Fe u = host::add(p.y, p.y, FQ), v = host::sqr(u, FQ), w = host::mul(u, v, FQ), s = host::mul(p.x, v, FQ);
    Fe xx = host::sqr(p.x, FQ), m = host::add(host::add(xx, xx, FQ), xx, FQ);
    r.x = host::sub(host::sub(host::sqr(m, FQ), s, FQ), s, FQ);
    r.y = host::sub(host::mul(m, host::sub(s, r.x, FQ), FQ), host::mul(w, p.y, FQ), FQ);
    r.zz = host::mul(v, p.zz, FQ);
    r.zzz = host::mul(w, p.zzz, FQ);
    return r;
}
inline Pt add(const Pt &a, const Pt &b) {
    if (is_inf(a)) return b;
    if (is_inf(b)) return a;
    Fe u1 = host::mul(a.x, b.zz, FQ), u2 = host::mul(b.x, a.zz, FQ);
    Fe s1 = host::mul(a.y, b.zzz, FQ), s2 = host::mul(b.y, a.zzz, FQ);
    Fe p = host::sub(u2, u1, FQ), r = host::sub(s2, s1, FQ);
    if (host::is_zero(p)) return host::is_zero(r) ? dbl(a) : inf();
    Fe pp = host::sqr(p, FQ), ppp = host::mul(p, pp, FQ), q = host::mul(u1, pp, FQ);
    Pt o;
    o.x = host::sub(host::sub(host::sub(host::sqr(r, FQ), ppp, FQ), q, FQ), q, FQ);
    o.y = host::sub(host::mul(r, host::sub(q, o.x, FQ), FQ), host::mul(s1, ppp, FQ), FQ);
    o.zz = host::mul(host::mul(a.zz, b.zz, FQ), pp, FQ);
    o.zzz = host::mul(host::mul(a.zzz, b.zzz, FQ), ppp, FQ);
    return o;
}
inline Pt mul_small(const Pt &p, uint32_t k) {                   // k * p, double-and-add from the top bit
    Pt acc = inf();
    for (int bit = 31; bit >= 0; --bit) {
        acc = dbl(acc);
        if ((k >> bit) & 1) acc = add(acc, p);
    }
    return acc;
}
inline Pt mul_scalar(const Pt &p, const uint64_t k[4]) {           // k * p for a canonical 256-bit scalar (host, a few per call)
    Pt acc = inf();
    for (int bit = 255; bit >= 0; --bit) {
        acc = dbl(acc);
        if ((k[bit >> 6] >> (bit & 63)) & 1) acc = add(acc, p);
    }
    return acc;
}
inline void to_affine(const Pt &p, uint64_t *out_xy /* AFF_W words */, int *is_inf_out) {
    if (is_inf(p)) { memset(out_xy, 0, AFF_BYTES); if (is_inf_out) *is_inf_out = 1; return; }
    Fe zi = host::inv(p.zzz, FQ);
    Fe zzi = host::sqr(host::mul(zi, p.zz, FQ), FQ);
    Fe x = host::mul(p.x, zzi, FQ), y = host::mul(p.y, zi, FQ);
    memcpy(out_xy, x.l, AFF_BYTES / 2);
    memcpy(out_xy + AFF_W / 2, y.l, AFF_BYTES / 2);
    if (is_inf_out) *is_inf_out = 0;
}
}  // namespace hec

// ------------------------------------------------------------------ planning / workspace
struct FixedBase {                       // fixed-base tables of the resident SRS (owned by MsmState)
    DevBuf rows;                         // W x n affine points, row w = 2^(c*w) * SRS
    uint32_t c = 0, W = 0, wide = 0;
    size_t n = 0;
};

// Balanced fixed-base windows: W = ceil(255 / c) windows, the first *wide of width c and the rest c - 1, so that
// no window is left with only a few bits (a 2-bit top window would pile n/4 points on each of 4 buckets).
uint32_t balanced_windows(uint32_t c, uint32_t *wide) {
    // widths must sum to the scalar's bit length plus one spare bit (255 on BN254), so that the top window's raw value plus
    // the incoming carry never exceeds half of its range (no carry out of the top window)
    const uint32_t BITS = SCALAR_BITS + 1;
    uint32_t W = (BITS + c - 1) / c;                           // fewest c-bit windows that cover BITS
    uint32_t slack = W * c - BITS;                              // bits to give back by narrowing windows to c - 1
    *wide = W - (slack < W ? slack : W);
    return W;
}

// Cost in units of one bucket insertion (0.155 ns measured): n*W insertions, stretched when the accumulation grid (one
// thread per bucket) is only a few waves of the 148 x 3 x 128 resident threads (measured: x1.95 at 0.58 waves, x1.25
// at 1.15, x1.0 from ~4 waves on), plus a per-bucket charge for the weighted bucket sums (~0.5 ns per bucket with the
// index-digit recursion) and, beyond 2^19 buckets, for the counting sort's atomics (~0.45 ns per bucket).
uint32_t pick_window(size_t n, bool shared_buckets, int sm_count) {
    uint32_t best_c = 8;
    double best = 1e300;
    const double wave = (double)sm_count * 3 * 128;
    for (uint32_t c = 6; c <= 22; ++c) {
        uint32_t wide = 1, W = shared_buckets ? balanced_windows(c, &wide) : SCALAR_BITS / c + 1;
        if (shared_buckets && wide == 0) continue;               // equivalent to c - 1 with uniform windows
        const double buckets = (shared_buckets ? 1.0 : (double)W) * (double)(1u << (c - 1));
        const double waves = buckets / wave;
        const double stretch = waves >= 4.0 ? 1.0 : waves >= 1.0 ? 1.0 + 0.3 / waves : 1.15 / waves;
        const double big = buckets > 524288.0 ? buckets - 524288.0 : 0.0;       // the sort's atomics stop being cache friendly
        double cost = (double)n * W * stretch + 3.4 * buckets + 3.0 * big;
        if (cost < best) { best = cost; best_c = c; }
    }
    return best_c;
}

// pair rounds for an MSM of `entries` bucket insertions over nb buckets.  mode >= 0: that many rounds; mode < 0: automatic --
// a round pays while buckets still hold several entries (it halves them at 788 instead of 1232 multiply-adds per addition,
// for ~6 small launches), so: floor(log2(mean load)) - 1 rounds, none for small MSMs (latency-bound) or when the
// references / the pool would not fit their 30 index bits or a sane share of HBM.
uint32_t pick_pair_rounds(int mode, uint64_t entries, uint64_t nb, uint64_t max_id) {
    if (!ZKB_HAVE_PAIR_ROUNDS || mode == 0 || entries < 2 || max_id >= pairs::POOL || entries >= pairs::POOL || entries > (1ull << 28)) return 0;
    if (mode > 0) return (uint32_t)std::min(mode, pairs::MAX_ROUNDS);
    if (entries < (1ull << 19)) return 0;
    const double mean = (double)entries / (double)std::max<uint64_t>(1, std::min(nb, entries));
    int r = (int)std::floor(std::log2(std::max(mean, 1.0))) - 1;
    return (uint32_t)std::max(0, std::min(r, 4));
}

MsmPlan make_plan(size_t n, int force_c, const FixedBase *fb, size_t offset, int sm_count, int pair_mode = 0) {
    MsmPlan pl;
    if (fb) {
        pl.c = fb->c; pl.W = fb->W; pl.G = 1; pl.wide = fb->wide;
        pl.id_base = (uint32_t)offset; pl.id_stride = (uint32_t)fb->n;
    } else {
        pl.c = force_c > 0 ? (uint32_t)force_c : pick_window(n, false, sm_count);
        pl.W = SCALAR_BITS / pl.c + 1; pl.G = pl.W; pl.wide = pl.W;
        pl.id_base = 0; pl.id_stride = 0;
    }
    pl.B = 1u << (pl.c - 1);
    pl.nbuckets = (uint64_t)pl.G * pl.B;
    // weighted-sum recursion: the first level shortens the bucket array R0-to-1 with R0 chosen so that its grid is one
    // full wave of resident CTAs (throughput-bound: ~2 additions per bucket), every later level is 2-to-1 (chains of
    // depth one: the tail is latency-bound, and depth log2 is the least a tree can have)
    pl.red_levels = 0;
    pl.red_m[0] = pl.B;
    pl.red_buf_elems[0] = pl.red_buf_elems[1] = 1;
    const uint64_t wave = (uint64_t)sm_count * RED_CTAS_PER_SM * RED_THREADS;
    uint32_t r0 = (uint32_t)((pl.nbuckets + wave - 1) / wave);
    if ((uint64_t)pl.G * ((pl.B + r0 - 1) / (r0 ? r0 : 1)) > wave && r0) ++r0;   // per-group rounding may spill over the wave
    while (pl.red_m[pl.red_levels] > 1) {
        const uint32_t l = pl.red_levels, m = pl.red_m[l];
        pl.red_r[l] = (l == 0 && r0 >= 4) ? r0 : 2;
        pl.red_m[l + 1] = (m + pl.red_r[l] - 1) / pl.red_r[l];
        const uint64_t out_elems = (uint64_t)(l + 2) * pl.G * pl.red_m[l + 1];
        pl.red_buf_elems[l & 1] = std::max(pl.red_buf_elems[l & 1], out_elems);
        ++pl.red_levels;
    }
    // task length: at least 64 and 2.5 times the mean bucket load (rounded up to a power of two, capped), so that
    // ordinary buckets stay a single task while an oversized bucket (skewed scalars) is cut into pieces of a few times the
    // typical thread's work: the accumulation ends when its longest task does, and a lone warp needs ~4 us per addition
    // (256-entry tasks kept the witness-like 2^20 MSM at 1.38 ms of accumulation for 43 % of the insertions)
    double mean = (double)n * pl.W / (double)pl.nbuckets;
    uint32_t seg = 64;
    while (seg < 2.5 * mean && seg < SEG_MAX) seg <<= 1;
    pl.seg = seg;
    const uint64_t max_id = fb ? (uint64_t)pl.id_base + (uint64_t)(pl.W - 1) * pl.id_stride + n : n;
    pl.pp = pairs::make_plan(pick_pair_rounds(pair_mode, (uint64_t)n * pl.W, pl.nbuckets, max_id), (uint64_t)n * pl.W, pl.nbuckets,
                             (uint64_t)sm_count * 4 * pairs::THREADS);
    return pl;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int carve_ws(zkb_ctx *ctx, DevBuf &buf, const MsmPlan &pl, size_t n, MsmWs &ws, uint64_t *max_tasks_out, uint64_t *max_heavy_tasks_out) {
    const uint64_t nb = pl.nbuckets;
    const uint64_t entries = (uint64_t)n * pl.W;
    const uint64_t max_tasks = nb + entries / pl.seg + 1;
    const uint64_t max_heavy_tasks = 2 * (entries / pl.seg) + 2;   // tasks of buckets with more than seg points
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    // counts, the size histogram and the counters are adjacent: one memset clears the three
    size_t o_counts = take(nb * 4), o_hist = take((SEG_MAX + 1) * 4), o_misc = take(64), o_zero_end = off,
           o_starts = take(nb * 4), o_cursor = take(nb * 4), o_ntasks = take(nb * 4),
           o_tbase = take(nb * 4), o_sorted = take(entries * 4 + 4), o_scan = take((nb / SCAN_TILE + 2) * 4), o_heavy = take(nb * 4),
           o_hslot = take(nb * 4), o_multi = take(nb * 4), o_citems = take((nb + max_tasks / HEAVY_CHUNK + 2) * 8),
           o_cout = take((max_tasks / (HEAVY_CHUNK / 2) + 4) * sizeof(g1x_t)),
           o_order = take(max_tasks * 8), o_out = take(max_tasks * sizeof(g1x_t)), o_bval = take(nb * sizeof(g1x_t)),
           o_red0 = take(pl.red_buf_elems[0] * sizeof(g1x_t)), o_red1 = take(pl.red_buf_elems[1] * sizeof(g1x_t));
    size_t o_pr[2] = {0, 0}, o_pc[2] = {0, 0}, o_ps[2] = {0, 0}, o_pk = 0, o_pscan = 0, o_prefs = 0, o_ppre = 0, o_pool = 0;
#if ZKB_HAVE_PAIR_ROUNDS
    if (pl.pp.rounds) {
        for (int k = 0; k < 2; ++k) {
            o_pr[k] = take(pl.pp.e_ub[1] * 4 + 4);
            o_pc[k] = take(nb * 4);
            o_ps[k] = take(nb * 4);
        }
        o_pk = take((nb + 1) * 8);
        o_pscan = take(((nb + 1) / pairs::SCAN_TILE64 + 2) * 8);
        o_prefs = take(pl.pp.scratch_elems * sizeof(uint2));
        o_ppre = take(pl.pp.scratch_elems * sizeof(fq_t));
        o_pool = take((pl.pp.pool_base[pl.pp.rounds] + 1) * sizeof(g1a_t));
    }
#endif
    int rc = zkb_reserve(ctx, buf, off);
    if (rc) return rc;
    char *p = (char *)buf.p;
    ws.counts = (uint32_t *)(p + o_counts); ws.starts = (uint32_t *)(p + o_starts); ws.cursor = (uint32_t *)(p + o_cursor);
    ws.ntasks = (uint32_t *)(p + o_ntasks); ws.task_base = (uint32_t *)(p + o_tbase); ws.sorted = (uint32_t *)(p + o_sorted);
    ws.scan_tmp = (uint32_t *)(p + o_scan); ws.size_hist = (uint32_t *)(p + o_hist); ws.size_cursor = ws.size_hist;
    ws.zero_bytes = o_zero_end - o_counts;
    ws.misc = (uint32_t *)(p + o_misc); ws.heavy_list = (uint32_t *)(p + o_heavy); ws.task_order = (uint2 *)(p + o_order);
    ws.heavy_slot = (uint32_t *)(p + o_hslot); ws.multi_list = (uint32_t *)(p + o_multi); ws.chunk_items = (uint2 *)(p + o_citems);
    ws.chunk_out = (g1x_t *)(p + o_cout);
    ws.task_out = (g1x_t *)(p + o_out); ws.bucket_val = (g1x_t *)(p + o_bval);
    ws.red_buf[0] = (g1x_t *)(p + o_red0); ws.red_buf[1] = (g1x_t *)(p + o_red1);
    if (pl.pp.rounds) {
        for (int k = 0; k < 2; ++k) {
            ws.pw.refs[k] = (uint32_t *)(p + o_pr[k]); ws.pw.counts[k] = (uint32_t *)(p + o_pc[k]); ws.pw.starts[k] = (uint32_t *)(p + o_ps[k]);
        }
        ws.pw.pk = (unsigned long long *)(p + o_pk); ws.pw.scan_tmp = (unsigned long long *)(p + o_pscan);
        ws.pw.pairrefs = (uint2 *)(p + o_prefs); ws.pw.prefix = (fq_t *)(p + o_ppre); ws.pw.pool = (g1a_t *)(p + o_pool);
    }
    *max_tasks_out = max_tasks;
    *max_heavy_tasks_out = max_heavy_tasks;
    return ZKB_OK;
}

MsmState *state(zkb_ctx *ctx) {
    if (!ctx->msm_state) ctx->msm_state = new MsmState();
    return (MsmState *)ctx->msm_state;
}

// One part of a multi-part MSM (msm_run_parts): the parts are point ranges of ONE MSM that share the window plan and the
// dense bucket array; every part sorts its own entries (on the sort stream, under the previous part's accumulation) and
// accumulates INTO the buckets; only the last part runs the window reduction.
struct PartCtl {
    int part, parts;
    g1x_t *buckets;                      // the shared bucket array
    cudaEvent_t input_ready;             // this part's scalars are in place (may be null)
};

// Enqueue the whole MSM; the G group sums end up in the slot's pinned buffer.
// fb == nullptr: bases are d_points[0..n); else: bases are fb rows, ids offset by `offset`.
// pipelined == false: everything on ctx->stream.  pipelined == true: sorting + accumulation on ctx->stream, the
// latency-bound tail (window reduction, final fold, D2H) on the high-priority tail stream so that it overlaps the
// next MSM's sort and accumulation; the caller alternates slots and waits on slot.tail_done.
int msm_enqueue(zkb_ctx *ctx, const g1a_t *d_points, const uint4 *d_scalars, size_t n, int force_c, const FixedBase *fb,
                size_t offset, MsmPlan *plan_out, int slot_id = 0, bool pipelined = false, const PartCtl *pc = nullptr) {
    if (n >= (1ull << 31)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm: n must be < 2^31");
    MsmPlan pl = make_plan(n, force_c, fb, offset, ctx->sm_count, pc ? 0 : ctx->msm_mode);     // no pair rounds inside a multi-part MSM
    if ((uint64_t)n * pl.W >= (1ull << 32)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm: n * windows must be < 2^32");
    MsmWs ws;
    uint64_t max_tasks, max_heavy;
    MsmState *st = state(ctx);
    MsmSlot &sl = st->slot[slot_id];
    if (!sl.acc_done) {
        ZKB_CUDA(ctx, cudaEventCreateWithFlags(&sl.acc_done, cudaEventDisableTiming));
        ZKB_CUDA(ctx, cudaEventCreateWithFlags(&sl.tail_done, cudaEventDisableTiming));
    }
    if (pipelined && !st->tail_stream) {
        int lo_prio = 0, hi_prio = 0;
        ZKB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        ZKB_CUDA(ctx, cudaStreamCreateWithPriority(&st->tail_stream, cudaStreamNonBlocking, hi_prio));
    }
    int rc = carve_ws(ctx, sl.ws, pl, n, ws, &max_tasks, &max_heavy);
    if (rc) return rc;
    size_t out_bytes = (size_t)pl.G * (pl.red_levels + 1) * sizeof(g1x_t);      // per group: total, A_0 .. A_{L-1}
    if (sl.pinned_bytes < out_bytes) {
        if (sl.pinned) cudaFreeHost(sl.pinned);
        ZKB_CUDA(ctx, cudaMallocHost(&sl.pinned, out_bytes));
        sl.pinned_bytes = out_bytes;
    }
    cudaStream_t const main_stream = ctx->stream;
    const uint32_t nb = (uint32_t)pl.nbuckets, n32 = (uint32_t)n;
    const uint32_t gstride = pl.G == 1 ? 0 : pl.B;
    if (!st->ev[0][0]) {
        for (auto &part : st->ev) for (cudaEvent_t &e : part) ZKB_CUDA(ctx, cudaEventCreate(&e));
        for (cudaEvent_t &e : st->part_sorted) ZKB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (cudaEvent_t &e : st->part_acc) ZKB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) ZKB_CUDA(ctx, cudaEventCreate(&st->ev_pairs[k]));
    }
    const int part = pc ? pc->part : 0;
    const bool first_part = part == 0, last_part = !pc || part == pc->parts - 1;
    cudaEvent_t *ev = st->ev[part];
    if (first_part) { st->last_entries = 0; st->ev_parts = pc ? pc->parts : 1; }
    st->last_entries += (uint64_t)n * pl.W; st->last_c = pl.c; st->last_W = pl.W;
    // sort phase: on the sort stream for the parts of a multi-part MSM (helpers enqueue on ctx->stream: swapped for the phase)
    cudaStream_t s = main_stream;
    if (pc) {
        s = st->sort_stream;
        ws.bucket_val = pc->buckets;
        if (pc->input_ready) ZKB_CUDA(ctx, cudaStreamWaitEvent(s, pc->input_ready, 0));
        if (part >= 2) ZKB_CUDA(ctx, cudaStreamWaitEvent(s, st->part_acc[part - 2], 0));   // the slot's sort buffers are free again
        else if (first_part) {                                              // everything enqueued before this MSM comes first
            ZKB_CUDA(ctx, cudaEventRecord(st->part_acc[MsmState::MAX_PARTS - 1], main_stream));
            ZKB_CUDA(ctx, cudaStreamWaitEvent(s, st->part_acc[MsmState::MAX_PARTS - 1], 0));
        }
        ctx->stream = s;
    }
    struct Restore { zkb_ctx *c; cudaStream_t m; ~Restore() { c->stream = m; } } restore{ctx, main_stream};
    ZKB_CUDA(ctx, cudaEventRecord(ev[0], s));

    ZKB_CUDA(ctx, cudaMemsetAsync(ws.counts, 0, ws.zero_bytes, s));           // counts, size_hist, misc
    if (first_part) ZKB_CUDA(ctx, cudaMemsetAsync(ws.bucket_val, 0, (size_t)nb * sizeof(g1x_t), s));
    if (n32) msm_count_kernel<<<(n32 + 255) / 256, 256, 0, s>>>(d_scalars, n32, pl.c, pl.W, pl.wide, gstride, ws.counts);
    rc = exclusive_scan(ctx, ws.counts, ws.starts, nb, ws.scan_tmp, nullptr, ws.cursor);
    if (rc) return rc;
    if (n32) msm_scatter_kernel<<<(n32 + 255) / 256, 256, 0, s>>>(d_scalars, n32, pl.c, pl.W, pl.wide, gstride, pl.id_base, pl.id_stride,
                                                                  ws.cursor, ws.sorted);
    // ---- batched-affine pair rounds (msm_pairs.cuh): each halves every bucket; the XYZZ accumulation below then works on
    // what is left, through the references / counts / starts of the last round
    const g1a_t *pool = nullptr;
    ZKB_CUDA(ctx, cudaEventRecord(st->ev_pairs[0], s));
#if ZKB_HAVE_PAIR_ROUNDS
    if (pl.pp.rounds && n32) {
        const uint32_t *refs = ws.sorted, *cnts = ws.counts, *sts = ws.starts;
        const uint32_t nscan = nb + 1, ntiles = (nscan + pairs::SCAN_TILE64 - 1) / pairs::SCAN_TILE64;
        for (uint32_t r = 0; r < pl.pp.rounds; ++r) {
            uint32_t *nrefs = ws.pw.refs[r & 1], *ncnts = ws.pw.counts[r & 1], *nsts = ws.pw.starts[r & 1];
            pairs::pack_kernel<<<(nscan + 255) / 256, 256, 0, s>>>(cnts, nb, ws.pw.pk);
            pairs::scan64_reduce_kernel<<<ntiles, 256, 0, s>>>(ws.pw.pk, nscan, ws.pw.scan_tmp);
            pairs::scan64_sums_kernel<<<1, 256, 0, s>>>(ws.pw.scan_tmp, ntiles);
            pairs::scan64_apply_kernel<<<ntiles, 256, 0, s>>>(ws.pw.pk, nscan, ws.pw.scan_tmp);
            pairs::finish_kernel<<<(nb + 255) / 256, 256, 0, s>>>(cnts, sts, refs, (const uint2 *)ws.pw.pk, nb, ncnts, nsts, nrefs);
            pairs::pair_add_kernel<<<pl.pp.grid[r], pairs::THREADS, 0, s>>>(d_points, ws.pw.pool, refs, sts, (const uint2 *)ws.pw.pk, nb,
                                                                         pl.pp.m[r], (uint32_t)pl.pp.pool_base[r], nrefs, ws.pw.pairrefs,
                                                                         ws.pw.prefix);
            refs = nrefs; cnts = ncnts; sts = nsts;
        }
        ZKB_CUDA(ctx, cudaGetLastError());
        ws.sorted = const_cast<uint32_t *>(refs); ws.counts = const_cast<uint32_t *>(cnts); ws.starts = const_cast<uint32_t *>(sts);
        pool = ws.pw.pool;
        ctx->launches += 6 * pl.pp.rounds;
    }
#endif
    ZKB_CUDA(ctx, cudaEventRecord(st->ev_pairs[1], s));
    st->last_rounds = n32 ? pl.pp.rounds : 0;
    msm_ntasks_kernel<<<(nb + 255) / 256, 256, 0, s>>>(ws.counts, nb, pl.seg, ws.ntasks, ws.size_hist, ws.misc, ws.heavy_list, ws.heavy_slot,
                                                       ws.multi_list, ws.chunk_items);
    rc = exclusive_scan(ctx, ws.ntasks, ws.task_base, nb, ws.scan_tmp, ws.misc + 1);
    if (rc) return rc;
    scan_sums_kernel<<<1, 256, 0, s>>>(ws.size_hist, pl.seg + 1, nullptr);   // <= SEG_MAX + 1 bins: one CTA, in place (size_cursor == size_hist)
    msm_task_scatter_kernel<<<(nb + 255) / 256, 256, 0, s>>>(ws.counts, ws.ntasks, nb, pl.seg, ws.size_cursor, ws.task_order);
    if (pc) {                                                           // accumulation: back on the main stream, after this part's sort
        ZKB_CUDA(ctx, cudaEventRecord(st->part_sorted[part], s));
        s = main_stream;
        ctx->stream = main_stream;
        ZKB_CUDA(ctx, cudaStreamWaitEvent(s, st->part_sorted[part], 0));
    }
    const uint32_t accum = first_part ? 0u : 1u;
    ZKB_CUDA(ctx, cudaEventRecord(ev[1], s));
    msm_accumulate_kernel<<<(unsigned)((max_tasks + 127) / 128), 128, 0, s>>>(d_points, pool, ws.sorted, ws.counts, ws.starts, ws.ntasks,
                                                                             ws.task_base, ws.task_order, ws.misc, pl.seg, ws.task_out,
                                                                             ws.bucket_val, (uint32_t)ctx->msm_prefetch, accum);
    ZKB_CUDA(ctx, cudaEventRecord(ev[2], s));
    msm_combine_chunks_kernel<<<ctx->sm_count * 2, 128, 0, s>>>(ws.misc, ws.heavy_list, ws.heavy_slot, ws.chunk_items, ws.ntasks, ws.task_base,
                                                                ws.task_out, ws.chunk_out, ws.bucket_val, accum);
    msm_combine_final_kernel<<<ctx->sm_count, 128, 0, s>>>(ws.misc, ws.heavy_list, ws.heavy_slot, ws.multi_list, ws.ntasks, ws.chunk_out,
                                                          ws.bucket_val, accum);
    ZKB_CUDA(ctx, cudaEventRecord(ev[3], s));
    if (pc) ZKB_CUDA(ctx, cudaEventRecord(st->part_acc[part], s));
    if (!last_part) {                                                   // the window reduction belongs to the last part
        ZKB_CUDA(ctx, cudaEventRecord(ev[4], s));
        ZKB_CUDA(ctx, cudaGetLastError());
        ctx->launches += 14;
        *plan_out = pl;
        return ZKB_OK;
    }
    cudaStream_t ts = s;
    if (pipelined) {
        ts = st->tail_stream;
        ZKB_CUDA(ctx, cudaEventRecord(sl.acc_done, s));
        ZKB_CUDA(ctx, cudaStreamWaitEvent(ts, sl.acc_done, 0));
    }
    const g1x_t *level_in = ws.bucket_val;
    for (uint32_t l = 0; l < pl.red_levels; ++l) {
        const uint32_t m_in = pl.red_m[l], m_out = pl.red_m[l + 1];
        dim3 grid((m_out + RED_THREADS - 1) / RED_THREADS, l + 1, pl.G);
        g1x_t *level_out = ws.red_buf[l & 1];
        // binary levels: four lanes per addition (msm_wsum_level_coop_kernel; 0.423 -> 0.386 ms at 2^19 buckets,
        // profiles/r02z_msm_coop.jsonl).  A first version of that idea (every lane loading both points, fourteen products laid
        // out over the lanes without regard to their dependencies, the last levels fused into one CTA) had measured 0.50 ms
        // (profiles/r02f_msm_coop.jsonl).  Also measured and dropped: the last ten levels as ONE launch of masked tree sums over
        // the 1024 partial sums per stream, 0.429 vs 0.429 ms (profiles/r02v_msm_tail.jsonl): a level costs its one dependent
        // XYZZ addition, not its launch
        if (pl.red_r[l] != 2) msm_wsum_level_kernel<true><<<grid, RED_THREADS, 0, ts>>>(level_in, m_in, m_out, pl.red_r[l], l, level_out);
        else if (ctx->msm_coop) msm_wsum_level_coop_kernel<<<dim3((4 * m_out + RED_THREADS - 1) / RED_THREADS, l + 1, pl.G), RED_THREADS, 0, ts>>>(level_in, m_in, m_out, l, level_out);
        else msm_wsum_level_kernel<false><<<grid, RED_THREADS, 0, ts>>>(level_in, m_in, m_out, 2, l, level_out);
        level_in = level_out;
    }
    ZKB_CUDA(ctx, cudaEventRecord(ev[4], ts));
    ZKB_CUDA(ctx, cudaGetLastError());
    st->ev_valid = true;
    ctx->launches += 14 + pl.red_levels; // 2 scans x 3 kernels + the bin scan, count, scatter, ntasks, task_scatter, accumulate, heavy x 2, reduction levels
    // the last level's output is [stream][group][1]: stream 0 = plain total, stream 1 + k = total of A_k
    ZKB_CUDA(ctx, cudaMemcpyAsync(sl.pinned, pl.red_levels ? (const void *)level_in : (const void *)ws.bucket_val, out_bytes,
                                  cudaMemcpyDeviceToHost, ts));
    ZKB_CUDA(ctx, cudaEventRecord(sl.tail_done, ts));
    *plan_out = pl;
    return ZKB_OK;
}

// fold the level totals of a finished MSM on the host; result in XYZZ.
// Per group: sum_j (j+1) X[j] = total + A_0 + R_0 * (A_1 + R_1 * (A_2 + ...)); then the windows (plain bases only).
hec::Pt msm_fold(const MsmPlan &pl, const void *pinned) {
    const hec::Pt *part = (const hec::Pt *)pinned;              // [stream][group]
    auto group_sum = [&](uint32_t g) {
        hec::Pt acc = hec::inf();
        for (uint32_t l = pl.red_levels; l-- > 0;) {
            acc = hec::mul_small(acc, pl.red_r[l]);
            acc = hec::add(acc, part[(size_t)(1 + l) * pl.G + g]);
        }
        return hec::add(acc, part[g]);
    };
    if (pl.G == 1) return group_sum(0);
    hec::Pt total = hec::inf();
    for (uint32_t w = pl.W; w-- > 0;) {
        for (uint32_t k = 0; k < pl.c; ++k) total = hec::dbl(total);
        total = hec::add(total, group_sum(w));
    }
    return total;
}

// wait for the stream, fold the group sums on the host; result in XYZZ
int msm_finish(zkb_ctx *ctx, const MsmPlan &pl, hec::Pt *out, int slot_id = 0) {
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = msm_fold(pl, state(ctx)->slot[slot_id].pinned);
    return ZKB_OK;
}

// ONE MSM over resident points [offset, offset + n) cut into `parts` point ranges that share the window plan and the bucket
// array (PartCtl).  What it buys: the sort phase (digit extraction, counting sort, task ordering: ~0.32 ms at 2^20, bound by
// atomics and launch latency, not by the multiplier) of part k + 1 runs on its own stream under the accumulation of part k, and
// with host scalars the upload of part k + 1 runs under both -- while the bucket reduction, the other fixed cost of an MSM, is
// still paid once (cutting an MSM into independent MSMs pays it per range: measured slower beyond two ranges, profiles/r02g).
// scalars_host != nullptr: canonical scalars in host memory, uploaded range by range on the copy stream into ctx->stage.
int msm_run_parts(zkb_ctx *ctx, const uint64_t *scalars_dev, const uint64_t *scalars_host, size_t offset, size_t n, int parts, hec::Pt *out) {
    MsmState *st = state(ctx);
    if (parts > MsmState::MAX_PARTS - 1) parts = MsmState::MAX_PARTS - 1;
    const FixedBase *fb = (const FixedBase *)st->fixed_base;
    if (fb && (ctx->msm_force_c > 0 || fb->n != ctx->srs_n)) fb = nullptr;
    const MsmPlan whole = make_plan(n, ctx->msm_force_c, fb, offset, ctx->sm_count, 0);
    const int force_c = (int)whole.c;                               // every part uses the whole MSM's window
    if (!st->sort_stream) {
        int lo_prio = 0, hi_prio = 0;
        ZKB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        ZKB_CUDA(ctx, cudaStreamCreateWithPriority(&st->sort_stream, cudaStreamNonBlocking, hi_prio));
    }
    if (scalars_host && !st->copy_stream) ZKB_CUDA(ctx, cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t &e : st->part_uploaded) if (!e) ZKB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // all memory up front: nothing may be reallocated while the streams of the earlier parts are still running
    int rc = zkb_reserve(ctx, st->shared_buckets, whole.nbuckets * sizeof(g1x_t));
    if (rc) return rc;
    size_t lo[MsmState::MAX_PARTS + 1];
    // Equal ranges.  A smaller first range (less to wait for before the GPU starts) measured SLOWER for host scalars at 2^20:
    // 3.05 ms with four quarters, 3.14 / 3.17 / 3.23 ms with a first range of 12 / 8 / 5 % (profiles/r02am, r02an): every range
    // costs a sort of ~12 launches, and the later, larger ranges then expose more of their own sort.  ZKB_MSM_FIRST_PCT keeps
    // the knob for other hosts.
    static const int first_pct_env = [] { const char *e = getenv("ZKB_MSM_FIRST_PCT"); int v = e ? atoi(e) : 0; return v >= 1 && v <= 90 ? v : 0; }();
    const int first_pct = scalars_host && parts > 1 ? first_pct_env : 0;
    lo[0] = 0;
    const size_t first = first_pct ? n * (size_t)first_pct / 100 : n / (size_t)parts;
    for (int k = 1; k < parts; ++k) lo[k] = first + (n - first) * (size_t)(k - 1) / (size_t)(parts - 1);
    lo[parts] = n;
    for (int k = 0; k < parts; ++k) {
        MsmWs tmp;
        uint64_t a, b;
        const MsmPlan pk = make_plan(lo[k + 1] - lo[k], force_c, fb, offset + lo[k], ctx->sm_count, 0);
        rc = carve_ws(ctx, st->slot[k & 1].ws, pk, lo[k + 1] - lo[k], tmp, &a, &b);
        if (rc) return rc;
    }
    const uint4 *scal = (const uint4 *)scalars_dev;
    if (scalars_host) {
        rc = zkb_reserve(ctx, ctx->stage, n * 32 + 32);
        if (rc) return rc;
        scal = (const uint4 *)ctx->stage.p;
        // the staging buffer may still be read by work enqueued earlier on the main stream: the copies wait for it
        ZKB_CUDA(ctx, cudaEventRecord(st->part_uploaded[7], ctx->stream));
        ZKB_CUDA(ctx, cudaStreamWaitEvent(st->copy_stream, st->part_uploaded[7], 0));
    }
    MsmPlan plan;
    for (int k = 0; k < parts; ++k) {
        const size_t cnt = lo[k + 1] - lo[k];
        if (scalars_host) {
            ZKB_CUDA(ctx, cudaMemcpyAsync((char *)ctx->stage.p + lo[k] * 32, scalars_host + 4 * lo[k], cnt * 32, cudaMemcpyHostToDevice, st->copy_stream));
            ZKB_CUDA(ctx, cudaEventRecord(st->part_uploaded[k], st->copy_stream));
        }
        PartCtl pc{k, parts, (g1x_t *)st->shared_buckets.p, scalars_host ? st->part_uploaded[k] : nullptr};
        const size_t off = offset + lo[k];
        rc = fb ? msm_enqueue(ctx, (const g1a_t *)fb->rows.p, scal + 2 * lo[k], cnt, 0, fb, off, &plan, k & 1, false, &pc)
                : msm_enqueue(ctx, (const g1a_t *)ctx->srs.p + off, scal + 2 * lo[k], cnt, force_c, nullptr, 0, &plan, k & 1, false, &pc);
        if (rc) { cudaStreamSynchronize(st->sort_stream); cudaStreamSynchronize(ctx->stream); return rc; }
    }
    return msm_finish(ctx, plan, out, (parts - 1) & 1);
}

}  // namespace

void zkb_msm_release(zkb_ctx *ctx) {
    MsmState *st = (MsmState *)ctx->msm_state;
    if (!st) return;
    for (MsmSlot &sl : st->slot) {
        if (sl.pinned) cudaFreeHost(sl.pinned);
        if (sl.ws.p) cudaFree(sl.ws.p);
        if (sl.acc_done) cudaEventDestroy(sl.acc_done);
        if (sl.tail_done) cudaEventDestroy(sl.tail_done);
    }
    if (st->tail_stream) cudaStreamDestroy(st->tail_stream);
    if (st->copy_stream) cudaStreamDestroy(st->copy_stream);
    for (cudaEvent_t e : st->part_uploaded) if (e) cudaEventDestroy(e);
    for (auto &part : st->ev) for (cudaEvent_t e : part) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : st->part_sorted) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : st->part_acc) if (e) cudaEventDestroy(e);
    if (st->sort_stream) cudaStreamDestroy(st->sort_stream);
    if (st->shared_buckets.p) cudaFree(st->shared_buckets.p);
    for (int k = 0; k < 2; ++k) if (st->ev_pairs[k]) cudaEventDestroy(st->ev_pairs[k]);
    if (st->fixed_base) {
        FixedBase *fb = (FixedBase *)st->fixed_base;
        if (fb->rows.p) cudaFree(fb->rows.p);
        delete fb;
    }
    delete st;
    ctx->msm_state = nullptr;
}

void zkb_commit_abort(zkb_ctx *ctx) {
    MsmState *st = (MsmState *)ctx->msm_state;
    if (!st || st->pipe_partial.empty()) return;
    for (size_t k = 0; k < st->pipe_pending.size(); ++k)
        if (st->pipe_pending[k]) cudaEventSynchronize(st->slot[k & 1].tail_done);
    st->pipe_partial.clear();
    st->pipe_pending.clear();
    st->pipe_expected = 1;
}

extern "C" {

int zkb_srs_precompute(zkb_ctx *ctx, int c);

int zkb_srs_load_g1(zkb_ctx *ctx, const uint64_t *xy_mont_host, size_t n) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!xy_mont_host && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_load_g1: null points");
    int rc = zkb_reserve(ctx, ctx->srs, n * AFF_BYTES + AFF_BYTES);
    if (rc) return rc;
    ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->srs.p, xy_mont_host, n * AFF_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->srs_n = n;
    ctx->srs_lo = 0;
    ctx->srs_global_n = n;
    ctx->srs_replicated = false;
    zkb_srs_precompute(ctx, -1);
    return ZKB_OK;
}

int zkb_srs_load_g1_dev(zkb_ctx *ctx, const uint64_t *xy_mont_dev, size_t n) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!xy_mont_dev && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_load_g1_dev: null points");
    int rc = zkb_reserve(ctx, ctx->srs, n * AFF_BYTES + AFF_BYTES);
    if (rc) return rc;
    ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->srs.p, xy_mont_dev, n * AFF_BYTES, cudaMemcpyDeviceToDevice, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->srs_n = n;
    ctx->srs_lo = 0;
    ctx->srs_global_n = n;
    ctx->srs_replicated = false;
    zkb_srs_precompute(ctx, -1);
    return ZKB_OK;
}

// size of the committer key (the whole key when this context holds one range of it: zkb_srs_set_range)
size_t zkb_srs_size(zkb_ctx *ctx) { return ctx ? ctx->srs_global_n : 0; }

// Build (c > 0: with that window size, c == 0: cost model) or drop (c < 0) the fixed-base tables of the resident SRS.
int zkb_srs_precompute(zkb_ctx *ctx, int c) {
    if (!ctx) return ZKB_ERR_INVALID;
    MsmState *st = state(ctx);
    if (st->fixed_base) {
        FixedBase *old = (FixedBase *)st->fixed_base;
        ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (old->rows.p) cudaFree(old->rows.p);
        delete old;
        st->fixed_base = nullptr;
    }
    if (c < 0) return ZKB_OK;
    if ((c > 0 && c < 3) || c > 24) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_precompute: window size out of range (3..24)");
    const size_t n = ctx->srs_n;
    if (n == 0) ZKB_FAIL(ctx, ZKB_ERR_NO_SRS, "zkb_srs_precompute: no SRS loaded");
    FixedBase *fb = new FixedBase();
    // replicated key on several GPUs: an MSM covers a share of the key (1/world of it when every commitment is sharded, about
    // 3/world in a fanned-out round of three), so the window is chosen for that size, not for the whole key
    const size_t n_typical = (ctx->srs_replicated && ctx->world > 1) ? std::max<size_t>(n * 2 / (size_t)ctx->world, 1024) : n;
    fb->c = c > 0 ? (uint32_t)c : pick_window(n_typical, true, ctx->sm_count);
    fb->W = balanced_windows(fb->c, &fb->wide);
    if (fb->wide == 0 && fb->c > 3) { fb->c -= 1; fb->wide = fb->W; }   // every window narrow: that is just c - 1 (half the buckets)
    fb->n = n;
    if ((uint64_t)n * fb->W >= (1ull << 31)) { delete fb; ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_precompute: n * windows must be < 2^31"); }
    int rc = zkb_reserve(ctx, fb->rows, (size_t)fb->W * n * sizeof(g1a_t));
    if (rc) { delete fb; return rc; }
    msm_precompute_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const g1a_t *)ctx->srs.p, (uint32_t)n, fb->c, fb->W,
                                                                                fb->wide, (g1a_t *)fb->rows.p);
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(fb->rows.p); delete fb; ctx->err = cudaGetErrorString(e); return ZKB_ERR_CUDA; }
    st->fixed_base = fb;
    return ZKB_OK;
}

int zkb_msm_set_mode(zkb_ctx *ctx, int mode) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (mode < -1 || mode > pairs::MAX_ROUNDS) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_set_mode: -1 = automatic, 0 = XYZZ accumulation only, 1..6 = that many batched-affine pair rounds first");
    ctx->msm_mode = mode;
    return ZKB_OK;
}

// A large single MSM runs as point ranges through shared buckets (msm_run_parts): `dev_parts` ranges when the scalars are in
// HBM, `host_parts` when they come from host memory (1 = the whole MSM at once), for MSMs of at least 2^min_log points.
// Defaults: 1 / 4 / 19 (ZKB_MSM_PARTS="d,h,min_log" in the environment overrides them at context creation).
int zkb_msm_set_parts(zkb_ctx *ctx, int dev_parts, int host_parts, int min_log) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (dev_parts < 1 || dev_parts > 7 || host_parts < 1 || host_parts > 7 || min_log < 4 || min_log > 31)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_set_parts: parts in 1..7, min_log in 4..31");
    ctx->msm_parts_dev = dev_parts;
    ctx->msm_parts_host = host_parts;
    ctx->msm_parts_min_log = min_log;
    return ZKB_OK;
}

int zkb_msm_set_window(zkb_ctx *ctx, int c) {
    if (!ctx || c < 0 || c > 24 || c == 1) return ZKB_ERR_INVALID;
    state(ctx);
    ctx->msm_force_c = c;
    return ZKB_OK;
}

// scalars on the device, bases = resident SRS[offset .. offset + n); result as an XYZZ partial sum (16 limbs)
int zkb_msm_g1_dev_partial(zkb_ctx *ctx, const uint64_t *scalars_dev, size_t offset, size_t n, uint64_t *out_xyzz /* XYZZ_W words */) {
    if (!ctx || !out_xyzz) return ZKB_ERR_INVALID;
    if (offset + n > ctx->srs_n) ZKB_FAIL(ctx, ZKB_ERR_NO_SRS, "zkb_msm: offset + n exceeds the loaded SRS");
    if (!scalars_dev && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm: null scalars");
    if (!state(ctx)->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "a zkb_commit_push batch is open (its scalars and results live in the buffers this call would reuse): call zkb_commit_finish first");
    MsmPlan pl;
    if (n >= ((size_t)1 << ctx->msm_parts_min_log) && ctx->msm_parts_dev > 1 && ctx->msm_mode == 0) {   // large: point ranges through shared buckets
        hec::Pt total;
        int rc = msm_run_parts(ctx, scalars_dev, nullptr, offset, n, ctx->msm_parts_dev, &total);
        if (rc) return rc;
        memcpy(out_xyzz, &total, XYZZ_BYTES);
        return ZKB_OK;
    }
    const FixedBase *fb = (const FixedBase *)state(ctx)->fixed_base;
    if (fb && (ctx->msm_force_c > 0 || fb->n != ctx->srs_n)) fb = nullptr;      // forced window: plain path
    int rc = fb ? msm_enqueue(ctx, (const g1a_t *)fb->rows.p, (const uint4 *)scalars_dev, n, 0, fb, offset, &pl)
                : msm_enqueue(ctx, (const g1a_t *)ctx->srs.p + offset, (const uint4 *)scalars_dev, n, ctx->msm_force_c, nullptr, 0, &pl);
    if (rc) return rc;
    hec::Pt total;
    rc = msm_finish(ctx, pl, &total);
    if (rc) return rc;
    memcpy(out_xyzz, &total, XYZZ_BYTES);
    return ZKB_OK;
}

int zkb_msm_g1_dev(zkb_ctx *ctx, const uint64_t *scalars_dev, size_t offset, size_t n, uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    uint64_t xyzz[XYZZ_W];
    int rc = zkb_msm_g1_dev_partial(ctx, scalars_dev, offset, n, xyzz);
    if (rc) return rc;
    hec::Pt p;
    memcpy(&p, xyzz, XYZZ_BYTES);
    hec::to_affine(p, out_xy, is_inf);
    return ZKB_OK;
}

// Point-range sharded MSM: every rank of the communicator calls this with the scalars of ITS resident range
// (scalars_dev[i] pairs with resident point offset + i); the ranks' partial sums are exchanged over NCCL (csrc/comm.cu) and
// every rank returns the same affine point.  world == 1: the same as zkb_msm_g1_dev.
int zkb_msm_g1_sharded_dev(zkb_ctx *ctx, const uint64_t *scalars_dev, size_t offset, size_t n, uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    if (ctx->world == 1) return zkb_msm_g1_dev(ctx, scalars_dev, offset, n, out_xy, is_inf);
    uint64_t mine[XYZZ_W];
    int rc = zkb_msm_g1_dev_partial(ctx, scalars_dev, offset, n, mine);
    if (rc) return rc;
    std::vector<uint64_t> all((size_t)ctx->world * XYZZ_W);
    rc = zkb_comm_allgather(ctx, mine, XYZZ_BYTES, all.data());
    if (rc) return rc;
    return zkb_g1_sum_partials(all.data(), (size_t)ctx->world, out_xy, is_inf);
}

// Host scalars.  From 2^18 points on the MSM runs as two or three point ranges through the two pipelined workspaces: the
// next range of the scalars crosses PCIe (copy stream) while the previous one is sorted and accumulated, and a range's
// window reduction overlaps the next range's accumulation; the partial sums are added on the host.
static int msm_host_partial(zkb_ctx *ctx, const uint64_t *scalars_host, size_t offset, size_t n, uint64_t *out_xyzz /* XYZZ_W words */) {
    if (!scalars_host && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_g1: null scalars");
    if (offset + n > ctx->srs_n) ZKB_FAIL(ctx, ZKB_ERR_NO_SRS, "zkb_msm: offset + n exceeds the loaded SRS");
    if (!state(ctx)->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "a zkb_commit_push batch is open (its scalars and results live in the buffers this call would reuse): call zkb_commit_finish first");
    int rc = zkb_reserve(ctx, ctx->stage, n * 32 + 32);
    if (rc) return rc;
    MsmState *st = state(ctx);
    if (n < ((size_t)1 << (ctx->msm_parts_min_log > 18 ? 18 : ctx->msm_parts_min_log))) {
        ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->stage.p, scalars_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        return zkb_msm_g1_dev_partial(ctx, (const uint64_t *)ctx->stage.p, offset, n, out_xyzz);
    }
    if (ctx->msm_parts_host > 1 && ctx->msm_mode == 0) {           // ranges through shared buckets, uploads under the earlier ranges' work
        hec::Pt total;
        rc = msm_run_parts(ctx, nullptr, scalars_host, offset, n, ctx->msm_parts_host, &total);
        if (rc) return rc;
        memcpy(out_xyzz, &total, XYZZ_BYTES);
        return ZKB_OK;
    }
    // zkb_msm_set_parts(ctx, d, 1, ..): the earlier scheme -- two independent MSMs over two ranges (each with its own buckets and reduction)
    if (!st->copy_stream) ZKB_CUDA(ctx, cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t &e : st->part_uploaded) if (!e) ZKB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const FixedBase *fb = (const FixedBase *)st->fixed_base;
    if (fb && (ctx->msm_force_c > 0 || fb->n != ctx->srs_n)) fb = nullptr;
    // the staging buffer may still be read by work enqueued earlier on the main stream: the copies wait for it
    ZKB_CUDA(ctx, cudaEventRecord(st->part_uploaded[0], ctx->stream));
    ZKB_CUDA(ctx, cudaStreamWaitEvent(st->copy_stream, st->part_uploaded[0], 0));
    // Two point ranges.  The GPU idles until the first range has crossed PCIe, so that one is the smaller; every range costs
    // a sort and a window reduction of its own (the first one's reduction runs on the tail stream under the second one's
    // accumulation).  Measured at 2^20 (profiles/r02g, r02h): three ranges of 20 / 40 / 40 % are SLOWER than two halves (3.71
    // vs 3.41 ms: the call is bound by GPU work, not by the transfer); a 30 / 70 split measured 3.37 ms against 3.44 for halves.
    static const int first_pct = [] { const char *e = getenv("ZKB_MSM_SPLIT"); int v = e ? atoi(e) : 0; return v >= 10 && v <= 90 ? v : 30; }();
    const int parts = 2;
    size_t part_lo[4] = {0, n * (size_t)first_pct / 100, n, n};
    MsmPlan plans[2];
    hec::Pt total = hec::inf();
    for (int k = 0; k < parts; ++k) {
        const int slot = k & 1;
        const size_t lo = part_lo[k], cnt = part_lo[k + 1] - lo;
        if (k >= 2) {                                             // the slot's previous range: fold its result, free its workspace
            ZKB_CUDA(ctx, cudaEventSynchronize(st->slot[slot].tail_done));
            total = hec::add(total, msm_fold(plans[slot], st->slot[slot].pinned));
        }
        char *dst = (char *)ctx->stage.p + lo * 32;
        ZKB_CUDA(ctx, cudaMemcpyAsync(dst, scalars_host + 4 * lo, cnt * 32, cudaMemcpyHostToDevice, st->copy_stream));
        ZKB_CUDA(ctx, cudaEventRecord(st->part_uploaded[k], st->copy_stream));
        ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, st->part_uploaded[k], 0));
        const size_t off = offset + lo;
        rc = fb ? msm_enqueue(ctx, (const g1a_t *)fb->rows.p, (const uint4 *)dst, cnt, 0, fb, off, &plans[slot], slot, true)
                : msm_enqueue(ctx, (const g1a_t *)ctx->srs.p + off, (const uint4 *)dst, cnt, ctx->msm_force_c, nullptr, 0, &plans[slot], slot, true);
        if (rc) return rc;
    }
    for (int k = parts >= 2 ? parts - 2 : 0; k < parts; ++k) {
        ZKB_CUDA(ctx, cudaEventSynchronize(st->slot[k & 1].tail_done));
        total = hec::add(total, msm_fold(plans[k & 1], st->slot[k & 1].pinned));
    }
    memcpy(out_xyzz, &total, XYZZ_BYTES);
    return ZKB_OK;
}

int zkb_msm_g1(zkb_ctx *ctx, const uint64_t *scalars_host, size_t offset, size_t n, uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    uint64_t xyzz[XYZZ_W];
    int rc = msm_host_partial(ctx, scalars_host, offset, n, xyzz);
    if (rc) return rc;
    return zkb_g1_sum_partials(xyzz, 1, out_xy, is_inf);
}

// zkb_msm_g1_sharded_dev with HOST scalars: every rank uploads the scalars of its resident range through the pipelined
// two-range path above (upload under accumulation), then the partial sums are exchanged
int zkb_msm_g1_sharded(zkb_ctx *ctx, const uint64_t *scalars_host, size_t offset, size_t n, uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    uint64_t mine[XYZZ_W];
    int rc = msm_host_partial(ctx, scalars_host, offset, n, mine);
    if (rc) return rc;
    if (ctx->world == 1) return zkb_g1_sum_partials(mine, 1, out_xy, is_inf);
    std::vector<uint64_t> all((size_t)ctx->world * XYZZ_W);
    rc = zkb_comm_allgather(ctx, mine, XYZZ_BYTES, all.data());
    if (rc) return rc;
    return zkb_g1_sum_partials(all.data(), (size_t)ctx->world, out_xy, is_inf);
}

// arbitrary bases (drop-in for VariableBaseMSM::multi_scalar_mul / HomomorphicCommitment::multi_scalar_mul)
int zkb_msm_g1_bases(zkb_ctx *ctx, const uint64_t *points_host, const uint64_t *scalars_host, size_t n,
                     uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    if ((!points_host || !scalars_host) && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_g1_bases: null input");
    if (!state(ctx)->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "a zkb_commit_push batch is open (its scalars and results live in the buffers this call would reuse): call zkb_commit_finish first");
    int rc = zkb_reserve(ctx, ctx->stage, (n + 1) * (AFF_BYTES + 32));
    if (rc) return rc;
    char *d_pts = (char *)ctx->stage.p, *d_sc = d_pts + n * AFF_BYTES;
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_pts, points_host, n * AFF_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_sc, scalars_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    MsmPlan pl;
    rc = msm_enqueue(ctx, (const g1a_t *)d_pts, (const uint4 *)d_sc, n, ctx->msm_force_c, nullptr, 0, &pl);
    if (rc) return rc;
    hec::Pt total;
    rc = msm_finish(ctx, pl, &total);
    if (rc) return rc;
    hec::to_affine(total, out_xy, is_inf);
    return ZKB_OK;
}

// MSM over points the CALLER holds in HBM (any key: the folded committer keys of the inner-product argument's rounds, ipa.cu):
// scalars on the device too, canonical integers or (scalars_mont != 0) Montgomery form as polynomial coefficients are kept.
int zkb_msm_g1_points_dev(zkb_ctx *ctx, const uint64_t *points_dev, const uint64_t *scalars_dev, size_t n, int scalars_mont,
                          uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    return zkb_msm_points_plus(ctx, points_dev, scalars_dev, n, scalars_mont, nullptr, nullptr, out_xy, is_inf);
}

// The same plus one term on the host: out = <scalars, points> + extra_scalar * extra (affine Montgomery point and canonical
// scalar in host memory; either NULL: no extra term).  The inner-product argument's L and R are such sums (ipa.cu): the extra
// multiplication is ~380 host group operations, cheaper than a second trip through the bucket pipeline for one point.
// Library-internal (declared in ctx.h), not part of the C ABI.
extern "C++" int zkb_msm_points_plus(zkb_ctx *ctx, const uint64_t *points_dev, const uint64_t *scalars_dev, size_t n, int scalars_mont,
                                     const uint64_t *extra_xy, const uint64_t *extra_scalar_canon, uint64_t *out_xy, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    if ((!points_dev || !scalars_dev) && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_g1_points_dev: null input");
    if (!state(ctx)->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "a zkb_commit_push batch is open (its scalars and results live in the buffers this call would reuse): call zkb_commit_finish first");
    const uint4 *sc = (const uint4 *)scalars_dev;
    if (scalars_mont && n) {
        int rc = zkb_reserve(ctx, ctx->stage, n * 32 + 32);
        if (rc) return rc;
        fr_from_mont_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sc, (uint4 *)ctx->stage.p, (uint32_t)n);
        ctx->launches += 1;
        ZKB_CUDA(ctx, cudaGetLastError());
        sc = (const uint4 *)ctx->stage.p;
    }
    MsmPlan pl;
    int rc = msm_enqueue(ctx, (const g1a_t *)points_dev, sc, n, ctx->msm_force_c, nullptr, 0, &pl);
    if (rc) return rc;
    hec::Pt extra = hec::inf();
    if (extra_xy && extra_scalar_canon) {                          // on the host while the GPU runs the MSM
        bool zero = true;
        for (int i = 0; i < AFF_W; ++i) zero = zero && extra_xy[i] == 0;
        if (!zero) {
            hec::Pt e;
            memcpy(e.x.l, extra_xy, AFF_BYTES / 2);
            memcpy(e.y.l, extra_xy + AFF_W / 2, AFF_BYTES / 2);
            e.zz = host::one(host::FQ);
            e.zzz = e.zz;
            extra = hec::mul_scalar(e, extra_scalar_canon);
        }
    }
    hec::Pt total;
    rc = msm_finish(ctx, pl, &total);
    if (rc) return rc;
    hec::to_affine(hec::add(total, extra), out_xy, is_inf);
    return ZKB_OK;
}

// sum of `count` XYZZ partial results (multi-GPU combine after the all-gather), affine out
int zkb_g1_sum_partials(const uint64_t *xyzz, size_t count, uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if ((!xyzz && count) || !out_xy) return ZKB_ERR_INVALID;
    hec::Pt total = hec::inf();
    for (size_t i = 0; i < count; ++i) {
        hec::Pt p;
        memcpy(&p, xyzz + XYZZ_W * i, XYZZ_BYTES);
        total = hec::add(total, p);
    }
    hec::to_affine(total, out_xy, is_inf);
    return ZKB_OK;
}

// kzg10::commit's inner product for one polynomial held in HBM in Montgomery form:
// converts the coefficients to canonical integers (into_repr) and runs the MSM against SRS[offset..].
int zkb_commit_dev(zkb_ctx *ctx, const uint64_t *coeffs_mont_dev, size_t offset, size_t n, uint64_t *out_xy /* AFF_W words */, int *is_inf) {
    if (!ctx || !out_xy) return ZKB_ERR_INVALID;
    if (!coeffs_mont_dev && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_dev: null coefficients");
    if (ctx->world > 1 || ctx->srs_global_n != ctx->srs_n) return zkb_commit_batch_dev(ctx, &coeffs_mont_dev, &offset, &n, 1, out_xy, is_inf);
    if (offset + n > ctx->srs_n) ZKB_FAIL(ctx, ZKB_ERR_NO_SRS, "zkb_commit_dev: offset + n exceeds the loaded SRS");
    if (!state(ctx)->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "a zkb_commit_push batch is open (its scalars and results live in the buffers this call would reuse): call zkb_commit_finish first");
    int rc = zkb_reserve(ctx, ctx->stage, n * 32 + 32);
    if (rc) return rc;
    if (n) fr_from_mont_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint4 *)coeffs_mont_dev, (uint4 *)ctx->stage.p, (uint32_t)n);
    ZKB_CUDA(ctx, cudaGetLastError());
    return zkb_msm_g1_dev(ctx, (const uint64_t *)ctx->stage.p, offset, n, out_xy, is_inf);
}

// out_points_dev[i] = scalars_dev[i] * base  (affine, Montgomery); scalars canonical.
int zkb_g1_fixed_base_mul_dev(zkb_ctx *ctx, const uint64_t *base_xy /* AFF_W words */, const uint64_t *scalars_dev, size_t n, uint64_t *out_points_dev) {
    if (!ctx || !base_xy) return ZKB_ERR_INVALID;
    if ((!scalars_dev || !out_points_dev) && n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_g1_fixed_base_mul_dev: null buffer");
    g1a_t b;
    memcpy(&b, base_xy, AFF_BYTES);
    if (n) g1_fixed_base_mul_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(b, (const uint4 *)scalars_dev, (uint32_t)n, (g1a_t *)out_points_dev);
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}

// Device time of the phases of the last MSM on this context, in ms (call after the MSM returned):
// out[0] digit extraction + counting sort + task ordering, out[1] bucket accumulation (the IMAD-bound kernel),
// out[2] oversized-bucket combine, out[3] window reduction, out[4] total; info[0] = n * windows (upper bound of
// bucket insertions), info[1] = c, info[2] = windows.
int zkb_msm_last_timing(zkb_ctx *ctx, float out_ms[5], uint64_t info[3]) {
    if (!ctx || !out_ms) return ZKB_ERR_INVALID;
    MsmState *st = state(ctx);
    if (!st->ev_valid) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_last_timing: no MSM has run on this context");
    // a multi-part MSM: the phases summed over the parts (their sort phases overlap the previous part's accumulation, so the
    // sum of the phases can exceed the total); the total runs from the first part's start to the end of the reduction
    const int P = st->ev_parts;
    ZKB_CUDA(ctx, cudaEventSynchronize(st->ev[P - 1][4]));
    for (int k = 0; k < 4; ++k) out_ms[k] = 0.f;
    for (int p = 0; p < P; ++p)
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            if (k == 3 && p != P - 1) continue;
            ZKB_CUDA(ctx, cudaEventElapsedTime(&ms, st->ev[p][k], st->ev[p][k + 1]));
            out_ms[k] += ms;
        }
    ZKB_CUDA(ctx, cudaEventElapsedTime(&out_ms[4], st->ev[0][0], st->ev[P - 1][4]));
    if (info) { info[0] = st->last_entries; info[1] = st->last_c; info[2] = st->last_W; }
    return ZKB_OK;
}

// the share of phase 0 above that the batched-affine pair rounds took, in ms, and how many rounds ran
int zkb_msm_last_pair_rounds(zkb_ctx *ctx, float *pairs_ms, int *rounds) {
    if (!ctx || !pairs_ms || !rounds) return ZKB_ERR_INVALID;
    MsmState *st = state(ctx);
    if (!st->ev_valid) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_msm_last_pair_rounds: no MSM has run on this context");
    ZKB_CUDA(ctx, cudaEventSynchronize(st->ev_pairs[1]));
    ZKB_CUDA(ctx, cudaEventElapsedTime(pairs_ms, st->ev_pairs[0], st->ev_pairs[1]));
    *rounds = (int)st->last_rounds;
    return ZKB_OK;
}

// ---- pipelined commitments: zkb_commit_push enqueues the MSM of one more HBM-resident polynomial, zkb_commit_finish waits for
// the open batch and returns the affine commitments.  The MSMs alternate between two workspaces: while MSM k's
// latency-bound window reduction and result download run on a high-priority stream, MSM k+1 already sorts and
// accumulates on the main stream; between pushes the caller is free to enqueue other work or block in a host copy
// (the prover uploads wire b while the GPU commits to wire a).
// Point-range view: offsets / lengths are global; this context holds SRS[srs_lo, srs_lo + srs_n) and works on the
// overlap (the whole polynomial on a single GPU).  With world > 1 the XYZZ partial sums of all ranks are all-gathered
// (128 B per rank and commitment, one NCCL call per batch) and every rank adds them identically.
static int commit_collect(zkb_ctx *ctx, MsmState *st, size_t k) {          // fold MSM k of the open batch on the host
    if (!st->pipe_pending[k]) return ZKB_OK;
    const int slot = (int)(k & 1);
    ZKB_CUDA(ctx, cudaEventSynchronize(st->slot[slot].tail_done));
    hec::Pt p = msm_fold(st->pipe_plan[slot], st->slot[slot].pinned);
    memcpy(st->pipe_partial[k].data(), &p, XYZZ_BYTES);
    st->pipe_pending[k] = 0;
    return ZKB_OK;
}

// ---- replicated committer key (zkb_srs_set_replicated): which part of commitment k of a batch of E this rank computes.
// "Shard": every commitment is cut into `world` point ranges -- every rank runs E small MSMs and pays E times the fixed cost
// of one (sort, bucket reduction over 2^(c-1) buckets, latency-bound tail: ~0.45 ms whatever the size).  "Fan-out": the
// ranks are split into E groups, group j computes commitment j alone (sharded by point range inside the group) -- every
// rank runs ONE larger MSM.  Fan-out wins while the saved fixed costs outweigh the coarser balance (groups of unequal
// size): E (F + P len / world)  vs  F + P len / (smallest group), F and P from measurements on B200 (profiles/r02*).
static bool fanout_wins(const zkb_ctx *ctx, size_t E, size_t len) {
    const size_t world = (size_t)ctx->world;
    if (E < 2 || world < E) return false;
    if (ctx->fanout >= 0) return ctx->fanout != 0;
    const double F = 0.45, P = 2.1e-6;                          // ms per MSM; ms per point (0.14 ns x ~15 windows)
    const double shard = (double)E * (F + P * (double)len / (double)world);
    const double fan = F + P * (double)len / (double)(world / E);
    return fan < shard;
}

static void replicated_share(const zkb_ctx *ctx, size_t E, size_t k, size_t offset, size_t len, size_t *lo, size_t *hi) {
    size_t parts = (size_t)ctx->world, idx = (size_t)ctx->rank;
    if (k < E && fanout_wins(ctx, E, len)) {
        const size_t base = parts / E, rem = parts % E;          // commitment j gets base + (j < rem) consecutive ranks
        size_t start = 0, mine = E, gsize = 0;
        for (size_t j = 0; j < E; ++j) {
            const size_t sz = base + (j < rem ? 1 : 0);
            if (idx >= start && idx < start + sz) { mine = j; gsize = sz; break; }
            start += sz;
        }
        if (mine != k) { *lo = *hi = offset; return; }           // another group's commitment
        parts = gsize;
        idx -= start;
    }
    const size_t q = len / parts, r = len % parts;                // balanced contiguous ranges
    *lo = offset + idx * q + std::min(idx, r);
    *hi = *lo + q + (idx < r ? 1 : 0);
}

// host-only test hook: the share of commitment k (of a batch of E, length len at `offset`) that `rank` of `world` computes
int zkb_test_replicated_share(int world, int rank, int fanout, size_t E, size_t k, size_t offset, size_t len, size_t out_lo_hi[2]) {
    if (world < 1 || rank < 0 || rank >= world || !out_lo_hi) return ZKB_ERR_INVALID;
    zkb_ctx fake;
    fake.world = world;
    fake.rank = rank;
    fake.fanout = fanout;
    replicated_share(&fake, E, k, offset, len, &out_lo_hi[0], &out_lo_hi[1]);
    return ZKB_OK;
}

int zkb_commit_expect(zkb_ctx *ctx, size_t count) {
    if (!ctx) return ZKB_ERR_INVALID;
    MsmState *st = state(ctx);
    if (!st->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_expect: a zkb_commit_push batch is already open");
    st->pipe_expected = count ? count : 1;
    return ZKB_OK;
}

int zkb_commit_push(zkb_ctx *ctx, const uint64_t *coeffs_mont_dev, size_t offset, size_t len) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!coeffs_mont_dev && len) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_push: null coefficients");
    if (offset + len > ctx->srs_global_n) ZKB_FAIL(ctx, ZKB_ERR_NO_SRS, "zkb_commit_push: offset + n exceeds the loaded SRS");
    MsmState *st = state(ctx);
    const FixedBase *fb = (const FixedBase *)st->fixed_base;
    if (fb && (ctx->msm_force_c > 0 || fb->n != ctx->srs_n)) fb = nullptr;
    const size_t k = st->pipe_partial.size();
    const int slot = (int)(k & 1);
    size_t g_lo = std::max(offset, ctx->srs_lo), g_hi = std::min(offset + len, ctx->srs_lo + ctx->srs_n);
    if (ctx->world > 1 && ctx->srs_replicated) replicated_share(ctx, st->pipe_expected, k, offset, len, &g_lo, &g_hi);
    const size_t n = g_hi > g_lo ? g_hi - g_lo : 0;
    int rc = ZKB_OK;
    if (k >= 2) {                                                          // slot reuse: MSM k-2 must be folded first
        rc = commit_collect(ctx, st, k - 2);
        if (rc) return rc;
    }
    st->pipe_partial.emplace_back();
    st->pipe_partial[k].fill(0);                                           // identity
    st->pipe_pending.push_back(0);
    if (!n) return ZKB_OK;
    const size_t stride = ctx->srs_n * 32 + 32;                             // canonical scalars, one buffer per slot
    if (ctx->stage.bytes < 2 * stride) {
        if (k) {                                                           // growing would free memory that MSMs in flight still read
            for (size_t j = 0; j < k; ++j) if ((rc = commit_collect(ctx, st, j))) return rc;
        }
        rc = zkb_reserve(ctx, ctx->stage, 2 * stride);
        if (rc) return rc;
    }
    uint4 *scal = (uint4 *)((char *)ctx->stage.p + (size_t)slot * stride);
    // slot reuse: the scalars of the slot's previous MSM must have been consumed (the events exist once the slot has run an MSM:
    // with a fanned-out batch a rank skips the commitments of the other groups)
    if (k >= 2 && st->slot[slot].tail_done) ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, st->slot[slot].tail_done, 0));
    fr_from_mont_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint4 *)(coeffs_mont_dev + 4 * (g_lo - offset)), scal, (uint32_t)n);
    const size_t off = g_lo - ctx->srs_lo;
    rc = fb ? msm_enqueue(ctx, (const g1a_t *)fb->rows.p, scal, n, 0, fb, off, &st->pipe_plan[slot], slot, true)
            : msm_enqueue(ctx, (const g1a_t *)ctx->srs.p + off, scal, n, ctx->msm_force_c, nullptr, 0, &st->pipe_plan[slot], slot, true);
    if (rc) return rc;
    st->pipe_pending[k] = 1;
    return ZKB_OK;
}

// the open batch's results as THIS rank's un-normalised XYZZ partial sums (16 limbs each), without any exchange: for callers
// that combine the ranks themselves, and for the single-GPU tests of the multi-GPU layouts
int zkb_commit_finish_partials(zkb_ctx *ctx, uint64_t *out_xyzz /* count x 16 */) {
    if (!ctx) return ZKB_ERR_INVALID;
    MsmState *st = state(ctx);
    const size_t count = st->pipe_partial.size();
    if (!out_xyzz && count) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_finish_partials: null output");
    int rc = ZKB_OK;
    for (size_t k = 0; k < count && !rc; ++k) rc = commit_collect(ctx, st, k);
    for (size_t k = 0; k < count && !rc; ++k) memcpy(out_xyzz + XYZZ_W * k, st->pipe_partial[k].data(), XYZZ_BYTES);
    st->pipe_partial.clear();
    st->pipe_pending.clear();
    st->pipe_expected = 1;
    return rc;
}

// test hook: make this context believe it is `rank` of `world` without a communicator (only zkb_commit_push /
// zkb_commit_finish_partials are meaningful then)
int zkb_test_set_rank_world(zkb_ctx *ctx, int rank, int world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return ZKB_ERR_INVALID;
    ctx->rank = rank;
    ctx->world = world;
    return ZKB_OK;
}

int zkb_commit_finish(zkb_ctx *ctx, uint64_t *out_xy /* count x 8 */, int *is_inf /* count or NULL */) {
    if (!ctx) return ZKB_ERR_INVALID;
    MsmState *st = state(ctx);
    const size_t count = st->pipe_partial.size();
    if (!out_xy && count) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_finish: null output");
    int rc = ZKB_OK;
    for (size_t k = 0; k < count && !rc; ++k) rc = commit_collect(ctx, st, k);
    std::vector<std::array<uint64_t, XYZZ_W>> part;
    part.swap(st->pipe_partial);
    st->pipe_pending.clear();
    st->pipe_expected = 1;
    if (rc) return rc;
    if (ctx->world > 1 && count) {
        std::vector<std::array<uint64_t, XYZZ_W>> all((size_t)ctx->world * count);
        rc = zkb_comm_allgather(ctx, part.data(), count * XYZZ_BYTES, all.data());
        if (rc) return rc;
        for (size_t k = 0; k < count; ++k) {
            hec::Pt total = hec::inf();
            for (int r = 0; r < ctx->world; ++r) {
                hec::Pt p;
                memcpy(&p, all[(size_t)r * count + k].data(), XYZZ_BYTES);
                total = hec::add(total, p);
            }
            hec::to_affine(total, out_xy + AFF_W * k, is_inf ? is_inf + k : nullptr);
        }
        return ZKB_OK;
    }
    for (size_t k = 0; k < count; ++k) {
        hec::Pt p;
        memcpy(&p, part[k].data(), XYZZ_BYTES);
        hec::to_affine(p, out_xy + AFF_W * k, is_inf ? is_inf + k : nullptr);
    }
    return ZKB_OK;
}

// kzg10::commit for `count` polynomials resident in HBM (PolynomialCommitment::commit takes a batch: prove.rs:133-135
// commits a, b, c together, :178-180 t, h1, h2, :249-251 z1, z2, :306-308 the three quotient parts).
int zkb_commit_batch_dev(zkb_ctx *ctx, const uint64_t *const *coeffs_mont_dev, const size_t *offsets, const size_t *lens, size_t count,
                         uint64_t *out_xy /* count x 8 */, int *is_inf /* count */) {
    if (!ctx) return ZKB_ERR_INVALID;
    if ((!coeffs_mont_dev || !lens || !out_xy) && count) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_batch_dev: null argument");
    MsmState *st = state(ctx);
    if (!st->pipe_partial.empty()) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_commit_batch_dev: a zkb_commit_push batch is still open");
    int rc = ZKB_OK;
    st->pipe_expected = count ? count : 1;
    for (size_t k = 0; k < count && !rc; ++k) rc = zkb_commit_push(ctx, coeffs_mont_dev[k], offsets ? offsets[k] : 0, lens[k]);
    if (rc) {                                                              // drain and drop the partial batch
        std::vector<uint64_t> scratch(AFF_W * st->pipe_partial.size() + AFF_W);
        std::string err = ctx->err;
        zkb_commit_finish(ctx, scratch.data(), nullptr);
        ctx->err = err;
        return rc;
    }
    return zkb_commit_finish(ctx, out_xy, is_inf);
}

}  // extern "C"
