// lookup.cu -- round 2's witness plumbing on the device: LookupTable::into_multiset (lookup/table.rs:52-61),
// f = q_lookup (*) c (proof_system/prove.rs:157-161) and MultiSet::combine_split (lookup/multiset.rs:103-146).
//
// The reference walks an IndexMap over 2n elements on one thread.  The prover's default path (prover.cu) keeps that work on
// the host but makes it sparse -- it touches only the table and the rows with a lookup gate -- which is right while lookup
// gates are a few percent of the rows.  A circuit whose rows are mostly lookup gates (range checks) would pay one hash probe
// per row on one host thread; this file is the dense path:
//
//   host    the table's distinct values in order of first appearance (the bucket order of combine_split), their
//           multiplicities in t (with the zero padding of into_multiset), an open-addressing hash of value -> bucket:
//           O(table_len) work, a few KB to upload;
//   K1      one thread per row: f_i = q_lookup_i * c_i, hash probe, warp-aggregated atomicAdd into the bucket counts;
//   K2      one CTA: per bucket, how many copies go to h1 and to h2 (halves alternate on odd counts: a running parity)
//           and where they start -- two prefix sums and a prefix parity over <= table_len + 1 buckets;
//   K3      one thread per output row of h1 / h2: binary search of the row in the bucket offsets, copy the bucket's value.
//
// Same outputs as the host path bit for bit (t, f, h1, h2 as n Montgomery elements each); an element of f that the table does
// not hold sets status bit 0 (ElementNotIndexedInTable, multiset.rs:121), halves that are not n long set bit 1.
#include <string.h>

#include <unordered_map>
#include <vector>

#include "ctx.h"
#include "ff.cuh"

using namespace zkb;

namespace {

constexpr uint32_t NONE = 0xffffffffu;

__host__ __device__ inline uint32_t key_hash(const uint32_t *w /* 8 words */) {
    uint32_t h = 0x9E3779B9u;
    for (int i = 0; i < 8; ++i) {
        h ^= w[i] + 0x7F4A7C15u + (h << 6) + (h >> 2);
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
    }
    return h;
}

__global__ void __launch_bounds__(256) lookup_f_count_kernel(const uint4 *__restrict__ q_lookup, const uint4 *__restrict__ c, uint32_t n,
                                                             const uint4 *__restrict__ keys, const uint32_t *__restrict__ slot_bucket,
                                                             uint32_t mask, uint32_t zero_bucket, uint4 *__restrict__ f_out,
                                                             uint32_t *__restrict__ fcount, int *__restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = NONE;
    bool live = i < n;
    if (live) {
        const fe_t q = fload_ro(q_lookup + 2 * (size_t)i);
        fe_t f = fzero<FrP>();
        if (!fis_zero<FrP>(q)) {
            const fe_t ci = fload_ro(c + 2 * (size_t)i);
            f = feq(q, fone<FrP>()) ? ci : fmul<FrP>(q, ci);
        }
        fstore(f_out + 2 * (size_t)i, f);
        if (fis_zero<FrP>(f)) {
            b = zero_bucket;
        } else {
            uint32_t s = key_hash(f.v) & mask;
            for (;;) {                                              // linear probing; the table is at most half full
                const uint32_t sb = slot_bucket[s];
                if (sb == NONE) break;
                if (feq(fload_ro(keys + 2 * (size_t)s), f)) { b = sb; break; }
                s = (s + 1) & mask;
            }
        }
        if (b == NONE) { atomicOr(status, 1); live = false; }      // ElementNotIndexedInTable
    }
    // lookups pile on few buckets (the zero bucket above all): one atomic per distinct bucket of the warp
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t peers = __match_any_sync(0xffffffffu, live ? b : NONE);
    if (live && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&fcount[b], (uint32_t)__popc(peers));
}

// exclusive prefix sum over a CTA of 1024 threads; *total = the CTA's sum (sm: 32 words)
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *sm, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) sm[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = sm[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        sm[lane] = s;                                               // inclusive totals of the warps
    }
    __syncthreads();
    const uint32_t r = (wid ? sm[wid - 1] : 0u) + x - v;
    *total = sm[31];
    __syncthreads();
    return r;
}

// One CTA of 1024 threads.  cnt_b = tcount_b + fcount_b; parity_b = (number of odd counts before b) & 1;
// to h1: cnt / 2 + (odd && !parity), to h2: cnt / 2 + (odd && parity)  (multiset.rs:131-144); offsets = exclusive sums.
__global__ void __launch_bounds__(1024) lookup_offsets_kernel(const uint32_t *__restrict__ tcount, const uint32_t *__restrict__ fcount,
                                                              uint32_t nb, uint32_t n, uint32_t *__restrict__ off0,
                                                              uint32_t *__restrict__ off1, int *__restrict__ status) {
    __shared__ uint32_t sm[32];
    uint32_t odd_before = 0, k0 = 0, k1 = 0;                         // carried from chunk to chunk (the same in every thread)
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t cnt = b < nb ? tcount[b] + fcount[b] : 0;
        const uint32_t odd = cnt & 1, half = cnt >> 1;
        uint32_t tot_odd, tot0, tot1;
        const uint32_t par = (odd_before + block_exclusive_scan_1024(odd, sm, &tot_odd)) & 1;
        const uint32_t a0 = half + ((odd && !par) ? 1u : 0u), a1 = half + ((odd && par) ? 1u : 0u);
        const uint32_t e0 = block_exclusive_scan_1024(a0, sm, &tot0), e1 = block_exclusive_scan_1024(a1, sm, &tot1);
        if (b < nb) { off0[b] = k0 + e0; off1[b] = k1 + e1; }
        odd_before += tot_odd; k0 += tot0; k1 += tot1;
    }
    if (threadIdx.x == 0) {
        off0[nb] = k0;
        off1[nb] = k1;
        if (k0 != n || k1 != n) atomicOr(status, 2);                // |t| = |f| = n: both halves must be n long
    }
}

// h[j] = value of the bucket whose range [off[b], off[b + 1]) holds j  (blockIdx.y: 0 -> h1, 1 -> h2)
__global__ void __launch_bounds__(256) lookup_expand_kernel(const uint4 *__restrict__ values, const uint32_t *__restrict__ off0,
                                                            const uint32_t *__restrict__ off1, uint32_t nb, uint32_t n,
                                                            uint4 *__restrict__ h1, uint4 *__restrict__ h2) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t *off = blockIdx.y ? off1 : off0;
    uint4 *h = blockIdx.y ? h2 : h1;
    if (j >= off[nb]) { fstore(h + 2 * (size_t)j, fzero<FrP>()); return; }   // short halves (an error the status word reports)
    uint32_t lo = 0, hi = nb;                                         // last b with off[b] <= j  (off[0] = 0)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= j) lo = mid; else hi = mid;
    }
    fstore(h + 2 * (size_t)j, fload_ro(values + 2 * (size_t)lo));
}

struct KeyHashH {
    size_t operator()(const host::Fe &k) const { return key_hash((const uint32_t *)k.l); }
};
struct KeyEqH {
    bool operator()(const host::Fe &a, const host::Fe &b) const { return memcmp(a.l, b.l, 32) == 0; }
};

}  // namespace

extern "C" {

// t, f, h1, h2 (n = 2^log_n Montgomery elements each, device) from the table (host, table_len <= n Montgomery elements, in
// the order the composer holds them), q_lookup's evaluations and the wire c (device).  Everything is enqueued on the context's
// stream; *status_host (pinned or pageable; may be NULL) is valid after the stream has been synchronised: 0 = ok,
// bit 0 = ElementNotIndexedInTable, bit 1 = the halves are not n long.
int zkb_lookup_multisets_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t *table_host, size_t table_len, const uint64_t *q_lookup_evals_dev,
                             const uint64_t *c_evals_dev, uint64_t *t_dev, uint64_t *f_dev, uint64_t *h1_dev, uint64_t *h2_dev,
                             int *status_host) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (log_n > 30) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_lookup_multisets_dev: log_n out of range (the bucket counts are 32-bit: 2n must stay below 2^32)");
    const size_t n = (size_t)1 << log_n;
    if ((!table_host && table_len) || !q_lookup_evals_dev || !c_evals_dev || !t_dev || !f_dev || !h1_dev || !h2_dev)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_lookup_multisets_dev: null argument");
    if (table_len > n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_lookup_multisets_dev: the table is longer than the domain");
    // ---- host: buckets in order of first appearance in t = table || zeros (multiset.rs:103-118)
    const host::Fe *table = (const host::Fe *)table_host;
    std::vector<host::Fe> values;
    std::vector<uint32_t> tcount;
    std::unordered_map<host::Fe, uint32_t, KeyHashH, KeyEqH> index;
    index.reserve(2 * table_len + 16);
    uint32_t zero_bucket = NONE;
    for (size_t i = 0; i < table_len; ++i) {
        const host::Fe &e = table[i];
        if (host::is_zero(e)) {
            if (zero_bucket == NONE) { zero_bucket = (uint32_t)values.size(); values.push_back(e); tcount.push_back(0); }
            ++tcount[zero_bucket];
            continue;
        }
        auto it = index.find(e);
        if (it == index.end()) { index.emplace(e, (uint32_t)values.size()); values.push_back(e); tcount.push_back(1); }
        else ++tcount[it->second];
    }
    if (table_len < n) {                                           // the padding of LookupTable::into_multiset
        if (zero_bucket == NONE) { zero_bucket = (uint32_t)values.size(); host::Fe z; memset(z.l, 0, 32); values.push_back(z); tcount.push_back(0); }
        tcount[zero_bucket] += (uint32_t)(n - table_len);
    }
    const uint32_t nb = (uint32_t)values.size();
    uint32_t slots = 16;
    while (slots < 2 * (nb + 1)) slots <<= 1;
    std::vector<host::Fe> keys(slots);
    std::vector<uint32_t> slot_bucket(slots, NONE);
    memset(keys.data(), 0, (size_t)slots * 32);
    for (uint32_t b = 0; b < nb; ++b) {
        if (b == zero_bucket) continue;
        uint32_t s = key_hash((const uint32_t *)values[b].l) & (slots - 1);
        while (slot_bucket[s] != NONE) s = (s + 1) & (slots - 1);
        slot_bucket[s] = b;
        keys[s] = values[b];
    }
    // ---- device workspace: values | keys | slot_bucket | tcount | fcount | off0 | off1 | status
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_val = 0, o_keys = o_val + up((size_t)(nb + 1) * 32), o_sb = o_keys + up((size_t)slots * 32), o_tc = o_sb + up((size_t)slots * 4),
                 o_fc = o_tc + up((size_t)(nb + 1) * 4), o_o0 = o_fc + up((size_t)(nb + 1) * 4), o_o1 = o_o0 + up((size_t)(nb + 2) * 4),
                 o_st = o_o1 + up((size_t)(nb + 2) * 4), total = o_st + 256;
    int rc = zkb_reserve(ctx, ctx->lookup_ws, total);
    if (rc) return rc;
    char *w = (char *)ctx->lookup_ws.p;
    cudaStream_t s = ctx->stream;
    // the host vectors die with this call: the copies below are from pageable memory, i.e. staged before cudaMemcpyAsync returns
    if (nb) ZKB_CUDA(ctx, cudaMemcpyAsync(w + o_val, values.data(), (size_t)nb * 32, cudaMemcpyHostToDevice, s));
    ZKB_CUDA(ctx, cudaMemcpyAsync(w + o_keys, keys.data(), (size_t)slots * 32, cudaMemcpyHostToDevice, s));
    ZKB_CUDA(ctx, cudaMemcpyAsync(w + o_sb, slot_bucket.data(), (size_t)slots * 4, cudaMemcpyHostToDevice, s));
    if (nb) ZKB_CUDA(ctx, cudaMemcpyAsync(w + o_tc, tcount.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, s));
    ZKB_CUDA(ctx, cudaMemsetAsync(w + o_fc, 0, o_st + 256 - o_fc, s));                    // fcount, offsets, status
    ZKB_CUDA(ctx, cudaMemsetAsync(t_dev, 0, n * 32, s));
    if (table_len) ZKB_CUDA(ctx, cudaMemcpyAsync(t_dev, table_host, table_len * 32, cudaMemcpyHostToDevice, s));
    int *d_status = (int *)(w + o_st);
    lookup_f_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uint4 *)q_lookup_evals_dev, (const uint4 *)c_evals_dev, (uint32_t)n,
                                                                   (const uint4 *)(w + o_keys), (const uint32_t *)(w + o_sb), slots - 1,
                                                                   zero_bucket, (uint4 *)f_dev, (uint32_t *)(w + o_fc), d_status);
    lookup_offsets_kernel<<<1, 1024, 0, s>>>((const uint32_t *)(w + o_tc), (const uint32_t *)(w + o_fc), nb, (uint32_t)n,
                                             (uint32_t *)(w + o_o0), (uint32_t *)(w + o_o1), d_status);
    lookup_expand_kernel<<<dim3((unsigned)((n + 255) / 256), 2), 256, 0, s>>>((const uint4 *)(w + o_val), (const uint32_t *)(w + o_o0),
                                                                            (const uint32_t *)(w + o_o1), nb, (uint32_t)n, (uint4 *)h1_dev,
                                                                            (uint4 *)h2_dev);
    ZKB_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    if (status_host) ZKB_CUDA(ctx, cudaMemcpyAsync(status_host, d_status, sizeof(int), cudaMemcpyDeviceToHost, s));
    return ZKB_OK;
}

}  // extern "C"
