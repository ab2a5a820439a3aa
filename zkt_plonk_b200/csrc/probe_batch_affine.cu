// probe_batch_affine.cu -- EXPERIMENT (not on the product path): does batch-affine bucket accumulation pay on B200?
//
// DESIGN.md section 7, item 1.  The bucket accumulation spends 1232 multiply-adds per insertion (XYZZ mixed addition with
// the dedicated squaring and the fused Y3).  An affine + affine addition needs 2M + 1S and one inversion; with
// Montgomery's trick a batch shares one inversion at 3M per member: 5M + 1S = 788 multiply-adds.  The open questions are
// whether the shared inversion and the CTA-wide synchronisation can be hidden, and how many CTAs fit next to the
// shared-memory state.  This probe answers them in isolation, with no sorting and no scheduling:
//
//   mode 0   baseline: every thread owns one XYZZ accumulator in registers and adds `steps` points of a small table
//            (L2-resident) with g1x_add_mixed -- the inner loop of msm_accumulate_kernel;
//   mode 1   batch-affine: every thread owns M affine accumulators in shared memory; a step adds one table point to each
//            of the CTA's 128 * M accumulators through ONE inversion: per-thread prefix products of the x differences, a
//            product tree over the 128 thread totals in shared memory, the root inverted by thread 0 with a binary extended
//            Euclid (ALU pipe only: ~530 shift / subtract steps instead of ~380 dependent field products), the tree and
//            the prefixes walked back, then lambda = dy * inv, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1.
// Both modes add the same points to the same accumulator ids, so their results must agree point for point once the XYZZ
// sums are normalised; zkb_probe_batch_affine reports additions per second of both and the number of mismatches.
// Equal-x pairs (doubling / cancellation) are not handled: the table holds distinct random points, so they do not occur.
#include "ctx.h"
#include "ec.cuh"
#include "ff_inv.cuh"

using namespace zkb;

namespace {

constexpr int PB_THREADS = 128;

// which table point accumulator `id` receives at step `s` (never its own start point: offsets differ)
__device__ __forceinline__ uint32_t pb_start(uint32_t id, uint32_t tmask) { return (id * 40499u + 7u) & tmask; }
__device__ __forceinline__ uint32_t pb_index(uint32_t id, uint32_t s, uint32_t tmask) {
    uint32_t k = (id * 2654435761u + s * 40503u + 1u) & tmask;
    return k == pb_start(id, tmask) ? k ^ 1u : k;                // never the accumulator's own start point (a doubling)
}

// ---- mode 0: the XYZZ loop
__global__ void __launch_bounds__(PB_THREADS) pb_xyzz_kernel(const g1a_t *__restrict__ table, uint32_t tmask, uint32_t steps, g1x_t *out) {
    const uint32_t id = blockIdx.x * PB_THREADS + threadIdx.x;
    g1x_t acc = g1x_from_affine(g1a_load(table + pb_start(id, tmask)));
    for (uint32_t s = 0; s < steps; ++s) g1x_add_mixed(acc, g1a_load(table + pb_index(id, s, tmask)));
    g1x_store(out + id, acc);
}
// untimed: normalise the XYZZ sums for the comparison
__global__ void __launch_bounds__(PB_THREADS) pb_to_affine_kernel(const g1x_t *in, uint32_t n, g1a_t *out) {
    const uint32_t id = blockIdx.x * PB_THREADS + threadIdx.x;
    if (id >= n) return;
    g1a_t r = g1x_to_affine(g1x_load(in + id));
    fstore(&out[id].x, r.x);
    fstore(&out[id].y, r.y);
}

// ---- mode 1: batch-affine with one inversion per CTA and step
// shared memory: prefixes [M][128], tree [2 * 128] and, unless GACC, the accumulators [M][128] (x, y).  GACC keeps the
// accumulators in global memory (`out` itself: L2-resident, as the bucket array of the real kernel would be), which
// leaves room for more resident CTAs to hide the inversion of the others.
template <int M, bool GACC>
__global__ void __launch_bounds__(PB_THREADS) pb_affine_kernel(const g1a_t *__restrict__ table, uint32_t tmask, uint32_t steps, g1a_t *out) {
    extern __shared__ uint4 pb_sm[];
    fe_t *pre = reinterpret_cast<fe_t *>(pb_sm);                 // prefix products of the thread's differences [M * 128]
    fe_t *tree = pre + M * PB_THREADS;                           // node k = node 2k * node 2k+1; leaves at 128 + t
    fe_t *sax = tree + 2 * PB_THREADS;                           // shared-memory accumulators (unused with GACC)
    fe_t *say = sax + M * PB_THREADS;
    const uint32_t t = threadIdx.x;
    // accumulator ids are laid out so that mode 0's thread `id` and this kernel's slot (block, i, t) coincide
    auto gid = [&](int i) { return (blockIdx.x * M + i) * PB_THREADS + t; };
    auto px = [&](int i) -> fe_t * { return GACC ? &out[gid(i)].x : &sax[i * PB_THREADS + t]; };
    auto py = [&](int i) -> fe_t * { return GACC ? &out[gid(i)].y : &say[i * PB_THREADS + t]; };
    for (int i = 0; i < M; ++i) {
        g1a_t p = g1a_load(table + pb_start(gid(i), tmask));
        fstore(px(i), p.x);
        fstore(py(i), p.y);
    }
    for (uint32_t s = 0; s < steps; ++s) {
        // 1. differences and their running product
        fe_t run;
        for (int i = 0; i < M; ++i) {
            fe_t x2 = fload_ro(&table[pb_index(gid(i), s, tmask)].x);
            fe_t d = fsub<Q>(x2, fload(px(i)));
            run = i ? fmul<Q>(run, d) : d;
            fstore(&pre[i * PB_THREADS + t], run);
        }
        fstore(&tree[PB_THREADS + t], run);
        __syncthreads();
        // 2. product tree over the 128 thread totals
        for (int w = PB_THREADS / 2; w >= 1; w >>= 1) {
            if (t < (uint32_t)w) fstore(&tree[w + t], fmul<Q>(fload(&tree[2 * (w + t)]), fload(&tree[2 * (w + t) + 1])));
            __syncthreads();
        }
        // 3. ONE inversion
        if (t == 0) fstore(&tree[1], finv_euclid<Q>(fload(&tree[1])));
        __syncthreads();
        // 4. walk the tree back: inverse of a child = inverse of the parent * the sibling's product
        for (int w = 1; w < PB_THREADS; w <<= 1) {
            if (t < (uint32_t)w) {
                fe_t inv_parent = fload(&tree[w + t]), left = fload(&tree[2 * (w + t)]), right = fload(&tree[2 * (w + t) + 1]);
                fstore(&tree[2 * (w + t)], fmul<Q>(inv_parent, right));
                fstore(&tree[2 * (w + t) + 1], fmul<Q>(inv_parent, left));
            }
            __syncthreads();
        }
        // 5. back-substitution through the thread's own differences, then the additions
        fe_t inv_run = fload(&tree[PB_THREADS + t]);
        for (int i = M - 1; i >= 0; --i) {
            g1a_t p2 = g1a_load(table + pb_index(gid(i), s, tmask));
            fe_t x1 = fload(px(i)), y1 = fload(py(i));
            fe_t d = fsub<Q>(p2.x, x1);
            fe_t inv_d = i ? fmul<Q>(inv_run, fload(&pre[(i - 1) * PB_THREADS + t])) : inv_run;
            if (i) inv_run = fmul<Q>(inv_run, d);
            fe_t lam = fmul<Q>(fsub<Q>(p2.y, y1), inv_d);
            fe_t x3 = fsub<Q>(fsub<Q>(fsqr<Q>(lam), x1), p2.x);
            fe_t y3 = fsub<Q>(fmul<Q>(lam, fsub<Q>(x1, x3)), y1);
            fstore(px(i), x3);
            fstore(py(i), y3);
        }
        __syncthreads();                                          // tree and prefixes are rewritten by the next step
    }
    if (!GACC)
        for (int i = 0; i < M; ++i) {
            fstore(&out[gid(i)].x, fload(px(i)));
            fstore(&out[gid(i)].y, fload(py(i)));
        }
}

__global__ void pb_compare_kernel(const g1a_t *a, const g1a_t *b, uint32_t n, uint32_t *mismatches) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_t ax = fload(&a[i].x), ay = fload(&a[i].y), bx = fload(&b[i].x), by = fload(&b[i].y);
    if (!feq(ax, bx) || !feq(ay, by)) atomicAdd(mismatches, 1u);
}

template <int M, bool GACC>
int run_affine(zkb_ctx *ctx, const g1a_t *table, uint32_t tmask, uint32_t steps, uint32_t blocks, g1a_t *out, float *ms) {
    const size_t smem = ((size_t)(GACC ? 1 : 3) * M * PB_THREADS + 2 * PB_THREADS) * sizeof(fe_t);
    ZKB_CUDA(ctx, cudaFuncSetAttribute(pb_affine_kernel<M, GACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    ZKB_CUDA(ctx, cudaEventCreate(&e0));
    ZKB_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        ZKB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        pb_affine_kernel<M, GACC><<<blocks, PB_THREADS, smem, ctx->stream>>>(table, tmask, steps, out);
        ZKB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        ZKB_CUDA(ctx, cudaEventSynchronize(e1));
        ZKB_CUDA(ctx, cudaGetLastError());
        float t;
        ZKB_CUDA(ctx, cudaEventElapsedTime(&t, e0, e1));
        if (rep && t < best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = best;
    return ZKB_OK;
}

}  // namespace

extern "C" {

// table_dev: 2^log_table distinct affine points in HBM (e.g. from zkb_g1_fixed_base_mul_dev); m = accumulators per thread
// of the batch-affine kernel (4, 8 or 16; negative: the same with the accumulators in global memory instead of shared
// memory); steps = additions per accumulator.  out[0] = XYZZ additions / s, out[1] = batch-affine
// additions / s, *mismatches = accumulators whose two results differ (must be 0).
int zkb_probe_batch_affine(zkb_ctx *ctx, const uint64_t *table_dev, unsigned log_table, int m, unsigned steps, double out[2],
                           unsigned *mismatches) {
    if (!ctx || !table_dev || !out || !mismatches) return ZKB_ERR_INVALID;
    const bool gacc = m < 0;
    if (gacc) m = -m;
    if ((m != 4 && m != 8 && m != 16) || (m == 16 && !gacc) || log_table < 4 || log_table > 24 || steps == 0)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_probe_batch_affine: |m| in {4, 8, 16} (16 only with global accumulators), 4 <= log_table <= 24");
    const uint32_t tmask = (1u << log_table) - 1;
    const uint32_t blocks = (uint32_t)ctx->sm_count * 8;          // batch-affine CTAs; the XYZZ kernel runs blocks * m CTAs
    const uint32_t n_acc = blocks * (uint32_t)m * PB_THREADS;
    int rc = zkb_reserve(ctx, ctx->stage, (size_t)n_acc * (64 * 2 + 128) + 64);
    if (rc) return rc;
    g1a_t *res_x = (g1a_t *)ctx->stage.p, *res_a = res_x + n_acc;
    g1x_t *raw_x = (g1x_t *)(res_a + n_acc);
    uint32_t *d_mis = (uint32_t *)(raw_x + n_acc);
    const g1a_t *table = (const g1a_t *)table_dev;
    cudaEvent_t e0, e1;
    ZKB_CUDA(ctx, cudaEventCreate(&e0));
    ZKB_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        ZKB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        pb_xyzz_kernel<<<blocks * m, PB_THREADS, 0, ctx->stream>>>(table, tmask, steps, raw_x);
        ZKB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        ZKB_CUDA(ctx, cudaEventSynchronize(e1));
        ZKB_CUDA(ctx, cudaGetLastError());
        float t;
        ZKB_CUDA(ctx, cudaEventElapsedTime(&t, e0, e1));
        if (rep && t < best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    out[0] = (double)n_acc * steps / (best * 1e-3);
    pb_to_affine_kernel<<<blocks * m, PB_THREADS, 0, ctx->stream>>>(raw_x, n_acc, res_x);
    float ms = 0;
    if (!gacc) rc = m == 4 ? run_affine<4, false>(ctx, table, tmask, steps, blocks, res_a, &ms) : run_affine<8, false>(ctx, table, tmask, steps, blocks, res_a, &ms);
    else rc = m == 4 ? run_affine<4, true>(ctx, table, tmask, steps, blocks, res_a, &ms)
            : m == 8 ? run_affine<8, true>(ctx, table, tmask, steps, blocks, res_a, &ms) : run_affine<16, true>(ctx, table, tmask, steps, blocks, res_a, &ms);
    if (rc) return rc;
    out[1] = (double)n_acc * steps / (ms * 1e-3);
    ZKB_CUDA(ctx, cudaMemsetAsync(d_mis, 0, 4, ctx->stream));
    pb_compare_kernel<<<(n_acc + 255) / 256, 256, 0, ctx->stream>>>(res_x, res_a, n_acc, d_mis);
    ZKB_CUDA(ctx, cudaMemcpyAsync(mismatches, d_mis, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

}  // extern "C"
