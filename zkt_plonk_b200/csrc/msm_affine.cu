// msm_affine.cu -- EXPERIMENTAL bucket accumulation with batched affine additions.
// Off by default: msm.cu launches it instead of msm_accumulate_kernel only when the context's msm_mode is 1
// (zkb_msm_set_mode, or ZKB_MSM_MODE=1 in the environment when the context is created).  UNMEASURED on the GPU at the time
// of writing: the isolated probe (probe_batch_affine.cu) validated the arithmetic and the batch inversion, not this kernel.
//
// Same inputs and outputs as msm_accumulate_kernel (sorted entry list cut into tasks of <= SEG entries of one bucket, one
// result per task in XYZZ form), different inner loop: a thread advances M tasks in lock step with the rest of its CTA;
// step k adds the k-th entry of every task, and all 128 * M additions of the step share ONE field inversion:
//   phase 1  per slot: the denominator d (x2 - x1, or 2 y1 for a doubling) and the thread's running product of them
//            (prefixes in shared memory); product tree over the 128 thread totals; thread 0 inverts the root (binary
//            extended Euclid, ALU pipe); the tree is walked back
//   phase 2  per slot: 1/d from the thread's inverse total and the prefixes, then lambda, x3 = lambda^2 - x1 - x2,
//            y3 = lambda (x1 - x3) - y1
// 5M + 1S per insertion (788 multiply-adds + ~50 for its share of the tree) against 1232 for the XYZZ mixed addition.
// Accumulators are affine and live in the task's output slot (global memory, L2-resident); an accumulator that is still
// empty takes the point as it is, equal x means doubling (joins the batch with d = 2 y1) or cancellation (empties the
// accumulator).  The task order puts tasks of similar length next to each other, so a CTA's slots finish together.
#include "ctx.h"
#include "ec.cuh"
#include "ff_inv.cuh"

using namespace zkb;

namespace {

constexpr uint32_t SIGN_BIT = 0x80000000u;          // msm.cu: entry = point id | sign of the digit
constexpr int MA_THREADS = 128;
enum : uint32_t { ACT_NONE = 0, ACT_ADD = 1, ACT_DBL = 2, ACT_SET = 3, ACT_CANCEL = 4 };

__device__ __forceinline__ void prefetch_point_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p) + 32));
}

template <int M>
__global__ void __launch_bounds__(MA_THREADS) msm_accumulate_affine_kernel(
    const g1a_t *__restrict__ points, const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ counts,
    const uint32_t *__restrict__ starts, const uint32_t *__restrict__ ntasks, const uint32_t *__restrict__ task_base,
    const uint2 *__restrict__ task_order, const uint32_t *__restrict__ misc, uint32_t SEG, g1x_t *__restrict__ task_out,
    g1x_t *__restrict__ bucket_val) {
    extern __shared__ uint4 ma_sm[];
    fe_t *pre = reinterpret_cast<fe_t *>(ma_sm);                 // [M][128] running products of the thread's denominators
    fe_t *tree = pre + M * MA_THREADS;                           // [256] node k = node 2k * node 2k+1, leaves at 128 + t
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(tree + 2 * MA_THREADS);   // [M][128] entries of the slot's task
    uint32_t *s_off = s_cnt + M * MA_THREADS;                    // [M][128] first entry in `sorted`
    uint32_t *s_dst = s_off + M * MA_THREADS;                    // [M][128] output slot: index | (1 << 31 if in bucket_val)
    __shared__ uint32_t s_max;
    const uint32_t t = threadIdx.x;
    const uint32_t n_tasks = misc[1];
    if (t == 0) s_max = 0;
    __syncthreads();
    uint32_t my_max = 0;
    for (int i = 0; i < M; ++i) {
        const uint32_t T = (blockIdx.x * M + i) * MA_THREADS + t;
        uint32_t cnt = 0, off = 0, dst = 0;
        if (T < n_tasks) {
            const uint2 task = task_order[T];
            const uint32_t b = task.x, s = task.y;
            cnt = min(SEG, counts[b] - s * SEG);
            off = starts[b] + s * SEG;
            dst = ntasks[b] == 1 ? (b | 0x80000000u) : task_base[b] + s;
        }
        s_cnt[i * MA_THREADS + t] = cnt;
        s_off[i * MA_THREADS + t] = off;
        s_dst[i * MA_THREADS + t] = dst;
        my_max = max(my_max, cnt);
    }
    atomicMax(&s_max, my_max);
    __syncthreads();
    const uint32_t steps = s_max;
    auto slot = [&](int i) -> g1x_t * {
        const uint32_t d = s_dst[i * MA_THREADS + t];
        return (d & 0x80000000u) ? bucket_val + (d & 0x7fffffffu) : task_out + d;
    };
    uint32_t empty = (1u << M) - 1;                              // bit i: accumulator i holds no point yet
    const fe_t one = fone<Q>();

    for (uint32_t k = 0; k < steps; ++k) {
        // ---- phase 1: what each slot does in this step, its denominator, the thread's running product
        uint32_t actions = 0;
        fe_t run = one;
        for (int i = 0; i < M; ++i) {
            uint32_t act = ACT_NONE;
            if (k < s_cnt[i * MA_THREADS + t]) {
                const uint32_t v = sorted[s_off[i * MA_THREADS + t] + k];
                const g1a_t *pp = points + (v & ~SIGN_BIT);
                if (k + 1 < s_cnt[i * MA_THREADS + t]) prefetch_point_l2(points + (sorted[s_off[i * MA_THREADS + t] + k + 1] & ~SIGN_BIT));
                if ((empty >> i) & 1) {
                    act = ACT_SET;
                } else {
                    g1x_t *acc = slot(i);
                    fe_t d = fsub<Q>(fload_ro(&pp->x), fload(&acc->x));
                    act = ACT_ADD;
                    if (fis_zero<Q>(d)) {                        // same x: the same point (double it) or its negative (cancel)
                        fe_t y2 = fload_ro(&pp->y);
                        if (v & SIGN_BIT) y2 = fneg<Q>(y2);
                        fe_t y1 = fload(&acc->y);
                        if (feq(y2, y1)) { act = ACT_DBL; d = fdbl<Q>(y1); }
                        else act = ACT_CANCEL;
                    }
                    if (act != ACT_CANCEL) run = fmul<Q>(run, d);
                }
            }
            actions |= act << (3 * i);
            fstore(&pre[i * MA_THREADS + t], run);
        }
        fstore(&tree[MA_THREADS + t], run);
        __syncthreads();
        for (int w = MA_THREADS / 2; w >= 1; w >>= 1) {
            if (t < (uint32_t)w) fstore(&tree[w + t], fmul<Q>(fload(&tree[2 * (w + t)]), fload(&tree[2 * (w + t) + 1])));
            __syncthreads();
        }
        if (t == 0) fstore(&tree[1], finv_euclid<Q>(fload(&tree[1])));      // never zero: every factor is a non-zero field element
        __syncthreads();
        for (int w = 1; w < MA_THREADS; w <<= 1) {
            if (t < (uint32_t)w) {
                fe_t inv_parent = fload(&tree[w + t]), left = fload(&tree[2 * (w + t)]), right = fload(&tree[2 * (w + t) + 1]);
                fstore(&tree[2 * (w + t)], fmul<Q>(inv_parent, right));
                fstore(&tree[2 * (w + t) + 1], fmul<Q>(inv_parent, left));
            }
            __syncthreads();
        }
        // ---- phase 2: the additions
        fe_t inv_run = fload(&tree[MA_THREADS + t]);
        for (int i = M - 1; i >= 0; --i) {
            const uint32_t act = (actions >> (3 * i)) & 7u;
            if (act == ACT_NONE) continue;
            g1x_t *acc = slot(i);
            if (act == ACT_CANCEL) { empty |= 1u << i; continue; }
            const uint32_t v = sorted[s_off[i * MA_THREADS + t] + k];
            g1a_t p2 = g1a_load(points + (v & ~SIGN_BIT));
            if (v & SIGN_BIT) p2.y = fneg<Q>(p2.y);
            if (act == ACT_SET) {
                fstore(&acc->x, p2.x);
                fstore(&acc->y, p2.y);
                empty &= ~(1u << i);
                continue;
            }
            const fe_t x1 = fload(&acc->x), y1 = fload(&acc->y);
            const fe_t d = act == ACT_ADD ? fsub<Q>(p2.x, x1) : fdbl<Q>(y1);
            const fe_t inv_d = i ? fmul<Q>(inv_run, fload(&pre[(i - 1) * MA_THREADS + t])) : inv_run;
            if (i) inv_run = fmul<Q>(inv_run, d);
            fe_t num;
            if (act == ACT_ADD) {
                num = fsub<Q>(p2.y, y1);
            } else {
                fe_t xx = fsqr<Q>(x1);
                num = fadd<Q>(fdbl<Q>(xx), xx);                  // 3 x1^2 (a = 0)
            }
            const fe_t lam = fmul<Q>(num, inv_d);
            const fe_t x3 = fsub<Q>(fsub<Q>(fsqr<Q>(lam), x1), p2.x);       // doubling: p2.x == x1
            const fe_t y3 = fsub<Q>(fmul<Q>(lam, fsub<Q>(x1, x3)), y1);
            fstore(&acc->x, x3);
            fstore(&acc->y, y3);
        }
        __syncthreads();                                         // prefixes and tree are rewritten by the next step
    }
    // ---- results in the XYZZ form the combine / reduction kernels read
    for (int i = 0; i < M; ++i) {
        const uint32_t T = (blockIdx.x * M + i) * MA_THREADS + t;
        if (T >= n_tasks) continue;
        g1x_t *acc = slot(i);
        if ((empty >> i) & 1) {
            g1x_store(acc, g1x_inf());
        } else {
            fstore(&acc->zz, one);
            fstore(&acc->zzz, one);
        }
    }
}

}  // namespace

// launched by msm_enqueue (msm.cu) in place of msm_accumulate_kernel when ctx->msm_mode == 1
int zkb_launch_accumulate_affine(zkb_ctx *ctx, cudaStream_t s, uint64_t max_tasks, const void *points, const uint32_t *sorted,
                                 const uint32_t *counts, const uint32_t *starts, const uint32_t *ntasks, const uint32_t *task_base,
                                 const void *task_order, const uint32_t *misc, uint32_t seg, void *task_out, void *bucket_val) {
    constexpr int M = 8;
    const size_t smem = ((size_t)M * MA_THREADS + 2 * MA_THREADS) * sizeof(fe_t) + (size_t)3 * M * MA_THREADS * sizeof(uint32_t);
    ZKB_CUDA(ctx, cudaFuncSetAttribute(msm_accumulate_affine_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((max_tasks + (uint64_t)M * MA_THREADS - 1) / ((uint64_t)M * MA_THREADS));
    if (grid)
        msm_accumulate_affine_kernel<M><<<grid, MA_THREADS, smem, s>>>((const g1a_t *)points, sorted, counts, starts, ntasks, task_base,
                                                                     (const uint2 *)task_order, misc, seg, (g1x_t *)task_out,
                                                                     (g1x_t *)bucket_val);
    ZKB_CUDA(ctx, cudaGetLastError());
    return ZKB_OK;
}
