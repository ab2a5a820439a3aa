// ctx.h -- internal runtime state behind the opaque zkb_ctx of include/zkb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/zkb200.h"
#include "host_ff.h"

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct zkb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;       // kernels enqueued through this context (bench.py reports it)

    // ---- NTT: cached twiddle tables, keyed by a small integer id (see ntt.cu)
    std::map<uint64_t, DevBuf> tables;
    DevBuf ntt_scratch;          // ping-pong buffer for multi-pass transforms
    bool ntt_no_direct = false;  // true: never build the fully expanded twiddle / coset tables (saves N x 32 B each)
    bool ntt_no_fold = false;    // true: forward coset transforms stream the expanded coset table instead of folding g^c (A/B)
    int ntt_kernel = 0;          // 0: radix-4 pass kernel where it measured faster (2048-element tiles), 1: the generic kernel
                                 // everywhere, 2: the radix-4 kernel for every tile of 256..2048 elements
    DevBuf stage;                // staging buffer for host-pointer entry points
    DevBuf ptr_stage;

    // ---- MSM (msm.cu)
    DevBuf srs;                  // resident affine G1 points
    size_t srs_n = 0;
    void *msm_state = nullptr;   // opaque (owned by msm.cu)
    int msm_force_c = 0;         // 0 = cost model picks the window size
    int msm_prefetch = 0;        // L2 prefetch distance (points) of the bucket accumulation's gathers; 0 = none (measured best)
    int msm_coop = 1;            // binary levels of the bucket reduction with four lanes per addition (0: one thread per addition)
    int msm_parts_dev = 1, msm_parts_host = 4, msm_parts_min_log = 19;   // a large single MSM as point ranges through shared
                                 // buckets (msm.cu msm_run_parts; zkb_msm_set_parts)
    int msm_mode = 0;            // batched-affine pair rounds in front of the XYZZ accumulation (msm_pairs.cuh): 0 = none,
                                 // 1..6 = that many, -1 = chosen per MSM from the mean bucket load

    // ---- multi-GPU (comm.cu): one process per GPU, commitments sharded by point range
    void *comm = nullptr;        // ncclComm_t (NCCL is dlopen'ed: the library has no link-time dependency on it)
    int rank = 0, world = 1;
    size_t srs_lo = 0;           // global index of the first resident SRS point
    size_t srs_global_n = 0;     // size of the whole committer key (== srs_n unless this context holds one range of it)
    bool srs_replicated = false; // every rank holds the WHOLE key: the library splits each batch of commitments among the ranks
    int fanout = -1;             // replicated key, batches of several commitments: 1 = one group of ranks per commitment,
                                 // 0 = every commitment sharded over all ranks, -1 = cost model (msm.cu fanout_wins)
    cudaStream_t comm_stream = nullptr;   // host-to-host exchanges (zkb_comm_allgather) run here, not behind the context's stream
    DevBuf comm_buf;             // device staging of the all-gather
    void *comm_pinned = nullptr; // pinned host staging of the all-gather
    size_t comm_pinned_bytes = 0;

    // ---- elementwise / scan kernels (poly.cu)
    DevBuf poly_ws;
    DevBuf lookup_ws;            // bucket tables of zkb_lookup_multisets_dev (lookup.cu)
    int gp_failed = 0;           // the last grand product met a zero denominator
    unsigned long long *len_slot = nullptr;   // device word of zkb_poly_effective_len_dev
};

#define ZKB_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char b_[512];                                                                     \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                     __FILE__, __LINE__);                                                     \
            (ctx)->err = b_;                                                                  \
            return ZKB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ZKB_FAIL(ctx, code, msg) \
    do {                         \
        (ctx)->err = (msg);      \
        return (code);           \
    } while (0)

// grow-only device buffer
inline int zkb_reserve(zkb_ctx *ctx, DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes) return ZKB_OK;
    if (b.p) {
        ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ZKB_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
        b.p = nullptr;
        return e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA;
    }
    b.bytes = bytes;
    return ZKB_OK;
}

// Array k (n_elems field elements) is complete on rank k % world; afterwards every rank holds elements
// [rank * chunk, (rank + 1) * chunk + halo) (cyclically, chunk = n_elems / world) of every array in its own buffer k.
int zkb_comm_spread_slices(zkb_ctx *ctx, uint64_t *const bufs[], int count, size_t n_elems, size_t halo, cudaStream_t stream);
extern "C" int zkb_quotient_evals_range_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t challenges[20], const uint64_t *const wit[9],
                                            const uint64_t *const epk[11], uint64_t *out_dev, size_t lo, size_t hi);   // internal (hidden visibility)
// implemented in ntt.cu / msm.cu / poly.cu
int zkb_ntt_run(zkb_ctx *ctx, uint64_t *d_data, size_t len, unsigned log_n, int inverse, int coset);
int zkb_ntt_run_batch(zkb_ctx *ctx, uint64_t *const *d_ptrs, size_t count, size_t len, unsigned log_n, int inverse, int coset);
void zkb_msm_release(zkb_ctx *ctx);
void zkb_comm_release(zkb_ctx *ctx);
void zkb_commit_abort(zkb_ctx *ctx);   // drain and drop an open zkb_commit_push batch (error recovery)
// all ranks: recv_host[r * bytes ..] = rank r's send_host[0 .. bytes)  (NCCL all-gather on the context's stream; synchronous)
int zkb_comm_allgather(zkb_ctx *ctx, const void *send_host, size_t bytes, void *recv_host);
// the same in place on the device, enqueued on `stream` (no synchronisation)
int zkb_comm_allgather_dev(zkb_ctx *ctx, void *buf_dev, size_t bytes_per_rank, cudaStream_t stream);
// two-level power tables base^e = lo[e & (2^s - 1)] * hi[e >> s], e < 2^lm (Montgomery Fr); hi is pre-scaled by hi_scale
int zkb_pow2lvl_cached(zkb_ctx *ctx, uint64_t key, unsigned lm, const zkb::host::Fe &base, const zkb::host::Fe &hi_scale,
                       const void **out, uint32_t *s_out);
int zkb_pow2lvl_build(zkb_ctx *ctx, void *out, unsigned lm, const zkb::host::Fe &base, const zkb::host::Fe &hi_scale,
                      uint32_t *s_out);

// library-internal (msm.cu): <scalars, points> over caller-held device points plus extra_scalar * extra on the host
int zkb_msm_points_plus(zkb_ctx *ctx, const uint64_t *points_dev, const uint64_t *scalars_dev, size_t n, int scalars_mont,
                        const uint64_t *extra_xy, const uint64_t *extra_scalar_canon, uint64_t *out_xy, int *is_inf);
