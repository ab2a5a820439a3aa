// api.cu -- extern "C" boundary of libzkb200.so (include/zkb200.h): context, memory, host-pointer wrappers.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ctx.h"
#include "ff.cuh"

using namespace zkb;

namespace {

template <class P>
__global__ void fp_binop_kernel(int op, uint4 *out, const uint4 *a, const uint4 *b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int Q4 = P::N / 4;                                  // 16-byte pieces per element
    fel_t<P::N> x = floadn<P::N>(a + Q4 * i), y = floadn<P::N>(b + Q4 * i), r;
    switch (op) {
        case 0: r = fmul<P>(x, y); break;
        case 1: r = fadd<P>(x, y); break;
        case 2: r = fsub<P>(x, y); break;
        case 3: r = fsqr<P>(x); break;
        case 4: r = finv<P>(x); break;
        case 5: r = fto_mont<P>(x); break;
        default: r = ffrom_mont<P>(x); break;
    }
    fstore(out + Q4 * i, r);
}


// ---- integer-pipe microbenchmarks (roofline denominators for the MSM / NTT kernels) ----
// mode 0: 32-bit IMAD (mad.lo.u32), mode 1: IMAD.WIDE.U32 (mad.wide.u32 + 64-bit accumulate),
// mode 2: dependent Montgomery products (fmul<FqP>), 4 independent chains per thread.
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, uint32_t iters, int mode) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 0) {
        uint32_t a[8], m = t | 1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = t + k;
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(a[(k + 1) & 7]));
        }
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r ^= a[k];
        out[t] = r;
    } else if (mode == 1) {
        unsigned long long a[8];
        uint32_t m = t | 1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = t + k;
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"((uint32_t)a[(k + 1) & 7]), "r"(m));
        }
        unsigned long long r = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r ^= a[k];
        out[t] = (uint32_t)(r ^ (r >> 32));
    } else if (mode == 3) {                                       // FP64 FMA (the other wide multiplier on the SM)
        double a[8], m = 1.0 + 1e-9 * (double)(t & 15);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (double)(t + k);
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(m), "d"(a[(k + 1) & 7]));
        }
        double r = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += a[k];
        out[t] = (uint32_t)__double2ll_rn(r);
    } else if (mode == 6 || mode == 7 || mode == 8) {            // pipe-sharing probes: IMAD + DFMA, IMAD + IADD3, DFMA + IADD3 in one thread
        uint32_t a[4], m = t | 1u, c[4];
        double d[4], md = 1.0 + 1e-9 * (double)(t & 15);
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = t + k; c[k] = t * 3 + k; d[k] = (double)(t + k); }
#define ZKB_PROBE_LOOP(BODY)                                  \
        for (uint32_t i = 0; i < iters; ++i) {                \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) { BODY } \
        }
#define ZKB_I asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(a[(k + 1) & 3]));
#define ZKB_D asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(md), "d"(d[(k + 1) & 3]));
#define ZKB_A asm volatile("add.u32 %0, %0, %1;" : "+r"(c[k]) : "r"(c[(k + 1) & 3])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(c[k]) : "r"(m));
        if (mode == 6) { ZKB_PROBE_LOOP(ZKB_I ZKB_D) }
        else if (mode == 7) { ZKB_PROBE_LOOP(ZKB_I ZKB_A) }
        else { ZKB_PROBE_LOOP(ZKB_D ZKB_A) }
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) r ^= a[k] ^ c[k] ^ (uint32_t)__double2ll_rn(d[k]);
        out[t] = r;
    } else {
        fq_t x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int j = 0; j < FqP::N; ++j) x[k].v[j] = (t * 2654435761u + k * 40503u + j) & 0x00ffffffu;
        }
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = fmul<FqP>(x[k], x[(k + 1) & 3]);
        }
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) r ^= x[k].v[0] ^ x[k].v[FqP::N - 1];
        out[t] = r;
    }
}

}  // namespace

extern "C" {

const char *zkb_version(void) { return "zkb200 0.1 (sm_100a, " ZKB_CURVE_NAME ")"; }

// The curve this library was compiled for (one shared object per curve, same entry points): 0 BN254, 1 BLS12-381, 2 BLS12-377;
// 64-bit words of a scalar (always 4) and of a base-field element (4 or 6: an affine point is twice, an XYZZ partial sum four
// times that); bit length of the scalar field's modulus; 1 when the prover driver (zkb_plonk_setup / zkb_plonk_prove) is compiled in
// (every build, like the key files and the pairing verifier; only EthereumTranscript is BN254's: ZKB_ERR_UNSUPPORTED elsewhere).
int zkb_curve_info(int *curve_id, int *fr_words, int *fq_words, int *fr_bits, int *has_prover) {
    if (curve_id) *curve_id = ZKB_CURVE;
    if (fr_words) *fr_words = host::FR_L;
    if (fq_words) *fq_words = host::FQ_L;
    if (fr_bits) *fr_bits = FrP::BITS;
    if (has_prover) *has_prover = 1;                              // zkb_plonk_setup / zkb_plonk_prove: every build
    return ZKB_OK;
}
// the G1 generator of the curve: affine, Montgomery form, 2 x fq_words words
int zkb_g1_generator(uint64_t *out_xy) {
    if (!out_xy) return ZKB_ERR_INVALID;
    memcpy(out_xy, host::G1_GEN_X, 8 * host::FQ_L);
    memcpy(out_xy + host::FQ_L, host::G1_GEN_Y, 8 * host::FQ_L);
    return ZKB_OK;
}

int zkb_ctx_create(int device, zkb_ctx **out) {
    if (!out) return ZKB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ZKB_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= count) return ZKB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return ZKB_ERR_CUDA;
    zkb_ctx *ctx = new zkb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    const char *mode = getenv("ZKB_MSM_MODE");                  // pair rounds (msm_pairs.cuh): "0".."6", or "-1" / "auto"
    if (mode && mode[0] >= '0' && mode[0] <= '6' && mode[1] == 0) ctx->msm_mode = mode[0] - '0';
    else if (mode && (!strcmp(mode, "-1") || !strcmp(mode, "auto"))) ctx->msm_mode = -1;
    const char *mp = getenv("ZKB_MSM_PARTS");                   // "dev_parts,host_parts,min_log": see zkb_msm_set_parts
    if (mp) {
        int d = 0, h = 0, m = 0;
        if (sscanf(mp, "%d,%d,%d", &d, &h, &m) == 3 && d >= 1 && d <= 7 && h >= 1 && h <= 7 && m >= 4 && m <= 31) {
            ctx->msm_parts_dev = d; ctx->msm_parts_host = h; ctx->msm_parts_min_log = m;
        }
    }
    const char *pf = getenv("ZKB_MSM_PF");                      // L2 prefetch distance of the accumulation's gathers (0..16)
    if (pf && atoi(pf) >= 0 && atoi(pf) <= 16) ctx->msm_prefetch = atoi(pf);
    const char *co = getenv("ZKB_MSM_COOP");                    // "0": one thread per addition in the binary reduction levels (A/B)
    if (co && (co[0] == '0' || co[0] == '1') && co[1] == 0) ctx->msm_coop = co[0] - '0';
    const char *nf = getenv("ZKB_NTT_FOLD");                    // "0": do not fold the coset factor into the inter-pass table (A/B)
    if (nf && nf[0] == '0' && nf[1] == 0) ctx->ntt_no_fold = true;
    const char *nk = getenv("ZKB_NTT_KERNEL");                  // "0" / "1" / "2": see zkb_ntt_set_kernel
    if (nk && nk[0] >= '0' && nk[0] <= '2' && nk[1] == 0) ctx->ntt_kernel = nk[0] - '0';
    *out = ctx;
    return ZKB_OK;
}

void zkb_ctx_destroy(zkb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->tables) cudaFree(kv.second.p);
    DevBuf *bufs[] = {&ctx->ntt_scratch, &ctx->stage, &ctx->ptr_stage, &ctx->srs, &ctx->poly_ws, &ctx->lookup_ws};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->len_slot) cudaFree(ctx->len_slot);
    zkb_msm_release(ctx);
    zkb_comm_release(ctx);
    delete ctx;
}

int zkb_ctx_set_stream(zkb_ctx *ctx, void *cuda_stream) {
    if (!ctx) return ZKB_ERR_INVALID;
    ctx->stream = (cudaStream_t)cuda_stream;
    return ZKB_OK;
}

int zkb_ctx_sync(zkb_ctx *ctx) {
    if (!ctx) return ZKB_ERR_INVALID;
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

uint64_t zkb_launch_count(zkb_ctx *ctx) { return ctx ? ctx->launches : 0; }

const char *zkb_last_error(zkb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int zkb_dev_alloc(zkb_ctx *ctx, size_t bytes, void **dptr) {
    if (!ctx || !dptr) return ZKB_ERR_INVALID;
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA;
    }
    return ZKB_OK;
}

int zkb_dev_free(zkb_ctx *ctx, void *dptr) {
    if (!ctx) return ZKB_ERR_INVALID;
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKB_CUDA(ctx, cudaFree(dptr));
    return ZKB_OK;
}

int zkb_h2d(zkb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    if (!ctx || (!dst_dev && bytes) || (!src_host && bytes)) return ZKB_ERR_INVALID;
    ZKB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

int zkb_d2h(zkb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    if (!ctx || (!dst_host && bytes) || (!src_dev && bytes)) return ZKB_ERR_INVALID;
    ZKB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

// ---------------------------------------------------------------------------------------------- NTT
int zkb_ntt_set_kernel(zkb_ctx *ctx, int kind) {
    if (!ctx || kind < 0 || kind > 2) return ZKB_ERR_INVALID;
    ctx->ntt_kernel = kind;
    return ZKB_OK;
}

int zkb_ntt_set_direct_tables(zkb_ctx *ctx, int enable) {
    if (!ctx) return ZKB_ERR_INVALID;
    ctx->ntt_no_direct = !enable;
    return ZKB_OK;
}

int zkb_ntt_dev(zkb_ctx *ctx, uint64_t *data_dev, size_t len, unsigned log_n, int inverse, int coset) {
    if (!ctx) return ZKB_ERR_INVALID;
    return zkb_ntt_run(ctx, data_dev, len, log_n, inverse, coset);
}

int zkb_ntt_batch_dev(zkb_ctx *ctx, uint64_t *const *ptrs_host, size_t count, size_t len, unsigned log_n,
                      int inverse, int coset) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!ptrs_host && count) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ntt_batch_dev: null pointer table");
    return zkb_ntt_run_batch(ctx, ptrs_host, count, len, log_n, inverse, coset);
}

int zkb_ntt(zkb_ctx *ctx, uint64_t *data_host, size_t len, unsigned log_n, int inverse, int coset) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!data_host) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ntt: null data pointer");
    if (log_n > host::FR_TWO_ADICITY || log_n > 31) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_ntt: log_n exceeds Fr TWO_ADICITY (28 on BN254) or 31");
    size_t n = (size_t)1 << log_n;
    if (len > n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_ntt: len > 2^log_n");
    int rc = zkb_reserve(ctx, ctx->stage, n * 32);
    if (rc) return rc;
    ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->stage.p, data_host, len * 32, cudaMemcpyHostToDevice, ctx->stream));
    rc = zkb_ntt_run(ctx, (uint64_t *)ctx->stage.p, len, log_n, inverse, coset);
    if (rc) return rc;
    ZKB_CUDA(ctx, cudaMemcpyAsync(data_host, ctx->stage.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

// ---------------------------------------------------------------------------------------------- test hooks
int zkb_test_fp_binop(zkb_ctx *ctx, int field, int op, uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n) {
    if (!ctx || !out || !a || !b) return ZKB_ERR_INVALID;
    const size_t eb = field == 0 ? 8 * host::FR_L : 8 * host::FQ_L;   // bytes per element
    int rc = zkb_reserve(ctx, ctx->stage, 3 * n * eb);
    if (rc) return rc;
    uint4 *da = (uint4 *)ctx->stage.p, *db = da + n * eb / 16, *dout = db + n * eb / 16;
    ZKB_CUDA(ctx, cudaMemcpyAsync(da, a, n * eb, cudaMemcpyHostToDevice, ctx->stream));
    ZKB_CUDA(ctx, cudaMemcpyAsync(db, b, n * eb, cudaMemcpyHostToDevice, ctx->stream));
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (field == 0) fp_binop_kernel<FrP><<<blocks, 128, 0, ctx->stream>>>(op, dout, da, db, n);
    else fp_binop_kernel<FqP><<<blocks, 128, 0, ctx->stream>>>(op, dout, da, db, n);
    ZKB_CUDA(ctx, cudaGetLastError());
    ZKB_CUDA(ctx, cudaMemcpyAsync(out, dout, n * eb, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

// ---------------------------------------------------------------------------------------------- microbenchmark
int zkb_bench_int(zkb_ctx *ctx, int mode, double *ops_per_sec) {
    if (!ctx || !ops_per_sec || mode < 0 || mode > 8 || mode == 4 || mode == 5) return ZKB_ERR_INVALID;
    const uint32_t blocks = ctx->sm_count * 8, threads = 256, iters = mode == 2 ? 512 : 4096;
    int rc = zkb_reserve(ctx, ctx->stage, (size_t)blocks * threads * 4);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    ZKB_CUDA(ctx, cudaEventCreate(&e0));
    ZKB_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        ZKB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        int_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t *)ctx->stage.p, iters, mode);
        ZKB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        ZKB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms;
        ZKB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double per_thread = (double)iters * (mode == 2 ? 4 : mode >= 6 ? 4 : 8);   // modes 6-8: per pair
    *ops_per_sec = per_thread * blocks * threads / (best * 1e-3);
    return ZKB_OK;
}

}  // extern "C"
