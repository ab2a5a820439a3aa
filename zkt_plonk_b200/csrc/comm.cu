// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference is single-process (rayon only: SURVEY.md 2.3), so nothing here replaces reference code; it is the
// exchange step of the point-range sharded commitments: every rank holds one contiguous range of the
// committer key (and its fixed-base tables), runs the bucket method on the matching slice of each polynomial, and
// the XYZZ partial sums (128 B per rank and commitment) are all-gathered and added identically on every rank.
//
// NCCL is loaded with dlopen at zkb_comm_init time, so libzkb200.so keeps no link-time dependency on it (a
// single-GPU host needs no NCCL at all) and a process that already loaded torch's bundled libnccl.so.2 shares it.
#include <dlfcn.h>
#include <nccl.h>

#include "ctx.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    if (api.handle || !api.err.empty()) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return &api;
    }
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(api.handle, name);
        if (!p) { ok = false; api.err = std::string("libnccl lacks ") + name; }
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    return &api;
}

#define ZKB_NCCL(ctx, api, call)                                                                  \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) {                                                                  \
            (ctx)->err = std::string(#call " failed: ") + (api)->GetErrorString(r_);              \
            return ZKB_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

}  // namespace

void zkb_comm_release(zkb_ctx *ctx) {
    if (ctx->comm) {
        NcclApi *api = nccl_api();
        if (api->handle) api->CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
    }
    if (ctx->comm_buf.p) { cudaFree(ctx->comm_buf.p); ctx->comm_buf = DevBuf(); }
    if (ctx->comm_pinned) { cudaFreeHost(ctx->comm_pinned); ctx->comm_pinned = nullptr; ctx->comm_pinned_bytes = 0; }
    if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
    ctx->rank = 0;
    ctx->world = 1;
}

int zkb_comm_allgather(zkb_ctx *ctx, const void *send_host, size_t bytes, void *recv_host) {
    if (ctx->world == 1) { memcpy(recv_host, send_host, bytes); return ZKB_OK; }
    if (!ctx->comm) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_comm_allgather: no communicator (zkb_comm_init)");
    NcclApi *api = nccl_api();
    const size_t total = bytes * (size_t)(ctx->world + 1);
    int rc = zkb_reserve(ctx, ctx->comm_buf, total);
    if (rc) return rc;
    if (ctx->comm_pinned_bytes < total) {
        if (ctx->comm_pinned) cudaFreeHost(ctx->comm_pinned);
        ctx->comm_pinned = nullptr;
        ctx->comm_pinned_bytes = 0;
        ZKB_CUDA(ctx, cudaMallocHost(&ctx->comm_pinned, total));
        ctx->comm_pinned_bytes = total;
    }
    // Host data in, host data out: nothing here depends on the context's stream, and the prover keeps that stream busy with
    // the next round's transforms while a batch of commitments is folded -- on it, the exchange queued behind them (8 B200:
    // 19.3 ms per 2^20-gate proof against 17.3 with the stream drained at every round, profiles/r02x_bench_n8.json).
    if (!ctx->comm_stream) {
        int lo_prio = 0, hi_prio = 0;
        ZKB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        ZKB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi_prio));
    }
    cudaStream_t cs = ctx->comm_stream;
    char *d_send = (char *)ctx->comm_buf.p, *d_recv = d_send + bytes;
    char *h_send = (char *)ctx->comm_pinned, *h_recv = h_send + bytes;
    memcpy(h_send, send_host, bytes);
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_send, h_send, bytes, cudaMemcpyHostToDevice, cs));
    ZKB_NCCL(ctx, api, api->AllGather(d_send, d_recv, bytes, ncclUint8, (ncclComm_t)ctx->comm, cs));
    ZKB_CUDA(ctx, cudaMemcpyAsync(h_recv, d_recv, bytes * (size_t)ctx->world, cudaMemcpyDeviceToHost, cs));
    ZKB_CUDA(ctx, cudaStreamSynchronize(cs));
    memcpy(recv_host, h_recv, bytes * (size_t)ctx->world);
    return ZKB_OK;
}

// in place on the device: buf[r * bytes_per_rank ..] holds rank r's slice afterwards (this rank's slice must already be there)
int zkb_comm_allgather_dev(zkb_ctx *ctx, void *buf_dev, size_t bytes_per_rank, cudaStream_t stream) {
    if (ctx->world == 1) return ZKB_OK;
    if (!ctx->comm) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_comm_allgather_dev: no communicator (zkb_comm_init)");
    NcclApi *api = nccl_api();
    ZKB_NCCL(ctx, api, api->AllGather((const char *)buf_dev + (size_t)ctx->rank * bytes_per_rank, buf_dev, bytes_per_rank, ncclUint8,
                                      (ncclComm_t)ctx->comm, stream));
    return ZKB_OK;
}

// Round 4 of a sharded proof: the nine coset NTTs are spread over the ranks (array k is transformed by rank k % world),
// every rank then needs only the slice of each result that its part of the quotient reads, plus a halo of 4 elements
// for the "next row" accesses.  One grouped exchange: the owner of an array sends every other rank its slice.
int zkb_comm_spread_slices(zkb_ctx *ctx, uint64_t *const bufs[], int count, size_t n_elems, size_t halo, cudaStream_t stream) {
    if (ctx->world == 1) return ZKB_OK;
    if (!ctx->comm) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_comm_spread_slices: no communicator (zkb_comm_init)");
    const size_t world = (size_t)ctx->world, chunk = n_elems / world;
    if (n_elems % world || halo > chunk) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_comm_spread_slices: the ranks must divide the array");
    NcclApi *api = nccl_api();
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    ZKB_NCCL(ctx, api, api->GroupStart());
    ncclResult_t bad = ncclSuccess;
    auto note = [&](ncclResult_t r) { if (r != ncclSuccess && bad == ncclSuccess) bad = r; };
    for (int k = 0; k < count; ++k) {
        const int owner = k % ctx->world;
        char *base = (char *)bufs[k];
        for (size_t r = 0; r < world; ++r) {
            if ((int)r == owner) continue;
            const size_t lo = r * chunk, hlo = ((r + 1) * chunk) % n_elems;       // slice and (cyclic) halo of rank r
            if (ctx->rank == owner) {
                note(api->Send(base + lo * 32, chunk * 32, ncclUint8, (int)r, comm, stream));
                if (halo) note(api->Send(base + hlo * 32, halo * 32, ncclUint8, (int)r, comm, stream));
            } else if ((size_t)ctx->rank == r) {
                note(api->Recv(base + lo * 32, chunk * 32, ncclUint8, owner, comm, stream));
                if (halo) note(api->Recv(base + hlo * 32, halo * 32, ncclUint8, owner, comm, stream));
            }
        }
    }
    ncclResult_t end = api->GroupEnd();
    if (bad != ncclSuccess || end != ncclSuccess) {
        ctx->err = std::string("zkb_comm_spread_slices: ") + api->GetErrorString(bad != ncclSuccess ? bad : end);
        return ZKB_ERR_CUDA;
    }
    return ZKB_OK;
}

extern "C" {

int zkb_comm_unique_id(uint8_t out[128]) {
    if (!out) return ZKB_ERR_INVALID;
    NcclApi *api = nccl_api();
    if (!api->handle) return ZKB_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return ZKB_ERR_CUDA;
    memcpy(out, &id, 128);
    return ZKB_OK;
}

int zkb_comm_init(zkb_ctx *ctx, const uint8_t id_bytes[128], int rank, int world) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (!id_bytes || world < 1 || rank < 0 || rank >= world) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_comm_init: bad rank / world / id");
    zkb_comm_release(ctx);
    if (world == 1) return ZKB_OK;
    NcclApi *api = nccl_api();
    if (!api->handle) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_comm_init: " + api->err);
    ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    ncclComm_t comm = nullptr;
    ZKB_NCCL(ctx, api, api->CommInitRank(&comm, world, id, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return ZKB_OK;
}

int zkb_comm_destroy(zkb_ctx *ctx) {
    if (!ctx) return ZKB_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    zkb_comm_release(ctx);
    return ZKB_OK;
}

int zkb_comm_rank(zkb_ctx *ctx) { return ctx ? ctx->rank : 0; }
int zkb_comm_world(zkb_ctx *ctx) { return ctx ? ctx->world : 1; }

int zkb_comm_allgather_host(zkb_ctx *ctx, const void *send_host, size_t bytes, void *recv_host) {
    if (!ctx) return ZKB_ERR_INVALID;
    if ((!send_host || !recv_host) && bytes) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_comm_allgather_host: null buffer");
    return zkb_comm_allgather(ctx, send_host, bytes, recv_host);
}

// The resident SRS (zkb_srs_load_g1*: srs_n points) is the range [global_lo, global_lo + srs_n) of a committer key
// of global_n powers.  Commitments (zkb_commit_dev / zkb_commit_batch_dev, hence zkb_plonk_setup / zkb_plonk_prove)
// then take global offsets and lengths, run on the overlap with the resident range and all-gather the partial sums.
int zkb_srs_set_range(zkb_ctx *ctx, size_t global_lo, size_t global_n) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (global_lo + ctx->srs_n > global_n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_set_range: the resident range exceeds the key");
    ctx->srs_lo = global_lo;
    ctx->srs_global_n = global_n;
    ctx->srs_replicated = false;
    return ZKB_OK;
}

// The other multi-GPU layout: EVERY rank holds the whole committer key (2^20 + 8 points are 64 MiB, with window tables
// under 1 GiB of 180), and the library decides per batch of commitments who computes what: one group of ranks per
// commitment ("fan-out": prove.rs:134,179,250,307 commit to 3 / 3 / 2 / 3 independent polynomials) or every commitment cut
// over all ranks -- see fanout_wins in msm.cu.  fanout: -1 = cost model, 0 = always shard, 1 = always fan out.
int zkb_srs_set_replicated(zkb_ctx *ctx, int fanout) {
    if (!ctx) return ZKB_ERR_INVALID;
    if (fanout < -1 || fanout > 1) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_set_replicated: fanout is -1 (cost model), 0 or 1");
    if (ctx->srs_lo != 0 || ctx->srs_global_n != ctx->srs_n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_set_replicated: this context holds a range of the key (zkb_srs_set_range)");
    ctx->srs_replicated = true;
    ctx->fanout = fanout;
    return ZKB_OK;
}

}  // extern "C"
