// prover.cu -- C++ round driver: setup and the five prover rounds of zkt-plonk behind two C-ABI calls.
//
//   zkb_plonk_setup   proof_system/setup.rs:42-166 (10 iFFTs, 10 commitments, extend_prover_key: keys/mod.rs:78-146)
//   zkb_plonk_prove   proof_system/prove.rs:59-470 round by round, Proof serialised as proof.rs:106-155 derives it
//
// Everything heavy is a kernel call on HBM-resident polynomials (the entry points of zkb200.h); this file is host
// code: the Merlin transcript (merlin 3.0 over STROBE-128 / Keccak-f[1600], pinned by merlin's published test vector
// through the identical Python restatement), the TranscriptProtocol byte forms (transcript.rs:49-109), the
// host-side lookup plumbing (prove.rs:145-167: f = q_lookup * c, MultiSet::combine_split multiset.rs:103-146),
// challenge-dependent scalars (linearization_poly.rs:19-121) and ark-serialize's compressed point / field encodings.
// A Rust FFI crate calls these two functions with the composer's vectors.
#include <atomic>
#include <chrono>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ctx.h"
#include "host_plonk.h"

using namespace zkb;
using host::Fe;

namespace {

struct KeyHash {
    size_t operator()(const Fe &k) const {
        uint64_t h = k.l[0] * 0x9E3779B97F4A7C15ULL ^ (k.l[1] + 0xBF58476D1CE4E5B9ULL) * 0x94D049BB133111EBULL;
        h ^= (k.l[2] * 0xD6E8FEB86659FD93ULL) ^ (k.l[3] * 0xFF51AFD7ED558CCDULL);
        return (size_t)(h ^ (h >> 29));
    }
};
struct KeyEq { bool operator()(const Fe &a, const Fe &b) const { return feq(a, b); } };

// MultiSet::combine_split (multiset.rs:103-146) on Montgomery limb rows (unique per field element): buckets in order
// of first appearance in t, every element of f must be in t, halves alternate on odd counts; returns false on
// ElementNotIndexedInTable.  Written for the shape the prover meets: t = table entries followed by
// zero padding, f = q_lookup * c, which is zero outside the lookup gates.  Only the table and the lookup rows are touched: the
// zero bucket (~2n elements) is counted arithmetically and its halves are never written, because the staging
// columns are kept zero outside the small regions recorded in `dirty` (cleared before the next proof writes).
// h1 / h2 come out as [prefix | zeros | suffix]; dirty[h][0] = end of the prefix, dirty[h][1] = start of the suffix.
bool combine_split_sparse(const Fe *table, size_t table_len, size_t n, const Fe *f_rows /* f on the lookup rows */, size_t n_rows,
                          Fe *h1, Fe *h2, size_t dirty[2][2], size_t *n1, size_t *n2) {
    std::vector<std::pair<Fe, size_t>> buckets;
    std::unordered_map<Fe, size_t, KeyHash, KeyEq> index;
    index.reserve(2 * table_len + 16);
    size_t zero_bucket = (size_t)-1;
    for (size_t i = 0; i < table_len; ++i) {
        const Fe &e = table[i];
        if (host::is_zero(e)) {
            if (zero_bucket == (size_t)-1) { zero_bucket = buckets.size(); buckets.push_back({e, 0}); }
            ++buckets[zero_bucket].second;
            continue;
        }
        auto it = index.find(e);
        if (it == index.end()) { index.emplace(e, buckets.size()); buckets.push_back({e, 1}); }
        else ++buckets[it->second].second;
    }
    if (table_len < n) {                                           // the padding of LookupTable::into_multiset
        if (zero_bucket == (size_t)-1) { zero_bucket = buckets.size(); buckets.push_back({Fe{{0, 0, 0, 0}}, 0}); }
        buckets[zero_bucket].second += n - table_len;
    }
    size_t f_zeros = n - n_rows;                                   // rows without a lookup gate
    for (size_t r = 0; r < n_rows; ++r) {
        const Fe &e = f_rows[r];
        if (host::is_zero(e)) { ++f_zeros; continue; }
        auto it = index.find(e);
        if (it == index.end()) return false;
        ++buckets[it->second].second;
    }
    if (f_zeros) {
        if (zero_bucket == (size_t)-1) return false;               // a zero in f that the table does not hold
        buckets[zero_bucket].second += f_zeros;
    }
    Fe *h[2] = {h1, h2};
    for (int j = 0; j < 2; ++j) {                                  // forget what the previous proof left behind
        memset(h[j], 0, dirty[j][0] * sizeof(Fe));
        if (dirty[j][1] < n) memset(h[j] + dirty[j][1], 0, (n - dirty[j][1]) * sizeof(Fe));
        dirty[j][0] = 0;
        dirty[j][1] = n;
    }
    size_t k[2] = {0, 0};
    bool parity = false, past_zero = false;
    for (size_t b = 0; b < buckets.size(); ++b) {
        const size_t half = buckets[b].second / 2;
        const bool odd = buckets[b].second & 1;
        const size_t add0 = half + ((odd && !parity) ? 1 : 0), add1 = half + ((odd && parity) ? 1 : 0);
        if (k[0] + add0 > n || k[1] + add1 > n) return false;      // cannot happen for |t| = |f| = n; keeps the writes in bounds
        if (b == zero_bucket) {
            dirty[0][0] = k[0]; dirty[1][0] = k[1];
            k[0] += add0; k[1] += add1;
            dirty[0][1] = k[0]; dirty[1][1] = k[1];
            past_zero = true;
        } else {
            for (size_t j = 0; j < add0; ++j) h1[k[0] + j] = buckets[b].first;
            for (size_t j = 0; j < add1; ++j) h2[k[1] + j] = buckets[b].first;
            k[0] += add0; k[1] += add1;
        }
        if (odd) parity = !parity;
    }
    if (!past_zero) { dirty[0][0] = k[0]; dirty[1][0] = k[1]; }
    *n1 = k[0]; *n2 = k[1];
    return true;
}

// device polynomial: a slice of the key's arena (or an owned allocation for key material)
struct DPoly {
    uint64_t *d = nullptr;
    size_t len = 0;      // effective coefficients (DensePolynomial truncation)
    size_t cap = 0;
};

}  // namespace

struct zkb_plonk_pk {
    unsigned log_n = 0;
    size_t n = 0, table_size = 0;
    DPoly poly[10];                              // q_m q_l q_r q_o q_c q_lookup q_table sigma1 sigma2 sigma3 (coefficients)
    uint64_t *sigma_evals[3] = {nullptr, nullptr, nullptr};
    uint64_t *epk[11] = {};                      // q_m q_l q_r q_o q_c q_lookup q_table sigma1 sigma2 sigma3 l_1 (4n cosets)
    std::vector<uint32_t> lookup_rows;           // rows with q_lookup != 0 (the lookup gates)
    std::vector<Fe> lookup_q;                    // q_lookup on those rows, for f = q_lookup * c
    uint64_t *d_qlookup_evals = nullptr;         // q_lookup on the domain, in HBM: the device path of round 2's plumbing (lookup.cu)
    int lookup_mode = 0;                         // 0: device path when more than n / 8 rows are lookup gates, 1: host, 2: device
    int *lookup_status = nullptr;                // pinned word the device path reports into
    uint32_t *d_lookup_rows = nullptr;           // the same rows in HBM (f is scattered there from its compact upload)
    uint64_t *d_lookup_vals = nullptr;
    size_t dirty_t = 0;                          // non-zero regions the previous proof left in the pinned staging
    size_t dirty_h[2][2] = {{0, 0}, {0, 0}};
    std::vector<size_t> pi_pos;                  // sorted public-input rows
    Pt vk[10];                                   // q_m q_l q_r q_o q_c sigma1 sigma2 sigma3 q_lookup q_table (VerifierKey order)
    std::vector<void *> owned;                   // device allocations of the key
    Fe *stage = nullptr;                         // pinned host staging: t, f, h1, h2 (4 x n elements)
    Fe *wire_stage = nullptr;                    // pinned host staging of the three wires (3 x n elements)
    cudaStream_t copy_stream = nullptr;          // wire uploads (+ their all-gather on several GPUs), issued by the proving thread
    cudaStream_t lookup_stream = nullptr;        // uploads of the lookup multisets, issued by the worker thread: a stream of
                                                 // their own, so that 4n pinned elements never queue in front of a wire
    cudaEvent_t lookup_uploaded = nullptr, wire_uploaded = nullptr, wire_ev[3] = {nullptr, nullptr, nullptr};
    char *arena = nullptr;                       // per-proof scratch: reset at the start of every prove
    size_t arena_bytes = 0, arena_off = 0;
    int transcript_kind = 0;                     // 0 MerlinTranscript (default binary), 1 EthereumTranscript
    // ProvingComposer::wire_evals on the device (zkb_plonk_pk_set_wiring): the circuit's wire maps, 0 = Variable::Zero
    uint32_t *d_wiring[3] = {nullptr, nullptr, nullptr};   // w_l, w_r, w_o: n variable indices each, in HBM
    std::vector<uint32_t> lookup_wo;             // w_o on the lookup rows (the host side of round 2 needs c there)
    size_t n_vars_max = 0;                       // 1 + the largest variable index the wiring refers to
    uint64_t *d_vars = nullptr;                  // the assignment of the current proof in HBM (grown on demand)
    size_t d_vars_cap = 0;
};

namespace {

enum { P_QM, P_QL, P_QR, P_QO, P_QC, P_QLK, P_QT, P_S1, P_S2, P_S3 };

int dev_alloc_owned(zkb_ctx *ctx, zkb_plonk_pk *pk, size_t bytes, uint64_t **out) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 32);
    if (e != cudaSuccess) ZKB_FAIL(ctx, ZKB_ERR_OOM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    pk->owned.push_back(p);
    *out = (uint64_t *)p;
    return ZKB_OK;
}

uint64_t *arena_take(zkb_plonk_pk *pk, size_t elems) {
    size_t bytes = (elems * 32 + 255) / 256 * 256;
    if (pk->arena_off + bytes > pk->arena_bytes) return nullptr;
    uint64_t *p = (uint64_t *)(pk->arena + pk->arena_off);
    pk->arena_off += bytes + 9 * 512;            // skew: the quotient kernel walks ~20 buffers in lockstep; keep them off a
    return p;                                     // common power-of-two stride
}

#define TRY(call) do { int rc_ = (call); if (rc_ != ZKB_OK) return rc_; } while (0)
#define TAKE(ptr, elems) do { (ptr) = arena_take(const_cast<zkb_plonk_pk *>(pk), (elems)); \
    if (!(ptr)) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_plonk_prove: scratch arena exhausted"); } while (0)

// evals (host) -> coefficients on the device (poly_from_evals: iFFT + truncation)
int poly_from_evals_host(zkb_ctx *ctx, const uint64_t *evals_host, unsigned log_n, uint64_t *dst, size_t cap, DPoly *out) {
    const size_t n = (size_t)1 << log_n;
    ZKB_CUDA(ctx, cudaMemsetAsync(dst, 0, cap * 32, ctx->stream));
    ZKB_CUDA(ctx, cudaMemcpyAsync(dst, evals_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    TRY(zkb_ntt_dev(ctx, dst, n, log_n, 1, 0));
    out->d = dst; out->cap = cap;
    return zkb_poly_effective_len_dev(ctx, dst, n, &out->len);
}

int commit_many(zkb_ctx *ctx, const DPoly *const *polys, size_t count, Pt *out) {
    std::vector<const uint64_t *> ptrs(count);
    std::vector<size_t> lens(count);
    std::vector<uint64_t> xy(AFF_W * count);
    std::vector<int> inf(count);
    for (size_t k = 0; k < count; ++k) { ptrs[k] = polys[k]->d; lens[k] = polys[k]->len; }
    TRY(zkb_commit_batch_dev(ctx, ptrs.data(), nullptr, lens.data(), count, xy.data(), inf.data()));
    for (size_t k = 0; k < count; ++k) {
        out[k].inf = inf[k] != 0;
        memcpy(out[k].x.l, &xy[AFF_W * k], FQB);
        memcpy(out[k].y.l, &xy[AFF_W * k + AFF_W / 2], FQB);
    }
    return ZKB_OK;
}

// dst[pos[k]] = vals[k]  (PublicInputs::as_evals: pi.rs:75-82, a handful of rows in an otherwise zero column)
__global__ void scatter_fe_kernel(uint4 *dst, const unsigned long long *pos, const uint4 *vals, size_t count) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    dst[2 * pos[k]] = vals[2 * k];
    dst[2 * pos[k] + 1] = vals[2 * k + 1];
}

// ProvingComposer::wire_evals (prove.rs:49-55): wire[i] = value_of_var(w[i]); vars[0] is Variable::Zero's value (zero)
__global__ void __launch_bounds__(256) gather_wire_kernel(uint4 *__restrict__ wire, const uint32_t *__restrict__ w, const uint4 *__restrict__ vars, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t v = w[i];
    wire[2 * i] = __ldg(vars + 2 * v);
    wire[2 * i + 1] = __ldg(vars + 2 * v + 1);
}

__global__ void scatter_fe_rows_kernel(uint4 *dst, const uint32_t *rows, const uint4 *vals, size_t count) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    dst[2 * (size_t)rows[k]] = vals[2 * k];
    dst[2 * (size_t)rows[k] + 1] = vals[2 * k + 1];
}

int commit_finish_pts(zkb_ctx *ctx, size_t count, Pt *out) {
    std::vector<uint64_t> xy(AFF_W * count);
    std::vector<int> inf(count);
    TRY(zkb_commit_finish(ctx, xy.data(), inf.data()));
    for (size_t k = 0; k < count; ++k) {
        out[k].inf = inf[k] != 0;
        memcpy(out[k].x.l, &xy[AFF_W * k], FQB);
        memcpy(out[k].y.l, &xy[AFF_W * k + AFF_W / 2], FQB);
    }
    return ZKB_OK;
}

// The caller's vectors are pageable: cudaMemcpyAsync would stage them through the driver at ~12 GB/s with the host blocked
// (0.67 ms per 8 MiB wire on the B200 box, profiles/r02t_trace_2^18.jsonl).  Instead a few threads copy 1 MiB pieces into pinned
// memory and each piece is handed to the DMA engine as soon as it is staged, so staging and PCIe overlap.
// Returns false if a CUDA call failed.  Runs on the uploader thread of a proof.
bool staged_upload(int device, uint64_t *dst_dev, const void *src_host, void *pinned, size_t bytes, cudaStream_t stream) {
    const size_t piece = (size_t)1 << 20, pieces = (bytes + piece - 1) / piece;
    // threads: one per 8 MiB, 1..4.  More starve the DMA engine of host memory bandwidth (8 threads: 18.8 instead of 17.6 ms
    // per 2^18-gate proof, one or two: 17.6), a single one cannot feed it at 32 MiB per wire (55.8 instead of 51.5 ms at
    // 2^20; profiles/r02v_ab_prove*.jsonl); the 1/world slices of a multi-GPU proof take one thread per rank
    const unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>(4, std::max<size_t>(1, bytes >> 23));
    if (hw < 4) nt = 1;
    if (const char *e = getenv("ZKB_STAGE_THREADS")) nt = (size_t)std::max(1, atoi(e));   // for A/B measurements
    nt = pieces < nt ? pieces : nt;
    std::atomic<bool> ok{true};
    auto run = [&](size_t t) {
        if (cudaSetDevice(device) != cudaSuccess) { ok.store(false); return; }
        for (size_t i = t; i < pieces; i += nt) {
            const size_t lo = i * piece, len = lo + piece < bytes ? piece : bytes - lo;
            memcpy((char *)pinned + lo, (const char *)src_host + lo, len);
            if (cudaMemcpyAsync((char *)dst_dev + lo, (char *)pinned + lo, len, cudaMemcpyHostToDevice, stream) != cudaSuccess) ok.store(false);
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nt; ++t) pool.emplace_back(run, t);
    run(0);
    for (auto &t : pool) t.join();
    return ok.load();
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// test hook (host only, no GPU): the prover's sparse combine_split on caller-owned persistent staging
int zkb_test_combine_split(const uint64_t *table, size_t table_len, size_t n, const uint64_t *f, const uint32_t *rows, size_t n_rows,
                           uint64_t *h1, uint64_t *h2, size_t dirty[4], size_t out_lens[2]) {
    if ((!table && table_len) || !f || (!rows && n_rows) || !h1 || !h2 || !dirty || !out_lens) return ZKB_ERR_INVALID;
    size_t d[2][2] = {{dirty[0], dirty[1]}, {dirty[2], dirty[3]}};
    std::vector<Fe> f_rows(n_rows);
    for (size_t r = 0; r < n_rows; ++r) {
        if (rows[r] >= n) return ZKB_ERR_INVALID;
        f_rows[r] = fe_from(f + 4 * (size_t)rows[r]);
    }
    bool ok = combine_split_sparse((const Fe *)table, table_len, n, f_rows.data(), n_rows, (Fe *)h1, (Fe *)h2, d, &out_lens[0], &out_lens[1]);
    dirty[0] = d[0][0]; dirty[1] = d[0][1]; dirty[2] = d[1][0]; dirty[3] = d[1][1];
    return ok ? ZKB_OK : ZKB_ERR_INVALID;
}

int zkb_test_transcript(int kind, const uint8_t *ops, size_t n_ops, const uint64_t *args, uint8_t *challenges_out) {
    if ((kind != 0 && kind != 1) || (!ops && n_ops) || !args || !challenges_out) return ZKB_ERR_INVALID;
    Transcript tr("test", kind);
    size_t n_ch = 0;
    for (size_t i = 0; i < n_ops; ++i) {
        const uint64_t *a = args + (AFF_W + 1) * i;                // 9 words per op on BN254
        switch (ops[i]) {
            case 0: tr.append_u64("a", a[0]); break;
            case 1: tr.append_scalar("b", fe_from(a)); break;
            case 2: { Pt p; p.x = fq_from(a); p.y = fq_from(a + AFF_W / 2); p.inf = a[AFF_W] != 0; tr.append_commitment("c", p); break; }
            case 3: { Fe c = tr.challenge_scalar("a"); fe_bytes(c, host::FR, challenges_out + 32 * n_ch++); break; }
            default: return ZKB_ERR_INVALID;
        }
    }
    return ZKB_OK;
}

// Where round 2's witness plumbing (t, f = q_lookup * c, combine_split) runs: 0 = on the device when more than n / 8 rows are
// lookup gates, else sparse on a host thread under round 1 (default); 1 = always the host path; 2 = always the device path.
// The proof bytes are the same either way.
int zkb_plonk_pk_set_lookup_mode(zkb_plonk_pk *pk, int mode) {
    if (!pk || mode < 0 || mode > 2) return ZKB_ERR_INVALID;
    pk->lookup_mode = mode;
    return ZKB_OK;
}

int zkb_plonk_pk_set_transcript(zkb_plonk_pk *pk, int kind) {
    if (!pk || (kind != 0 && kind != 1)) return ZKB_ERR_INVALID;
    if (kind == 1 && ZKB_CURVE != ZKB_CURVE_BN254) return ZKB_ERR_UNSUPPORTED;   // EthereumTranscript is bound to Bn254 upstream
    pk->transcript_kind = kind;
    return ZKB_OK;
}

// bytes of a serialised Proof on this build's curve (proof.rs:112-154): 802 on BN254, 1010 on BLS12-381 / BLS12-377
size_t zkb_plonk_proof_bytes(void) { return PROOF_BYTES; }

void zkb_plonk_pk_destroy(zkb_ctx *ctx, zkb_plonk_pk *pk) {
    if (!pk) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (void *p : pk->owned) cudaFree(p);
    if (pk->arena) cudaFree(pk->arena);
    if (pk->stage) cudaFreeHost(pk->stage);
    if (pk->wire_stage) cudaFreeHost(pk->wire_stage);
    if (pk->d_vars) cudaFree(pk->d_vars);
    if (pk->copy_stream) cudaStreamDestroy(pk->copy_stream);
    if (pk->lookup_status) cudaFreeHost(pk->lookup_status);
    if (pk->lookup_stream) cudaStreamDestroy(pk->lookup_stream);
    if (pk->lookup_uploaded) cudaEventDestroy(pk->lookup_uploaded);
    if (pk->wire_uploaded) cudaEventDestroy(pk->wire_uploaded);
    for (cudaEvent_t e : pk->wire_ev) if (e) cudaEventDestroy(e);
    delete pk;
}

namespace {

// shared head of the two ways to build a key: argument checks and the empty key object
int key_begin(zkb_ctx *ctx, unsigned log_n, size_t table_size, const size_t *pi_positions, size_t n_pi, zkb_plonk_pk **pk_out) {
    if (log_n + 2 > host::FR_TWO_ADICITY || log_n + 2 > 30 || log_n < 3) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_plonk_setup: need 8 <= n and 4n <= 2^TWO_ADICITY (28 on BN254; InvalidEvalDomainSize)");
    const size_t n = (size_t)1 << log_n;
    if (table_size >= n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_setup: max table size is equal or larger than n (lookup/table.rs:43)");
    if (n + 8 > ctx->srs_global_n) ZKB_FAIL(ctx, ZKB_ERR_NO_SRS, "zkb_plonk_setup: the committer key must hold at least n + 8 powers");
    for (size_t i = 0; i < n_pi; ++i)
        if (pi_positions[i] >= n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_setup: public-input row outside the domain");
    zkb_plonk_pk *pk = new zkb_plonk_pk();
    pk->log_n = log_n; pk->n = n; pk->table_size = table_size;
    pk->pi_pos.assign(pi_positions, pi_positions + n_pi);
    *pk_out = pk;
    return ZKB_OK;
}

// shared tail: lookup rows to HBM, verifier-key commitments (unless given), the extended key, arena, staging, streams.
// Expects pk->poly[], pk->sigma_evals[], pk->lookup_rows / lookup_q to be in place.
int key_finish(zkb_ctx *ctx, zkb_plonk_pk *pk, bool have_vk) {
    const unsigned log_n = pk->log_n;
    const size_t n = pk->n, n4 = 4 * n;
    {
        uint64_t *rows_dev, *vals_dev;
        const size_t n_rows = pk->lookup_rows.size();
        int rc = dev_alloc_owned(ctx, pk, n_rows * 4 + 32, &rows_dev);
        if (rc) return rc;
        rc = dev_alloc_owned(ctx, pk, n_rows * 32 + 32, &vals_dev);
        if (rc) return rc;
        pk->d_lookup_rows = (uint32_t *)rows_dev;
        pk->d_lookup_vals = vals_dev;
        if (n_rows && cudaMemcpy(pk->d_lookup_rows, pk->lookup_rows.data(), n_rows * 4, cudaMemcpyHostToDevice) != cudaSuccess)
            ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_setup: H2D copy failed");
    }
    int rc;
    {                                                                      // q_lookup on the domain (forward NTT of its coefficients)
        const DPoly &p = pk->poly[P_QLK];
        rc = dev_alloc_owned(ctx, pk, n * 32, &pk->d_qlookup_evals);
        if (rc) return rc;
        if (cudaMemsetAsync(pk->d_qlookup_evals, 0, n * 32, ctx->stream) != cudaSuccess ||
            (p.len && cudaMemcpyAsync(pk->d_qlookup_evals, p.d, p.len * 32, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess))
            ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_setup: device copy failed");
        rc = zkb_ntt_dev(ctx, pk->d_qlookup_evals, p.len, log_n, 0, 0);
        if (rc) return rc;
        if (cudaMallocHost((void **)&pk->lookup_status, 64) != cudaSuccess) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_plonk_setup: cannot allocate pinned staging");
        *pk->lookup_status = 0;
    }
    if (!have_vk) {                                                        // verifier key commitments in VerifierKey order (setup.rs:104-121)
        const int vk_order[10] = {P_QM, P_QL, P_QR, P_QO, P_QC, P_S1, P_S2, P_S3, P_QLK, P_QT};
        const DPoly *cp[10];
        for (int k = 0; k < 10; ++k) cp[k] = &pk->poly[vk_order[k]];
        rc = commit_many(ctx, cp, 10, pk->vk);
        if (rc) return rc;
    }
    // extended key: 10 coset tables on 4n + l_1 (keys/mod.rs:96-119); x_coset / zh_coset live inside the quotient kernel
    for (int k = 0; k < 11; ++k) {
        rc = dev_alloc_owned(ctx, pk, n4 * 32, &pk->epk[k]);
        if (rc) return rc;
        if (k < 10) {
            const DPoly &p = pk->poly[k];
            if (cudaMemsetAsync(pk->epk[k], 0, n4 * 32, ctx->stream) != cudaSuccess ||
                cudaMemcpyAsync(pk->epk[k], p.d, p.len * 32, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
                ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_setup: device copy failed");
            rc = zkb_ntt_dev(ctx, pk->epk[k], p.len, log_n + 2, 0, 1);
        } else {
            rc = zkb_l1_coset_dev(ctx, log_n, pk->epk[k]);
        }
        if (rc) return rc;
    }
    // scratch arena for one proof: 9 witness cosets + the quotient (4n each) and ~30 n-sized buffers
    pk->arena_bytes = (10 * n4 + 34 * (n + 16)) * 32 + 64 * 9 * 512;
    if (cudaMalloc((void **)&pk->arena, pk->arena_bytes) != cudaSuccess)
        ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_plonk_setup: cannot allocate the prover arena");
    if (cudaMallocHost((void **)&pk->stage, 4 * n * sizeof(Fe)) != cudaSuccess)
        ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_plonk_setup: cannot allocate pinned staging");
    if (cudaMallocHost((void **)&pk->wire_stage, 3 * n * sizeof(Fe)) != cudaSuccess)
        ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_plonk_setup: cannot allocate pinned staging");
    memset(pk->stage, 0, 4 * n * sizeof(Fe));                              // kept zero outside the regions a proof writes
    pk->dirty_h[0][1] = pk->dirty_h[1][1] = n;
    if (cudaStreamCreateWithFlags(&pk->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&pk->lookup_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&pk->lookup_uploaded, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pk->wire_uploaded, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pk->wire_ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pk->wire_ev[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pk->wire_ev[2], cudaEventDisableTiming) != cudaSuccess)
        ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_setup: cannot create the copy stream");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_setup: stream error");
    return ZKB_OK;
}

}  // namespace

int zkb_plonk_setup(zkb_ctx *ctx, unsigned log_n, const uint64_t *const selectors[6], const uint64_t *const sigma[3],
                    size_t table_size, const size_t *pi_positions, size_t n_pi, zkb_plonk_pk **out) {
    if (!ctx || !out) return ZKB_ERR_INVALID;
    *out = nullptr;
    if (!selectors || !sigma || (!pi_positions && n_pi)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_setup: null argument");
    zkb_plonk_pk *pk = nullptr;
    int rc = key_begin(ctx, log_n, table_size, pi_positions, n_pi, &pk);
    if (rc) return rc;
    const size_t n = pk->n;
    auto fail = [&](int rc) { zkb_plonk_pk_destroy(ctx, pk); return rc; };
    // selector / sigma polynomials (setup.rs:72-90) and the q_table mask (lookup/table.rs:42-48)
    const uint64_t *src[10] = {selectors[0], selectors[1], selectors[2], selectors[3], selectors[4], selectors[5], nullptr,
                               sigma[0], sigma[1], sigma[2]};
    std::vector<Fe> qtable(n);
    for (size_t i = 0; i < n; ++i) qtable[i] = i < table_size ? Fe{{0, 0, 0, 0}} : FR_ONE();
    src[P_QT] = (const uint64_t *)qtable.data();
    for (int k = 0; k < 10; ++k) {
        if (!src[k]) return fail((ctx->err = "zkb_plonk_setup: null selector / sigma column", ZKB_ERR_INVALID));
        uint64_t *d;
        rc = dev_alloc_owned(ctx, pk, n * 32, &d);
        if (rc) return fail(rc);
        rc = poly_from_evals_host(ctx, src[k], log_n, d, n, &pk->poly[k]);
        if (rc) return fail(rc);
    }
    for (int k = 0; k < 3; ++k) {
        rc = dev_alloc_owned(ctx, pk, n * 32, &pk->sigma_evals[k]);
        if (rc) return fail(rc);
        if (cudaMemcpyAsync(pk->sigma_evals[k], sigma[k], n * 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
            return fail((ctx->err = "zkb_plonk_setup: H2D copy failed", ZKB_ERR_CUDA));
    }
    for (size_t i = 0; i < n; ++i) {
        const Fe q = fe_from(selectors[5] + 4 * i);
        if (!host::is_zero(q)) { pk->lookup_rows.push_back((uint32_t)i); pk->lookup_q.push_back(q); }
    }
    rc = key_finish(ctx, pk, false);
    if (rc) return fail(rc);
    *out = pk;
    return ZKB_OK;
}

// The same key from what the reference's `compile` wrote: the ten ProverKey polynomials in coefficient form (pk file
// order: q_m q_l q_r q_o q_c sigma1 sigma2 sigma3 q_lookup q_table) and, optionally, the VerifierKey commitments.  The
// evaluation tables the prover needs (sigma1..3 and q_lookup over the domain, keys/mod.rs:121-145) come back by forward
// NTTs, exactly (the transforms are bijections), so a key built this way proves byte-identically to one built by
// zkb_plonk_setup from the composer's columns.
int zkb_plonk_pk_from_polys(zkb_ctx *ctx, unsigned log_n, const uint64_t *const polys[10], const size_t lens[10], size_t table_size,
                            const size_t *pi_positions, size_t n_pi, const uint64_t *vk_xy, const int *vk_inf, zkb_plonk_pk **out) {
    if (!ctx || !out) return ZKB_ERR_INVALID;
    *out = nullptr;
    if (!polys || !lens || (!pi_positions && n_pi)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_pk_from_polys: null argument");
    zkb_plonk_pk *pk = nullptr;
    int rc = key_begin(ctx, log_n, table_size, pi_positions, n_pi, &pk);
    if (rc) return rc;
    const size_t n = pk->n;
    auto fail = [&](int rc) { zkb_plonk_pk_destroy(ctx, pk); return rc; };
    const int file_to_key[10] = {P_QM, P_QL, P_QR, P_QO, P_QC, P_S1, P_S2, P_S3, P_QLK, P_QT};
    for (int f = 0; f < 10; ++f) {
        const int k = file_to_key[f];
        if (lens[f] > n || (!polys[f] && lens[f]))
            return fail((ctx->err = "zkb_plonk_pk_from_polys: a key polynomial has more than n coefficients", ZKB_ERR_INVALID));
        uint64_t *d;
        rc = dev_alloc_owned(ctx, pk, n * 32, &d);
        if (rc) return fail(rc);
        if (cudaMemsetAsync(d, 0, n * 32, ctx->stream) != cudaSuccess ||
            (lens[f] && cudaMemcpyAsync(d, polys[f], lens[f] * 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess))
            return fail((ctx->err = "zkb_plonk_pk_from_polys: H2D copy failed", ZKB_ERR_CUDA));
        pk->poly[k].d = d;
        pk->poly[k].cap = n;
        size_t len = lens[f];                                              // DensePolynomial invariant: no trailing zeros
        while (len && !(polys[f][4 * (len - 1)] | polys[f][4 * (len - 1) + 1] | polys[f][4 * (len - 1) + 2] | polys[f][4 * (len - 1) + 3])) --len;
        pk->poly[k].len = len;
    }
    // evaluation tables: sigma1..3 for z1 (keys/permutation.rs:74-92), q_lookup for f = q_lookup * c (prove.rs:157-161)
    for (int k = 0; k < 4; ++k) {
        uint64_t *d;
        rc = dev_alloc_owned(ctx, pk, n * 32, &d);
        if (rc) return fail(rc);
        const DPoly &p = pk->poly[k < 3 ? P_S1 + k : P_QLK];
        if (cudaMemcpyAsync(d, p.d, n * 32, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
            return fail((ctx->err = "zkb_plonk_pk_from_polys: device copy failed", ZKB_ERR_CUDA));
        rc = zkb_ntt_dev(ctx, d, p.len, log_n, 0, 0);
        if (rc) return fail(rc);
        if (k < 3) { pk->sigma_evals[k] = d; continue; }
        std::vector<Fe> q(n);
        if (cudaMemcpyAsync(q.data(), d, n * 32, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            return fail((ctx->err = "zkb_plonk_pk_from_polys: D2H copy failed", ZKB_ERR_CUDA));
        for (size_t i = 0; i < n; ++i)
            if (!host::is_zero(q[i])) { pk->lookup_rows.push_back((uint32_t)i); pk->lookup_q.push_back(q[i]); }
    }
    if (vk_xy) {
        for (int k = 0; k < 10; ++k) {
            pk->vk[k].x = fq_from(vk_xy + AFF_W * k);
            pk->vk[k].y = fq_from(vk_xy + AFF_W * k + AFF_W / 2);
            pk->vk[k].inf = (vk_inf && vk_inf[k]) || (host::is_zero(pk->vk[k].x) && host::is_zero(pk->vk[k].y));
        }
    }
    rc = key_finish(ctx, pk, vk_xy != nullptr);
    if (rc) return fail(rc);
    *out = pk;
    return ZKB_OK;
}

// ProverKey + VerifierKey files of the reference CLI (bin/src/main.rs:274-281 reads them before every proof).
// Public-input rows are recovered from vk.pi_roots (= omega^row, setup.rs:123).
int zkb_plonk_load_keys(zkb_ctx *ctx, const char *pk_path, const char *vk_path, size_t table_size, zkb_plonk_pk **out) {
    if (!ctx || !out) return ZKB_ERR_INVALID;
    *out = nullptr;
    size_t n = 0, n_roots = 0;
    uint64_t vk_xy[10 * AFF_W];
    int vk_inf[10];
    if (zkb_vk_file_read(vk_path, &n, nullptr, 0, &n_roots, vk_xy, vk_inf) != ZKB_OK)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: not a VerifierKey file (ark-serialize unchecked, KZG10 / Bn254)");
    if (n < 8 || (n & (n - 1)) || n > ((size_t)1 << 26)) ZKB_FAIL(ctx, ZKB_ERR_DOMAIN, "zkb_plonk_load_keys: vk.n is not a supported domain size");
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < n) ++log_n;
    std::vector<Fe> roots(n_roots ? n_roots : 1);
    if (zkb_vk_file_read(vk_path, &n, (uint64_t *)roots.data(), n_roots, &n_roots, vk_xy, vk_inf) != ZKB_OK)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: cannot re-read the VerifierKey file");
    std::vector<size_t> pos(n_roots);
    {
        std::unordered_map<Fe, size_t, KeyHash, KeyEq> want;
        for (size_t i = 0; i < n_roots; ++i) want.emplace(roots[i], i);
        if (want.size() != n_roots) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: repeated public-input root");
        const Fe w = host::fr_root_of_unity(log_n);
        Fe x = FR_ONE();
        size_t found = 0;
        for (size_t i = 0; i < n && found < n_roots; ++i) {
            auto it = want.find(x);
            if (it != want.end()) { pos[it->second] = i; ++found; }
            x = host::mul(x, w, host::FR);
        }
        if (found != n_roots) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: a public-input root is not in the domain of size vk.n");
        for (size_t i = 1; i < n_roots; ++i)
            if (pos[i] <= pos[i - 1]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: public-input roots are not in row order");
    }
    size_t lens[10];
    if (zkb_pk_file_info(pk_path, lens) != ZKB_OK)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: not a ProverKey file (ark-serialize unchecked, labels of setup.rs:93-102)");
    std::vector<std::vector<uint64_t>> store(10);
    uint64_t *ptrs[10];
    for (int k = 0; k < 10; ++k) {
        if (lens[k] > n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: a ProverKey polynomial has more than vk.n coefficients");
        store[k].resize(4 * (lens[k] ? lens[k] : 1));
        ptrs[k] = store[k].data();
    }
    if (zkb_pk_file_read(pk_path, ptrs, lens, lens) != ZKB_OK)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_load_keys: a ProverKey coefficient is not a canonical Fr element");
    return zkb_plonk_pk_from_polys(ctx, log_n, ptrs, lens, table_size, pos.data(), n_roots, vk_xy, vk_inf, out);
}

// What `compile` writes (main.rs:106-112), from a key built here: pk and vk files the reference's CLI can read.
int zkb_plonk_save_keys(zkb_ctx *ctx, const zkb_plonk_pk *pk, const char *pk_path, const char *vk_path) {
    if (!ctx || !pk) return ZKB_ERR_INVALID;
    if (pk_path) {
        const int file_to_key[10] = {P_QM, P_QL, P_QR, P_QO, P_QC, P_S1, P_S2, P_S3, P_QLK, P_QT};
        std::vector<std::vector<uint64_t>> store(10);
        const uint64_t *ptrs[10];
        size_t lens[10];
        for (int f = 0; f < 10; ++f) {
            const DPoly &p = pk->poly[file_to_key[f]];
            store[f].resize(4 * (p.len ? p.len : 1));
            lens[f] = p.len;
            ptrs[f] = store[f].data();
            if (p.len && cudaMemcpyAsync(store[f].data(), p.d, p.len * 32, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_save_keys: D2H copy failed");
        }
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_save_keys: stream error");
        if (zkb_pk_file_write(pk_path, ptrs, lens) != ZKB_OK) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_save_keys: cannot write the ProverKey file");
    }
    if (vk_path) {
        const Fe w = host::fr_root_of_unity(pk->log_n);
        std::vector<Fe> roots(pk->pi_pos.size() ? pk->pi_pos.size() : 1);
        for (size_t i = 0; i < pk->pi_pos.size(); ++i) roots[i] = host::pow_u64(w, pk->pi_pos[i], host::FR);
        uint64_t xy[10 * AFF_W];
        int inf[10];
        zkb_plonk_vk_commitments(pk, xy, inf);
        if (zkb_vk_file_write(vk_path, pk->n, (const uint64_t *)roots.data(), pk->pi_pos.size(), xy, inf) != ZKB_OK)
            ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_save_keys: cannot write the VerifierKey file");
    }
    return ZKB_OK;
}

int zkb_plonk_vk_commitments(const zkb_plonk_pk *pk, uint64_t out_xy[80], int is_inf[10]) {
    if (!pk || !out_xy) return ZKB_ERR_INVALID;
    for (int k = 0; k < 10; ++k) {
        memcpy(out_xy + AFF_W * k, pk->vk[k].x.l, FQB);
        memcpy(out_xy + AFF_W * k + AFF_W / 2, pk->vk[k].y.l, FQB);
        if (pk->vk[k].inf) memset(out_xy + AFF_W * k, 0, 2 * FQB);
        if (is_inf) is_inf[k] = pk->vk[k].inf;
    }
    return ZKB_OK;
}

}  // extern "C"

// Both entry points: the wires come either as three host vectors (a, b, c) or, after zkb_plonk_pk_set_wiring, as the
// variable assignment `vars` (n_vars elements) from which the device gathers them.
static int prove_impl(zkb_ctx *ctx, const zkb_plonk_pk *pk, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                      const uint64_t *vars, size_t n_vars, const uint64_t *table, size_t table_len, const uint64_t *pi_values,
                      const uint64_t *blinders, uint8_t proof_out[802], float timings_ms[8]) {
    if (!ctx || !pk) return ZKB_ERR_INVALID;
    if ((!vars && (!a || !b || !c)) || (!table && table_len) || (!pi_values && !pk->pi_pos.empty()) || !blinders || !proof_out)
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove: null argument");
    if (vars) {
        if (!pk->d_wiring[0]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove_vars: the key has no wiring (zkb_plonk_pk_set_wiring)");
        if (n_vars < pk->n_vars_max) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove_vars: the wiring refers to variables beyond n_vars");
        const uint64_t *z = vars;
        if (z[0] | z[1] | z[2] | z[3]) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove_vars: vars[0] is Variable::Zero and must be zero");
    }
    if (table_len > pk->table_size) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove: table size exceeds max size (lookup/table.rs:57)");
    const unsigned log_n = pk->log_n;
    const size_t n = pk->n, n4 = 4 * n, cap = n + 8, n_pi = pk->pi_pos.size();
    const_cast<zkb_plonk_pk *>(pk)->arena_off = 0;
    zkb_commit_abort(ctx);                                                 // a previous call may have failed between push and finish
    cudaStream_t s = ctx->stream;
    double t_mark = now_ms(), t_start = t_mark;
    auto tick = [&](int slot) {
        if (!timings_ms) return;
        cudaStreamSynchronize(s);
        double t = now_ms();
        timings_ms[slot] = (float)(t - t_mark);
        t_mark = t;
    };
    const Fe *bl = (const Fe *)blinders;
    int bl_used = 0;

    Transcript tr("ZKT Plonk", pk->transcript_kind);                       // plonk.rs:107
    tr.append_u64("circuit_size", (uint64_t)n);                           // VerifierKey::seed_transcript (keys/mod.rs:260-275)
    {
        const char *labels[10] = {"q_m_commit", "q_l_commit", "q_r_commit", "q_o_commit", "q_c_commit", "sigma1_commit",
                                  "sigma2_commit", "sigma3_commit", "q_lookup_commit", "q_table_commit"};
        for (int k = 0; k < 10; ++k) tr.append_commitment(labels[k], pk->vk[k]);
    }
    tr.append_scalars("pi", (const Fe *)pi_values, n_pi);                  // prove.rs:110

    // ---- round 2's host side (prove.rs:145-167) depends only on the witness and the key, not on any challenge: a
    // worker thread builds t, f = q_lookup * c and combine_split(t, f) in pinned memory while the GPU runs round 1
    // and uploads the four columns on its own stream
    Fe *t_vals = pk->stage, *f_vals = pk->stage + n, *h1_vals = pk->stage + 2 * n, *h2_vals = pk->stage + 3 * n;
    uint64_t *ev_a, *ev_b, *ev_c, *ev_t, *ev_f, *ev_h1, *ev_h2;
    TAKE(ev_a, n); TAKE(ev_b, n); TAKE(ev_c, n);
    TAKE(ev_t, n); TAKE(ev_f, n); TAKE(ev_h1, n); TAKE(ev_h2, n);
    // Dense lookups: one hash probe per lookup row on one host thread would dominate the proof, so when more than n / 8 rows are
    // lookup gates (or the caller says so: zkb_plonk_pk_set_lookup_mode) t, f, h1, h2 are built on the device instead (lookup.cu).
    const bool dev_lookup = pk->lookup_mode == 2 || (pk->lookup_mode == 0 && pk->lookup_rows.size() > n / 8);
    std::atomic<int> lookup_status{0};
    std::thread lookup_worker([&]() {
        if (dev_lookup) return;
        zkb_plonk_pk *mpk = const_cast<zkb_plonk_pk *>(pk);               // staging bookkeeping of the key object
        if (mpk->dirty_t > table_len) memset(t_vals + table_len, 0, (mpk->dirty_t - table_len) * sizeof(Fe));
        if (table_len) memcpy(t_vals, table, table_len * 32);           // LookupTable::into_multiset: entries then zeros
        mpk->dirty_t = table_len;
        const Fe one = FR_ONE();
        const Fe *cv = (const Fe *)c, *vv = (const Fe *)vars;
        const size_t n_rows = pk->lookup_rows.size();
        for (size_t r = 0; r < n_rows; ++r) {                           // f is zero outside the lookup gates: kept as a compact list
            const Fe &ci = vv ? vv[pk->lookup_wo[r]] : cv[pk->lookup_rows[r]];   // c on that row: given, or value_of_var(w_o[row])
            const Fe &q = pk->lookup_q[r];
            f_vals[r] = feq(q, one) ? ci : fmul(q, ci);
        }
        size_t n_h1 = 0, n_h2 = 0;
        if (!combine_split_sparse((const Fe *)table, table_len, n, f_vals, n_rows, h1_vals, h2_vals, mpk->dirty_h, &n_h1, &n_h2)) {
            lookup_status.store(1);
            return;
        }
        if (n_h1 != n || n_h2 != n) { lookup_status.store(2); return; }
        // pinned -> HBM while the main stream runs round 1.  The four columns are zero almost everywhere, so they are
        // cleared in HBM and only their non-zero regions cross PCIe (t: the table; h1, h2: prefix and suffix around
        // the zero bucket; f: the compact list, scattered to the lookup rows by a kernel)
        cudaStream_t cs = pk->lookup_stream;
        bool ok = cudaSetDevice(ctx->device) == cudaSuccess;
        auto clear = [&](uint64_t *d) { ok = ok && cudaMemsetAsync(d, 0, n * 32, cs) == cudaSuccess; };
        auto put = [&](uint64_t *d, const Fe *h, size_t lo, size_t hi) {
            if (hi > lo) ok = ok && cudaMemcpyAsync(d + 4 * lo, h + lo, (hi - lo) * 32, cudaMemcpyHostToDevice, cs) == cudaSuccess;
        };
        clear(ev_t); clear(ev_f); clear(ev_h1); clear(ev_h2);
        put(ev_t, t_vals, 0, table_len);
        put(ev_h1, h1_vals, 0, mpk->dirty_h[0][0]); put(ev_h1, h1_vals, mpk->dirty_h[0][1], n);
        put(ev_h2, h2_vals, 0, mpk->dirty_h[1][0]); put(ev_h2, h2_vals, mpk->dirty_h[1][1], n);
        if (n_rows) {
            ok = ok && cudaMemcpyAsync(pk->d_lookup_vals, f_vals, n_rows * 32, cudaMemcpyHostToDevice, cs) == cudaSuccess;
            if (ok) scatter_fe_rows_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, cs>>>((uint4 *)ev_f, pk->d_lookup_rows,
                                                                                          (const uint4 *)pk->d_lookup_vals, n_rows);
            ok = ok && cudaGetLastError() == cudaSuccess;
        }
        ok = ok && cudaEventRecord(pk->lookup_uploaded, cs) == cudaSuccess;
        if (!ok) lookup_status.store(3);
    });
    struct Joiner {                                                     // never leave the scope with a joinable thread
        std::thread &t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } lookup_joiner{lookup_worker};

    // ---- round 1: wires (prove.rs:116-140).  The caller's vectors are pageable: a few host threads copy each wire into
    // pinned staging, the DMA runs on the copy stream, and the uploads are interleaved with the commitments: the GPU
    // commits to wire a on the main stream while wire b is staged and crosses PCIe.
    // evals (device) -> blinded coefficient polynomial
    auto blinded_from_dev_evals = [&](const uint64_t *evals, int k_blind, DPoly *out) -> int {
        uint64_t *d;
        TAKE(d, cap);
        ZKB_CUDA(ctx, cudaMemsetAsync(d + 4 * n, 0, (cap - n) * 32, s));
        ZKB_CUDA(ctx, cudaMemcpyAsync(d, evals, n * 32, cudaMemcpyDeviceToDevice, s));
        TRY(zkb_ntt_dev(ctx, d, n, log_n, 1, 0));
        out->d = d; out->cap = cap;
        TRY(zkb_poly_effective_len_dev(ctx, d, n, &out->len));
        if (k_blind) {
            TRY(zkb_poly_add_blinders_dev(ctx, d, out->len, (const uint64_t *)(bl + bl_used), (size_t)k_blind));
            out->len += k_blind;
            bl_used += k_blind;
        }
        return ZKB_OK;
    };
    // several evaluation vectors at once: one batched iNTT launch per pass (a lone 2^18 transform is 0.3 waves of CTAs)
    auto blinded_from_dev_evals_many = [&](const uint64_t *const *evals, const int *k_blind, DPoly *const *out, int count) -> int {
        uint64_t *bufs[8];
        for (int k = 0; k < count; ++k) {
            TAKE(bufs[k], cap);
            ZKB_CUDA(ctx, cudaMemsetAsync(bufs[k] + 4 * n, 0, (cap - n) * 32, s));
            if (evals[k] != bufs[k]) ZKB_CUDA(ctx, cudaMemcpyAsync(bufs[k], evals[k], n * 32, cudaMemcpyDeviceToDevice, s));
        }
        TRY(zkb_ntt_batch_dev(ctx, bufs, (size_t)count, n, log_n, 1, 0));
        for (int k = 0; k < count; ++k) {
            out[k]->d = bufs[k]; out[k]->cap = cap;
            TRY(zkb_poly_effective_len_dev(ctx, bufs[k], n, &out[k]->len));
            if (k_blind[k]) {
                TRY(zkb_poly_add_blinders_dev(ctx, bufs[k], out[k]->len, (const uint64_t *)(bl + bl_used), (size_t)k_blind[k]));
                out[k]->len += k_blind[k];
                bl_used += k_blind[k];
            }
        }
        return ZKB_OK;
    };
    DPoly pa, pb, pc, pt, ph1, ph2, ppi;
    struct { const uint64_t *host; uint64_t *dev; DPoly *poly; } wires[3] = {{a, ev_a, &pa}, {b, ev_b, &pb}, {c, ev_c, &pc}};
    // Several GPUs (SPMD): every rank needs the whole wire for its NTTs, but the ranks of one box share the host's
    // memory bandwidth, so a rank uploads only its 1/world slice and the slices are all-gathered over NVLink.
    const bool split = ctx->world > 1 && ctx->comm && n % (size_t)ctx->world == 0;
    const size_t chunk = split ? n / (size_t)ctx->world : n, first = split ? (size_t)ctx->rank * chunk : 0;
    // the variable assignment instead of the wires: it crosses PCIe once (n_vars elements instead of 3n), every rank its
    // 1/world slice, then the three wires are gathered in HBM (ProvingComposer::wire_evals, prove.rs:49-55)
    size_t vchunk = 0, vfirst = 0, vcount = 0;
    if (vars) {
        zkb_plonk_pk *mpk = const_cast<zkb_plonk_pk *>(pk);
        const size_t world = split ? (size_t)ctx->world : 1, vpad = (n_vars + world - 1) / world * world;
        if (mpk->d_vars_cap < vpad) {
            if (mpk->d_vars) { ZKB_CUDA(ctx, cudaStreamSynchronize(pk->copy_stream)); cudaFree(mpk->d_vars); mpk->d_vars = nullptr; }
            ZKB_CUDA(ctx, cudaMalloc((void **)&mpk->d_vars, vpad * 32));
            mpk->d_vars_cap = vpad;
        }
        vchunk = vpad / world;
        vfirst = split ? (size_t)ctx->rank * vchunk : 0;
        vcount = vfirst < n_vars ? std::min(vchunk, n_vars - vfirst) : 0;
    }
    // All uploads are issued up front by one thread (staging through pinned memory, DMA and -- on several GPUs -- the
    // all-gather on the copy stream), so wire b crosses PCIe while the proving thread is still enqueueing the commitment to
    // wire a: issued from the proving thread between the commitments, the copies left the GPU idle for 0.35 ms per wire at
    // n = 2^18 (profiles/r02t_trace_2^18.jsonl).  up_recorded = number of wires whose event has been recorded.
    std::atomic<int> up_recorded{0}, up_failed{0};
    std::mutex up_mutex;
    std::condition_variable up_cv;                                      // the proving thread sleeps until a wire's event exists:
    auto up_publish = [&](int wires_recorded) {                         // several ranks share one host, a spinning thread per rank starves the stagers
        { std::lock_guard<std::mutex> lock(up_mutex); up_recorded.store(wires_recorded, std::memory_order_release); }
        up_cv.notify_all();
    };
    std::thread uploader([&]() {
        bool ok = cudaSetDevice(ctx->device) == cudaSuccess;
        cudaStream_t cs = pk->copy_stream;
        if (vars) {
            uint64_t *dst = pk->d_vars + 4 * vfirst;
            const uint64_t *src = vars + 4 * vfirst;
            if (ok && vcount) {
                if (vcount <= 3 * n) ok = staged_upload(ctx->device, dst, src, pk->wire_stage, vcount * 32, cs);
                else ok = cudaMemcpyAsync(dst, src, vcount * 32, cudaMemcpyHostToDevice, cs) == cudaSuccess;
            }
            if (ok && split) ok = zkb_comm_allgather_dev(ctx, pk->d_vars, vchunk * 32, cs) == ZKB_OK;
            for (int k = 0; k < 3 && ok; ++k)
                gather_wire_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>((uint4 *)wires[k].dev, pk->d_wiring[k], (const uint4 *)pk->d_vars, n);
            ok = ok && cudaGetLastError() == cudaSuccess && cudaEventRecord(pk->wire_ev[2], cs) == cudaSuccess;
            if (!ok) up_failed.store(1);
            up_publish(3);
            return;
        }
        for (int k = 0; k < 3; ++k) {
            ok = ok && staged_upload(ctx->device, wires[k].dev + 4 * first, wires[k].host + 4 * first, pk->wire_stage + (size_t)k * n + first,
                                     chunk * 32, cs);
            if (ok && split) ok = zkb_comm_allgather_dev(ctx, wires[k].dev, chunk * 32, cs) == ZKB_OK;
            ok = ok && cudaEventRecord(pk->wire_ev[k], cs) == cudaSuccess;
            if (!ok) up_failed.store(1);
            up_publish(k + 1);
        }
    });
    Joiner uploader_joiner{uploader};
    // a, b, c are one batch upstream (prove.rs:133-135), but here they are pushed as the wires cross PCIe: every push is
    // cut over ALL ranks (no zkb_commit_expect).  Fanning this batch out -- one group of ranks per wire -- serialises it
    // behind the uploads: the last group starts its (larger) MSM only when the last wire has arrived (measured on 8 B200:
    // round 1 4.6 instead of 3.5 ms at n = 2^20, profiles/r02i_bench_n8.json).
    for (int k = 0; k < 3; ++k) {
        {
            std::unique_lock<std::mutex> lock(up_mutex);
            up_cv.wait(lock, [&]() { return up_recorded.load(std::memory_order_acquire) > k; });
        }
        if (up_failed.load()) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_prove: upload of the wires failed");
        ZKB_CUDA(ctx, cudaStreamWaitEvent(s, pk->wire_ev[vars ? 2 : k], 0));
        if (k == 0) tick(0);
        TRY(blinded_from_dev_evals(wires[k].dev, 2, wires[k].poly));
        TRY(zkb_commit_push(ctx, wires[k].poly->d, 0, wires[k].poly->len));
    }
    uploader.join();
    tick(1);

    // ---- round 2: lookup multisets on the host (prove.rs:145-167).  t, h1, h2 depend on the witness and the table only, and
    // their commitments enter the transcript before any challenge is drawn (prove.rs:183-199), so they join the batch of
    // a, b, c: six pipelined commitments, one exposed reduction tail and one exchange between the ranks instead of two.
    lookup_worker.join();                                               // started before round 1 (see above)
    if (lookup_status.load() == 1) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "ElementNotIndexedInTable (lookup/multiset.rs:121)");
    if (lookup_status.load() == 2) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove: combine_split halves are not n long");
    if (lookup_status.load() == 3) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "zkb_plonk_prove: upload of the lookup multisets failed");
    if (dev_lookup) {                                                   // the wire c is in HBM by now: f, the bucket counts and the halves there
        *pk->lookup_status = 0;
        TRY(zkb_lookup_multisets_dev(ctx, log_n, table, table_len, pk->d_qlookup_evals, ev_c, ev_t, ev_f, ev_h1, ev_h2, pk->lookup_status));
    } else {
        ZKB_CUDA(ctx, cudaStreamWaitEvent(s, pk->lookup_uploaded, 0));
    }
    tick(2);
    {
        const uint64_t *evs[3] = {ev_t, ev_h1, ev_h2};
        const int kb[3] = {0, 3, 2};
        DPoly *outs[3] = {&pt, &ph1, &ph2};
        TRY(blinded_from_dev_evals_many(evs, kb, outs, 3));
        for (int k = 0; k < 3; ++k) TRY(zkb_commit_push(ctx, outs[k]->d, 0, outs[k]->len));
    }
    // ---- round 4's challenge-independent part, enqueued behind the six commitments: the public-input polynomial and the
    // coset NTTs of a, b, c, pi, t, h1, h2 run on the main stream while the reductions of the last commitments finish on the
    // tail stream and the host folds and exchanges the results (the GPU was idle there)
    std::vector<unsigned long long> pi_pos_host(pk->pi_pos.begin(), pk->pi_pos.end());
    {
        for (size_t k = 0; k < n_pi; ++k)
            if (pk->pi_pos[k] >= n) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove: public input position out of range");
        uint64_t *d, *d_vals = nullptr, *d_pos = nullptr;                  // PublicInputs::as_evals (pi.rs:75-82), built in HBM
        TAKE(d, n);
        ZKB_CUDA(ctx, cudaMemsetAsync(d, 0, n * 32, s));
        if (n_pi) {
            TAKE(d_vals, n_pi); TAKE(d_pos, (n_pi + 3) / 4);
            ZKB_CUDA(ctx, cudaMemcpyAsync(d_vals, pi_values, n_pi * 32, cudaMemcpyHostToDevice, s));
            ZKB_CUDA(ctx, cudaMemcpyAsync(d_pos, pi_pos_host.data(), n_pi * 8, cudaMemcpyHostToDevice, s));
            scatter_fe_kernel<<<(unsigned)((n_pi + 127) / 128), 128, 0, s>>>((uint4 *)d, (const unsigned long long *)d_pos, (const uint4 *)d_vals, n_pi);
            ZKB_CUDA(ctx, cudaGetLastError());
        }
        TRY(zkb_ntt_dev(ctx, d, n, log_n, 1, 0));
        ppi.d = d; ppi.cap = n;
        TRY(zkb_poly_effective_len_dev(ctx, d, n, &ppi.len));
    }
    // Several GPUs (SPMD): the nine coset NTTs are independent, so rank k % world transforms polynomial k, every
    // rank receives the slice (+ 4 halo elements for the "next row" reads) its part of the quotient needs, evaluates
    // that part, and the quotient slices are all-gathered before the (replicated) coset iNTT.
    const size_t world = (size_t)ctx->world;
    const bool fan = world > 1 && ctx->comm && n4 % world == 0 && n4 / world >= 4;
    DPoly pz1, pz2;
    const DPoly *wit_p[9] = {&pz1, &pz2, &pa, &pb, &pc, &ppi, &pt, &ph1, &ph2};
    uint64_t *wit_w[9];
    for (int k = 0; k < 9; ++k) TAKE(wit_w[k], n4);
    auto coset_ntts = [&](int k_lo, int k_hi) -> int {                  // polynomials wit_p[k_lo .. k_hi) -> wit_w, this rank's share
        uint64_t *own[9];
        size_t n_own = 0, max_len = 0;
        for (int k = k_lo; k < k_hi; ++k) {
            if (fan && (size_t)k % world != (size_t)ctx->rank) continue;
            uint64_t *d = wit_w[k];
            ZKB_CUDA(ctx, cudaMemsetAsync(d, 0, n4 * 32, s));
            ZKB_CUDA(ctx, cudaMemcpyAsync(d, wit_p[k]->d, wit_p[k]->len * 32, cudaMemcpyDeviceToDevice, s));
            own[n_own++] = d;
            max_len = wit_p[k]->len > max_len ? wit_p[k]->len : max_len;
        }
        if (!n_own) return ZKB_OK;
        return zkb_ntt_batch_dev(ctx, own, n_own, max_len, log_n + 2, 0, 1);   // the buffers are zero beyond their own lengths
    };
    TRY(coset_ntts(2, 9));
    Pt c_a[3], c_t[3];
    {
        Pt six[6];
        TRY(commit_finish_pts(ctx, 6, six));
        if (dev_lookup) {                                               // the status word crossed PCIe before the commitments finished
            ZKB_CUDA(ctx, cudaStreamSynchronize(s));
            if (*pk->lookup_status & 1) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "ElementNotIndexedInTable (lookup/multiset.rs:121)");
            if (*pk->lookup_status & 2) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove: combine_split halves are not n long");
        }
        for (int k = 0; k < 3; ++k) { c_a[k] = six[k]; c_t[k] = six[3 + k]; }
    }
    tr.append_commitment("a_commit", c_a[0]);
    tr.append_commitment("b_commit", c_a[1]);
    tr.append_commitment("c_commit", c_a[2]);
    tr.append_commitment("t_commit", c_t[0]);
    tr.append_commitment("h1_commit", c_t[1]);
    tr.append_commitment("h2_commit", c_t[2]);
    const Fe beta = tr.challenge_scalar("beta"), gamma = tr.challenge_scalar("gamma");
    const Fe delta = tr.challenge_scalar("delta"), epsilon = tr.challenge_scalar("epsilon");
    if (feq(beta, gamma) || feq(beta, delta) || feq(beta, epsilon) || feq(gamma, delta) || feq(gamma, epsilon) || feq(delta, epsilon))
        ZKB_FAIL(ctx, ZKB_ERR_INVALID, "challenges must be different (prove.rs:202-207)");
    tick(3);

    // ---- round 3: grand products (prove.rs:209-251)
    {
        uint64_t *z1e, *z2e;
        TAKE(z1e, n); TAKE(z2e, n);
        TRY(zkb_z1_evals_dev(ctx, log_n, beta.l, gamma.l, ev_a, ev_b, ev_c, pk->sigma_evals[0], pk->sigma_evals[1], pk->sigma_evals[2], z1e));
        if (zkb_grand_product_failed(ctx)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "compute_z1_poly: zero denominator (permutation/mod.rs:242)");
        TRY(zkb_z2_evals_dev(ctx, log_n, delta.l, epsilon.l, ev_f, ev_t, ev_h1, ev_h2, z2e));
        if (zkb_grand_product_failed(ctx)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "compute_z2_poly: zero denominator (lookup/mod.rs:73)");
        const uint64_t *evs[2] = {z1e, z2e};
        const int kb[2] = {3, 3};
        DPoly *outs[2] = {&pz1, &pz2};
        TRY(blinded_from_dev_evals_many(evs, kb, outs, 2));
    }
    // the coset NTTs of z1 and z2 need the polynomials, not alpha: enqueued behind the two commitments, they run while the
    // last reduction finishes on the tail stream and the host folds the results
    Pt c_z[2];
    TRY(zkb_commit_expect(ctx, 2));
    TRY(zkb_commit_push(ctx, pz1.d, 0, pz1.len));
    TRY(zkb_commit_push(ctx, pz2.d, 0, pz2.len));
    TRY(coset_ntts(0, 2));
    TRY(commit_finish_pts(ctx, 2, c_z));
    tr.append_commitment("z1_commit", c_z[0]);
    tr.append_commitment("z2_commit", c_z[1]);
    tick(4);

    // ---- round 4: quotient (prove.rs:258-308, quotient_poly.rs:20-227); its nine coset NTTs ran above
    const Fe alpha = tr.challenge_scalar("alpha");
    uint64_t *q_buf;
    {
        const uint64_t *wit[9];
        for (int k = 0; k < 9; ++k) wit[k] = wit_w[k];
        uint64_t ch[20];
        const Fe *cs[5] = {&alpha, &beta, &gamma, &delta, &epsilon};
        for (int k = 0; k < 5; ++k) memcpy(ch + 4 * k, cs[k]->l, 32);
        TAKE(q_buf, n4);
        if (fan) {
            const size_t chunk4 = n4 / world, lo = (size_t)ctx->rank * chunk4;
            TRY(zkb_comm_spread_slices(ctx, wit_w, 9, n4, 4, s));
            TRY(zkb_quotient_evals_range_dev(ctx, log_n, ch, wit, (const uint64_t *const *)pk->epk, q_buf, lo, lo + chunk4));
            TRY(zkb_comm_allgather_dev(ctx, q_buf, chunk4 * 32, s));
        } else {
            TRY(zkb_quotient_evals_dev(ctx, log_n, ch, wit, (const uint64_t *const *)pk->epk, q_buf));
        }
        TRY(zkb_ntt_dev(ctx, q_buf, n4, log_n + 2, 1, 1));
    }
    size_t q_len = 0;
    TRY(zkb_poly_effective_len_dev(ctx, q_buf, n4, &q_len));
    if (q_len < 2 * (n + 2)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "quotient shorter than 2(n+2): the reference's slice would panic (prove.rs:287-292)");
    // a satisfied circuit whose blinded polynomials all had n coefficients gives deg q <= 3n + 5.  Anything longer means the
    // division by Z_H was not exact: an unsatisfied witness, or a polynomial shorter than n before add_blinders_to_poly
    // (prove.rs:472-483 appends at the CURRENT length, which only equals + b(X)(X^n - 1) at full length -- e.g. z2 = 1 for a
    // circuit without any lookup).  The reference then fails in PC::commit (q_hi exceeds the committer key's degree).
    if (q_len - 2 * (n + 2) > cap) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "quotient longer than 3n + 6 coefficients: the division by Z_H was not exact (prove.rs:287-292; PC::commit would fail upstream)");
    DPoly q_part[3];
    {
        const size_t lo[3] = {0, n + 2, 2 * (n + 2)}, hi[3] = {n + 2, 2 * (n + 2), q_len};
        for (int k = 0; k < 3; ++k) {
            uint64_t *d;
            TAKE(d, cap);
            ZKB_CUDA(ctx, cudaMemsetAsync(d, 0, cap * 32, s));
            ZKB_CUDA(ctx, cudaMemcpyAsync(d, q_buf + 4 * lo[k], (hi[k] - lo[k]) * 32, cudaMemcpyDeviceToDevice, s));
            q_part[k].d = d; q_part[k].cap = cap;
            TRY(zkb_poly_effective_len_dev(ctx, d, hi[k] - lo[k], &q_part[k].len));
        }
        // q_lo.push(b0); q_mid[0] -= b0; q_mid.push(b1); q_hi[0] -= b1   (prove.rs:296-300)
        const Fe b0 = bl[bl_used], b1 = bl[bl_used + 1];
        bl_used += 2;
        Fe m0, h0;
        ZKB_CUDA(ctx, cudaMemcpyAsync(m0.l, q_part[1].d, 32, cudaMemcpyDeviceToHost, s));
        ZKB_CUDA(ctx, cudaMemcpyAsync(h0.l, q_part[2].d, 32, cudaMemcpyDeviceToHost, s));
        ZKB_CUDA(ctx, cudaStreamSynchronize(s));
        m0 = fsub(m0, b0);
        h0 = fsub(h0, b1);
        ZKB_CUDA(ctx, cudaMemcpyAsync(q_part[0].d + 4 * q_part[0].len, b0.l, 32, cudaMemcpyHostToDevice, s));
        ZKB_CUDA(ctx, cudaMemcpyAsync(q_part[1].d, m0.l, 32, cudaMemcpyHostToDevice, s));
        ZKB_CUDA(ctx, cudaMemcpyAsync(q_part[1].d + 4 * q_part[1].len, b1.l, 32, cudaMemcpyHostToDevice, s));
        ZKB_CUDA(ctx, cudaMemcpyAsync(q_part[2].d, h0.l, 32, cudaMemcpyHostToDevice, s));
        ZKB_CUDA(ctx, cudaStreamSynchronize(s));                        // the host temporaries above must outlive the copies
        q_part[0].len += 1;
        q_part[1].len += 1;
        if (q_part[2].len == 0) q_part[2].len = 1;                      // coeffs[0] -= b1 on an (improbable) empty q_hi would panic upstream
    }
    Pt c_q[3];
    { const DPoly *ps[3] = {&q_part[0], &q_part[1], &q_part[2]}; TRY(commit_many(ctx, ps, 3, c_q)); }
    tr.append_commitment("q_lo_commit", c_q[0]);
    tr.append_commitment("q_mid_commit", c_q[1]);
    tr.append_commitment("q_hi_commit", c_q[2]);
    const Fe xi = tr.challenge_scalar("xi");
    tick(5);

    // ---- round 5: linearisation (linearization_poly.rs:19-121) and the two openings (prove.rs:381-451)
    const Fe one = FR_ONE();
    const Fe w_n = host::fr_root_of_unity(log_n);
    const Fe shifted = fmul(xi, w_n);
    const Fe zh = fsub(host::pow_u64(xi, (uint64_t)n, host::FR), one);
    const Fe l1 = fmul(zh, host::inv(fmul(host::from_u64((uint64_t)n, host::FR), fsub(xi, one)), host::FR));
    Fe ev[12];                                                          // a b c sigma1 sigma2 z1_next q_lookup t t_next z2_next h1_next h2
    {
        struct { const DPoly *p; const Fe *at; } q[12] = {
            {&pa, &xi}, {&pb, &xi}, {&pc, &xi}, {&pk->poly[P_S1], &xi}, {&pk->poly[P_S2], &xi}, {&pz1, &shifted},
            {&pk->poly[P_QLK], &xi}, {&pt, &xi}, {&pt, &shifted}, {&pz2, &shifted}, {&ph1, &shifted}, {&ph2, &xi}};
        const uint64_t *ptrs[12];
        size_t lens[12];
        Fe at[12];
        for (int k = 0; k < 12; ++k) { ptrs[k] = q[k].p->d; lens[k] = q[k].p->len; at[k] = *q[k].at; }
        TRY(zkb_poly_eval_many_dev(ctx, 12, ptrs, lens, (const uint64_t *)at, (uint64_t *)ev));
    }
    const Fe &ea = ev[0], &eb = ev[1], &ec = ev[2], &es1 = ev[3], &es2 = ev[4], &ez1n = ev[5], &eql = ev[6], &et = ev[7],
             &etn = ev[8], &ez2n = ev[9], &eh1n = ev[10], &eh2 = ev[11];
    const Fe al2 = fmul(alpha, alpha), al3 = fmul(al2, alpha), al4 = fmul(al3, alpha), al5 = fmul(al4, alpha);
    const Fe bxi = fmul(beta, xi), opd = fadd(one, delta), eopd = fmul(epsilon, opd);
    const Fe k1 = host::from_u64(7, host::FR), k2 = host::from_u64(13, host::FR);
    const Fe s_z1 = fadd(fmul(fmul(fmul(alpha, fadd(fadd(bxi, ea), gamma)), fadd(fadd(fmul(bxi, k1), eb), gamma)),
                              fadd(fadd(fmul(bxi, k2), ec), gamma)), fmul(l1, al2));
    const Fe s_sigma3 = fneg(fmul(fmul(fmul(fmul(alpha, beta), ez1n), fadd(fadd(fmul(beta, es1), ea), gamma)),
                                  fadd(fadd(fmul(beta, es2), eb), gamma)));
    const Fe s_z2 = fadd(fmul(fmul(fmul(al3, opd), fadd(epsilon, fmul(eql, ec))), fadd(fadd(eopd, et), fmul(delta, etn))), fmul(al4, l1));
    const Fe s_h1 = fneg(fmul(fmul(al3, ez2n), fadd(fadd(eopd, eh2), fmul(delta, eh1n))));
    const Fe s_qt = fmul(al5, et);
    const Fe xn2 = fmul(fmul(fadd(zh, one), xi), xi);                    // xi^(n+2)
    const Fe nzh = fneg(zh);
    DPoly r_poly;
    {
        const DPoly *terms[13] = {&pk->poly[P_QM], &pk->poly[P_QL], &pk->poly[P_QR], &pk->poly[P_QO], &pk->poly[P_QC], &pz1,
                                  &pk->poly[P_S3], &pz2, &ph1, &pk->poly[P_QT], &q_part[0], &q_part[1], &q_part[2]};
        const Fe sc[13] = {fmul(ea, eb), ea, eb, ec, one, s_z1, s_sigma3, s_z2, s_h1, s_qt, nzh, fmul(nzh, xn2), fmul(fmul(nzh, xn2), xn2)};
        const uint64_t *ptrs[13];
        size_t lens[13], m = 0;
        for (int k = 0; k < 13; ++k) { ptrs[k] = terms[k]->d; lens[k] = terms[k]->len; m = lens[k] > m ? lens[k] : m; }
        TAKE(r_poly.d, m + 1);
        TRY(zkb_poly_lincomb_dev(ctx, 13, ptrs, lens, (const uint64_t *)sc, r_poly.d, m));
        TRY(zkb_poly_effective_len_dev(ctx, r_poly.d, m, &r_poly.len));
    }
    {
        const char *labels[12] = {"a_eval", "b_eval", "c_eval", "sigma1_eval", "sigma2_eval", "z1_next_eval", "q_lookup_eval",
                                  "t_eval", "t_next_eval", "z2_next_eval", "h1_next_eval", "h2_eval"};
        for (int k = 0; k < 12; ++k) tr.append_scalar(labels[k], ev[k]);
    }
    const Fe eta = tr.challenge_scalar("eta");
    // SonicKZG10::open: combined = sum eta^i p_i, witness = (combined - combined(z)) / (X - z)
    auto witness_of = [&](const DPoly *const *plist, int count, const Fe &point, DPoly *out) -> int {
        const uint64_t *ptrs[16];
        size_t lens[16], m = 0;
        Fe sc[16], cur = one;
        for (int k = 0; k < count; ++k) {
            ptrs[k] = plist[k]->d; lens[k] = plist[k]->len; m = lens[k] > m ? lens[k] : m;
            sc[k] = cur;
            cur = fmul(cur, eta);
        }
        uint64_t *comb;
        TAKE(comb, m + 1);
        TRY(zkb_poly_lincomb_dev(ctx, (size_t)count, ptrs, lens, (const uint64_t *)sc, comb, m));
        TAKE(out->d, m + 1);
        uint64_t evz[4];
        TRY(zkb_poly_divide_linear_dev(ctx, comb, m, point.l, out->d, evz));
        out->cap = m + 1;
        return zkb_poly_effective_len_dev(ctx, out->d, m ? m - 1 : 0, &out->len);
    };
    DPoly w1, w2;
    {
        const DPoly *l1p[9] = {&r_poly, &pa, &pb, &pc, &pk->poly[P_S1], &pk->poly[P_S2], &pk->poly[P_QLK], &pt, &ph2};
        const DPoly *l2p[4] = {&pz1, &pz2, &pt, &ph1};
        TRY(witness_of(l1p, 9, xi, &w1));
        TRY(witness_of(l2p, 4, shifted, &w2));
    }
    Pt c_w[2];
    { const DPoly *ps[2] = {&w1, &w2}; TRY(commit_many(ctx, ps, 2, c_w)); }
    tick(6);

    // ---- Proof (proof.rs:106-155): 11 compressed commitments, 2 x (w, None), 12 evaluations
    uint8_t *o = proof_out;
    const Pt *cm[11] = {&c_a[0], &c_a[1], &c_a[2], &c_t[0], &c_t[1], &c_t[2], &c_z[0], &c_z[1], &c_q[0], &c_q[1], &c_q[2]};
    for (int k = 0; k < 11; ++k, o += FQB) g1_compressed(*cm[k], o);
    for (int k = 0; k < 2; ++k) { g1_compressed(c_w[k], o); o[FQB] = 0; o += FQB + 1; }
    for (int k = 0; k < 12; ++k, o += 32) fe_bytes(ev[k], host::FR, o);
    if (timings_ms) timings_ms[7] = (float)(now_ms() - t_start);
    return ZKB_OK;
}

extern "C" {

int zkb_plonk_prove(zkb_ctx *ctx, const zkb_plonk_pk *pk, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                    const uint64_t *table, size_t table_len, const uint64_t *pi_values, const uint64_t *blinders,
                    uint8_t proof_out[802], float timings_ms[8]) {
    if (ctx && (!a || !b || !c)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove: null argument");
    return prove_impl(ctx, pk, a, b, c, nullptr, 0, table, table_len, pi_values, blinders, proof_out, timings_ms);
}

// ProvingComposer::wire_evals (prove.rs:49-55) moved to the device: the key keeps the circuit's wire maps,
// a proof uploads the variable assignment once and the three wires are gathered in HBM.
int zkb_plonk_pk_set_wiring(zkb_ctx *ctx, zkb_plonk_pk *pk, const uint32_t *w_l, const uint32_t *w_r, const uint32_t *w_o) {
    if (!ctx || !pk) return ZKB_ERR_INVALID;
    if (!w_l || !w_r || !w_o) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_pk_set_wiring: null wire map");
    const size_t n = pk->n;
    const uint32_t *w[3] = {w_l, w_r, w_o};
    uint32_t vmax = 0;
    for (int k = 0; k < 3; ++k)
        for (size_t i = 0; i < n; ++i) vmax = std::max(vmax, w[k][i]);
    for (int k = 0; k < 3; ++k) {
        if (!pk->d_wiring[k]) {
            uint64_t *d;
            TRY(dev_alloc_owned(ctx, pk, n * 4, &d));
            pk->d_wiring[k] = (uint32_t *)d;
        }
        ZKB_CUDA(ctx, cudaMemcpyAsync(pk->d_wiring[k], w[k], n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pk->lookup_wo.resize(pk->lookup_rows.size());
    for (size_t r = 0; r < pk->lookup_rows.size(); ++r) pk->lookup_wo[r] = w_o[pk->lookup_rows[r]];
    pk->n_vars_max = (size_t)vmax + 1;
    return ZKB_OK;
}

int zkb_plonk_prove_vars(zkb_ctx *ctx, const zkb_plonk_pk *pk, const uint64_t *var_values, size_t n_vars, const uint64_t *table,
                         size_t table_len, const uint64_t *pi_values, const uint64_t *blinders, uint8_t proof_out[802],
                         float timings_ms[8]) {
    if (ctx && (!var_values || !n_vars)) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_plonk_prove_vars: null assignment");
    return prove_impl(ctx, pk, nullptr, nullptr, nullptr, var_values, n_vars, table, table_len, pi_values, blinders, proof_out, timings_ms);
}

}  // extern "C"
