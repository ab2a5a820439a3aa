// zkb200.hpp -- C++ host-side mirror of the reference's generic seams over the C ABI of zkb200.h.
//
// The reference is compiled (Rust) code whose toolchain is absent from this image, so the host side above the C ABI
// is written in C++ with the reference's own names, argument meaning and error behaviour; INTEGRATION.md shows the
// one-to-one Rust newtypes.  Header-only; link with -lzkb200.
//
//   zkb::Fr, zkb::G1Affine        ark-ff Fp256 (4 x u64 LE Montgomery limbs) / ark-ec GroupAffine (x, y, infinity)
//   zkb::GpuDomain                D: EvaluationDomain<F> + EvaluationDomainExt<F>   (plonk-core/src/util.rs:27-140)
//   zkb::GpuKZG10                 PC: HomomorphicCommitment<F> = KZG10<Bn254>        (plonk-core/src/commitment.rs:10-46)
//   zkb::compute_z1_poly          plonk-core/src/permutation/mod.rs:181-257
//   zkb::compute_z2_poly          plonk-core/src/lookup/mod.rs:25-85
//   zkb::quotient_compute         plonk-core/src/proof_system/quotient_poly.rs:20-227
//   zkb::extend_prover_key        plonk-core/src/proof_system/keys/mod.rs:78-146
//   zkb::PlonkKey                 proof_system::setup / prove behind zkb_plonk_setup, zkb_plonk_load_keys, zkb_plonk_prove
//                                 (setup.rs:42-166, prove.rs:59-470; key files of bin/src/parser.rs:5-29)
//   zkb::verify                   Proof::verify (proof_system/proof.rs:285-503), host side
//
// Errors: the reference returns Error::InvalidEvalDomainSize / PC errors or panics on broken invariants; here every
// failure is a zkb::Error exception carrying the zkb_status code and the library's message (nothing crosses the C
// boundary as an exception).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkb200.h"

namespace zkb {

struct Fr {
    uint64_t l[4];                       // Montgomery form, little endian
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fr &o) const { return std::memcmp(l, o.l, 32) == 0; }
};
static_assert(sizeof(Fr) == 32, "Fr must match arkworks' in-memory Fp256");

struct G1Affine {
    uint64_t x[4], y[4];                 // Montgomery Fq; (0, 0) is the point at infinity at this boundary
    bool infinity() const { return (x[0] | x[1] | x[2] | x[3] | y[0] | y[1] | y[2] | y[3]) == 0; }
    bool operator==(const G1Affine &o) const { return std::memcmp(this, &o, 64) == 0; }
};
static_assert(sizeof(G1Affine) == 64, "G1Affine is x || y");

using DensePolynomial = std::vector<Fr>; // coefficients, low degree first; trailing zeros are dropped by truncate()

inline void truncate(DensePolynomial &p) {   // DensePolynomial::from_coefficients_vec
    while (!p.empty() && p.back().is_zero()) p.pop_back();
}

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("zkb200 error " + std::to_string(c) + ": " + m), code(c) {}
};

// One context per GPU and proving thread (prove() is !Send: prove.rs:62).
class Context {
  public:
    explicit Context(int device = 0) {
        zkb_ctx *c = nullptr;
        int rc = zkb_ctx_create(device, &c);
        if (rc != ZKB_OK) throw Error(rc, "zkb_ctx_create failed (no CUDA device; there is no CPU fallback)");
        ctx_.reset(c, zkb_ctx_destroy);
    }
    zkb_ctx *raw() const { return ctx_.get(); }
    void check(int rc) const {
        if (rc != ZKB_OK) throw Error(rc, zkb_last_error(ctx_.get()));
    }
    void sync() const { check(zkb_ctx_sync(raw())); }

  private:
    std::shared_ptr<zkb_ctx> ctx_;
};

// RAII device buffer of Fr / point words
class DeviceBuffer {
  public:
    DeviceBuffer(const Context &ctx, size_t bytes) : ctx_(ctx), bytes_(bytes) {
        ctx_.check(zkb_dev_alloc(ctx_.raw(), bytes, &p_));
    }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    DeviceBuffer(DeviceBuffer &&o) noexcept : ctx_(o.ctx_), p_(o.p_), bytes_(o.bytes_) { o.p_ = nullptr; }
    ~DeviceBuffer() { if (p_) zkb_dev_free(ctx_.raw(), p_); }
    uint64_t *words() const { return static_cast<uint64_t *>(p_); }
    size_t bytes() const { return bytes_; }
    void upload(const void *src, size_t n) const { ctx_.check(zkb_h2d(ctx_.raw(), p_, src, n)); }
    void download(void *dst, size_t n) const { ctx_.check(zkb_d2h(ctx_.raw(), dst, p_, n)); }

  private:
    Context ctx_;
    void *p_ = nullptr;
    size_t bytes_ = 0;
};

inline DeviceBuffer to_device(const Context &ctx, const std::vector<Fr> &v, size_t capacity_elems = 0) {
    size_t cap = capacity_elems > v.size() ? capacity_elems : v.size();
    std::vector<Fr> padded;
    const Fr *src = v.data();
    if (cap > v.size()) {                 // zero-extend (fft_in_place's resize)
        padded.assign(cap, Fr{{0, 0, 0, 0}});
        std::memcpy(padded.data(), v.data(), v.size() * 32);
        src = padded.data();
    }
    DeviceBuffer b(ctx, (cap ? cap : 1) * 32);
    if (cap) b.upload(src, cap * 32);
    return b;
}

// ---------------------------------------------------------------------------------------------------- GpuDomain
class GpuDomain {
  public:
    // EvaluationDomain::new: size = num_coeffs.next_power_of_two(); None above 2^TWO_ADICITY = 2^28
    static std::optional<GpuDomain> create(const Context &ctx, size_t num_coeffs) {
        unsigned log = 0;
        while ((size_t(1) << log) < num_coeffs) ++log;
        if (log > 28) return std::nullopt;   // -> Error::InvalidEvalDomainSize at the call site (prove.rs:77-81)
        return GpuDomain(ctx, log);
    }
    size_t size() const { return size_t(1) << log_size_; }
    unsigned log_size_of_group() const { return log_size_; }          // EvaluationDomainExt (util.rs:27-36)
    const Context &context() const { return ctx_; }

    std::vector<Fr> fft(const std::vector<Fr> &coeffs) const { auto v = coeffs; fft_in_place(v); return v; }
    std::vector<Fr> ifft(const std::vector<Fr> &evals) const { auto v = evals; ifft_in_place(v); return v; }
    std::vector<Fr> coset_fft(const std::vector<Fr> &coeffs) const { auto v = coeffs; coset_fft_in_place(v); return v; }
    std::vector<Fr> coset_ifft(const std::vector<Fr> &evals) const { auto v = evals; coset_ifft_in_place(v); return v; }
    void fft_in_place(std::vector<Fr> &v) const { run(v, 0, 0); }
    void ifft_in_place(std::vector<Fr> &v) const { run(v, 1, 0); }
    void coset_fft_in_place(std::vector<Fr> &v) const { run(v, 0, 1); }
    void coset_ifft_in_place(std::vector<Fr> &v) const { run(v, 1, 1); }
    // device-resident variants (the buffer must hold size() elements)
    void transform_dev(uint64_t *data_dev, size_t len, bool inverse, bool coset) const {
        ctx_.check(zkb_ntt_dev(ctx_.raw(), data_dev, len, log_size_, inverse, coset));
    }
    bool operator==(const GpuDomain &o) const { return log_size_ == o.log_size_; }

  private:
    GpuDomain(const Context &ctx, unsigned log) : ctx_(ctx), log_size_(log) {}
    void run(std::vector<Fr> &v, int inverse, int coset) const {
        size_t len = v.size() < size() ? v.size() : size();
        v.resize(size(), Fr{{0, 0, 0, 0}});                         // what ark-poly's *_in_place do first
        ctx_.check(zkb_ntt(ctx_.raw(), reinterpret_cast<uint64_t *>(v.data()), len, log_size_, inverse, coset));
    }
    Context ctx_;
    unsigned log_size_;
};

// ---------------------------------------------------------------------------------------------------- GpuKZG10
class GpuKZG10 {
  public:
    explicit GpuKZG10(const Context &ctx) : ctx_(ctx) {}
    // PC::trim: keep the committer key's powers_of_g resident; fixed_base builds the window tables once per key
    void trim(const std::vector<G1Affine> &powers_of_g, bool fixed_base = true) {
        ctx_.check(zkb_srs_load_g1(ctx_.raw(), reinterpret_cast<const uint64_t *>(powers_of_g.data()), powers_of_g.size()));
        if (fixed_base) ctx_.check(zkb_srs_precompute(ctx_.raw(), 0));
    }
    size_t supported_degree() const { size_t n = zkb_srs_size(ctx_.raw()); return n ? n - 1 : 0; }

    // PolynomialCommitment::commit(ck, polys, None): one commitment per polynomial (kzg10::commit: skip the leading
    // zero coefficients, into_repr, MSM).  Throws ZKB_ERR_NO_SRS where kzg10 returns TooManyCoefficients.
    std::vector<G1Affine> commit(const std::vector<const DensePolynomial *> &polys) const {
        std::vector<G1Affine> out;
        for (const DensePolynomial *p : polys) out.push_back(commit_one(*p));
        return out;
    }
    G1Affine commit_one(const DensePolynomial &p) const {
        size_t lo = 0, hi = p.size();
        while (hi > 0 && p[hi - 1].is_zero()) --hi;
        while (lo < hi && p[lo].is_zero()) ++lo;                  // skip_leading_zeros_and_convert_to_bigints
        G1Affine out{};
        if (lo == hi) return out;                                   // the zero polynomial commits to the identity
        DeviceBuffer d(ctx_, (hi - lo) * 32);
        d.upload(p.data() + lo, (hi - lo) * 32);
        int inf = 0;
        ctx_.check(zkb_commit_dev(ctx_.raw(), d.words(), lo, hi - lo, reinterpret_cast<uint64_t *>(&out), &inf));
        return out;
    }
    // HomomorphicCommitment::multi_scalar_mul (commitment.rs:31-46); scalars canonical (into_repr)
    G1Affine multi_scalar_mul(const std::vector<G1Affine> &commitments, const std::vector<std::array<uint64_t, 4>> &scalars) const {
        size_t n = commitments.size() < scalars.size() ? commitments.size() : scalars.size();
        G1Affine out{};
        int inf = 0;
        ctx_.check(zkb_msm_g1_bases(ctx_.raw(), reinterpret_cast<const uint64_t *>(commitments.data()),
                                    reinterpret_cast<const uint64_t *>(scalars.data()), n, reinterpret_cast<uint64_t *>(&out), &inf));
        return out;
    }
    const Context &context() const { return ctx_; }

  private:
    Context ctx_;
};

// ---------------------------------------------------------------------------------------------------- free functions
// compute_z1_poly(domain, beta, gamma, a, b, c, sigma1, sigma2, sigma3) -> DensePolynomial
inline DensePolynomial compute_z1_poly(const GpuDomain &domain, const Fr &beta, const Fr &gamma, const std::vector<Fr> &a,
                                       const std::vector<Fr> &b, const std::vector<Fr> &c, const std::vector<Fr> &sigma1,
                                       const std::vector<Fr> &sigma2, const std::vector<Fr> &sigma3) {
    const Context &ctx = domain.context();
    const size_t n = domain.size();
    for (const auto *v : {&a, &b, &c, &sigma1, &sigma2, &sigma3})
        if (v->size() != n) throw Error(ZKB_ERR_INVALID, "compute_z1_poly: assert_eq!(len, n) failed");   // mod.rs:197-202
    DeviceBuffer da = to_device(ctx, a), db = to_device(ctx, b), dc = to_device(ctx, c), d1 = to_device(ctx, sigma1),
                 d2 = to_device(ctx, sigma2), d3 = to_device(ctx, sigma3), out(ctx, n * 32);
    ctx.check(zkb_z1_evals_dev(ctx.raw(), domain.log_size_of_group(), beta.l, gamma.l, da.words(), db.words(), dc.words(),
                               d1.words(), d2.words(), d3.words(), out.words()));
    if (zkb_grand_product_failed(ctx.raw())) throw Error(ZKB_ERR_INVALID, "compute_z1_poly: zero denominator (reference: inverse().unwrap())");
    domain.transform_dev(out.words(), n, true, false);               // poly_from_evals
    DensePolynomial z(n);
    out.download(z.data(), n * 32);
    truncate(z);
    return z;
}

// compute_z2_poly(domain, delta, epsilon, f, t, h1, h2) -> DensePolynomial
inline DensePolynomial compute_z2_poly(const GpuDomain &domain, const Fr &delta, const Fr &epsilon, const std::vector<Fr> &f,
                                       const std::vector<Fr> &t, const std::vector<Fr> &h1, const std::vector<Fr> &h2) {
    const Context &ctx = domain.context();
    const size_t n = domain.size();
    for (const auto *v : {&f, &t, &h1, &h2})
        if (v->size() != n) throw Error(ZKB_ERR_INVALID, "compute_z2_poly: assert_eq!(len, n) failed");   // lookup/mod.rs:40-43
    DeviceBuffer df = to_device(ctx, f), dt = to_device(ctx, t), d1 = to_device(ctx, h1), d2 = to_device(ctx, h2), out(ctx, n * 32);
    ctx.check(zkb_z2_evals_dev(ctx.raw(), domain.log_size_of_group(), delta.l, epsilon.l, df.words(), dt.words(), d1.words(),
                               d2.words(), out.words()));
    if (zkb_grand_product_failed(ctx.raw())) throw Error(ZKB_ERR_INVALID, "compute_z2_poly: zero denominator (reference: inverse().unwrap())");
    domain.transform_dev(out.words(), n, true, false);
    DensePolynomial z(n);
    out.download(z.data(), n * 32);
    truncate(z);
    return z;
}

// The static coset tables of ExtendedProverKey that the quotient kernel streams (x_coset and zh_coset are computed
// inside the kernel): q_m, q_l, q_r, q_o, q_c, q_lookup, q_table, sigma1, sigma2, sigma3, l_1 -- resident in HBM.
struct ExtendedProverKey {
    std::vector<DeviceBuffer> tables;     // 11 buffers of 4n elements, in the order above
};

inline ExtendedProverKey extend_prover_key(const GpuDomain &domain, const std::array<const DensePolynomial *, 10> &polys) {
    const Context &ctx = domain.context();
    auto d4 = GpuDomain::create(ctx, 4 * domain.size());
    if (!d4) throw Error(ZKB_ERR_DOMAIN, "InvalidEvalDomainSize");
    ExtendedProverKey epk;
    for (const DensePolynomial *p : polys) {
        epk.tables.push_back(to_device(ctx, *p, d4->size()));
        d4->transform_dev(epk.tables.back().words(), p->size(), false, true);   // coset_evals_from_poly_ref
    }
    epk.tables.emplace_back(ctx, d4->size() * 32);
    ctx.check(zkb_l1_coset_dev(ctx.raw(), domain.log_size_of_group(), epk.tables.back().words()));
    return epk;
}

// quotient_poly::compute(domain, epk, alpha, beta, gamma, delta, epsilon, z1, z2, a, b, c, pi, h1, h2, t)
inline DensePolynomial quotient_compute(const GpuDomain &domain, const ExtendedProverKey &epk, const Fr &alpha, const Fr &beta,
                                        const Fr &gamma, const Fr &delta, const Fr &epsilon, const DensePolynomial &z1_poly,
                                        const DensePolynomial &z2_poly, const DensePolynomial &a_poly, const DensePolynomial &b_poly,
                                        const DensePolynomial &c_poly, const DensePolynomial &pi_poly, const DensePolynomial &h1_poly,
                                        const DensePolynomial &h2_poly, const DensePolynomial &t_poly) {
    const Context &ctx = domain.context();
    const size_t n = domain.size();
    if (n < 5) throw Error(ZKB_ERR_INVALID, "quotient_poly::compute: assert!(n >= 5)");                 // quotient_poly.rs:44
    auto d4 = GpuDomain::create(ctx, 4 * n);
    if (!d4) throw Error(ZKB_ERR_DOMAIN, "InvalidEvalDomainSize");
    const DensePolynomial *wit_polys[9] = {&z1_poly, &z2_poly, &a_poly, &b_poly, &c_poly, &pi_poly, &t_poly, &h1_poly, &h2_poly};
    std::vector<DeviceBuffer> wit;
    const uint64_t *wit_ptr[9], *epk_ptr[11];
    for (int k = 0; k < 9; ++k) {
        wit.push_back(to_device(ctx, *wit_polys[k], d4->size()));
        d4->transform_dev(wit.back().words(), wit_polys[k]->size(), false, true);
        wit_ptr[k] = wit.back().words();
    }
    if (epk.tables.size() != 11) throw Error(ZKB_ERR_INVALID, "quotient_compute: extended key must hold 11 tables");
    for (int k = 0; k < 11; ++k) epk_ptr[k] = epk.tables[k].words();
    uint64_t ch[20];
    const Fr *cs[5] = {&alpha, &beta, &gamma, &delta, &epsilon};
    for (int k = 0; k < 5; ++k) std::memcpy(ch + 4 * k, cs[k]->l, 32);
    DeviceBuffer out(ctx, d4->size() * 32);
    ctx.check(zkb_quotient_evals_dev(ctx.raw(), domain.log_size_of_group(), ch, wit_ptr, epk_ptr, out.words()));
    d4->transform_dev(out.words(), d4->size(), true, true);          // poly_from_coset_evals
    DensePolynomial q(d4->size());
    out.download(q.data(), q.size() * 32);
    truncate(q);
    return q;
}

// ---------------------------------------------------------------------------------------------------- whole prover / verifier
enum class Transcript : int { Merlin = 0, Ethereum = 1 };      // the T: TranscriptProtocol parameter (plonk.rs:39-46)

struct VerifierKeyData {                                       // keys/mod.rs:180-203, in this boundary's forms
    size_t n = 0;
    std::vector<Fr> pi_roots;
    std::array<G1Affine, 10> commitments{};                    // q_m q_l q_r q_o q_c sigma1 sigma2 sigma3 q_lookup q_table
};

struct G2Affine {                                              // x.c0 x.c1 y.c0 y.c1, Montgomery Fq (arkworks' in-memory order)
    uint64_t w[16];
};

// A proving key resident in HBM (ProverKey + ExtendedProverKey + VerifierKey commitments); the committer key must already
// be resident in the context (GpuKZG10::trim / zkb_srs_load_ck_file).
class PlonkKey {
  public:
    // proof_system::setup with extend = true, from the composer's padded columns (n = 2^log_n elements each)
    static PlonkKey setup(const Context &ctx, unsigned log_n, const std::array<const std::vector<Fr> *, 6> &selectors /* q_m q_l q_r q_o q_c q_lookup */,
                          const std::array<const std::vector<Fr> *, 3> &sigma_evals, size_t table_size, const std::vector<size_t> &pi_positions) {
        const size_t n = size_t(1) << log_n;
        const uint64_t *s[6], *g[3];
        for (int k = 0; k < 6; ++k) {
            if (selectors[k]->size() != n) throw Error(ZKB_ERR_INVALID, "PlonkKey::setup: selector column must hold n elements");
            s[k] = selectors[k]->data()->l;
        }
        for (int k = 0; k < 3; ++k) {
            if (sigma_evals[k]->size() != n) throw Error(ZKB_ERR_INVALID, "PlonkKey::setup: sigma column must hold n elements");
            g[k] = sigma_evals[k]->data()->l;
        }
        zkb_plonk_pk *pk = nullptr;
        ctx.check(zkb_plonk_setup(ctx.raw(), log_n, s, g, table_size, pi_positions.data(), pi_positions.size(), &pk));
        return PlonkKey(ctx, pk);
    }
    // deserialize_from_file::<ProverKey> + ::<VerifierKey> (bin/src/main.rs:274-281)
    static PlonkKey load(const Context &ctx, const std::string &pk_path, const std::string &vk_path, size_t table_size) {
        zkb_plonk_pk *pk = nullptr;
        ctx.check(zkb_plonk_load_keys(ctx.raw(), pk_path.c_str(), vk_path.c_str(), table_size, &pk));
        return PlonkKey(ctx, pk);
    }
    void save(const std::string &pk_path, const std::string &vk_path) const {
        ctx_.check(zkb_plonk_save_keys(ctx_.raw(), pk_.get(), pk_path.c_str(), vk_path.c_str()));
    }
    void set_transcript(Transcript t) { ctx_.check(zkb_plonk_pk_set_transcript(pk_.get(), static_cast<int>(t))); }
    std::array<G1Affine, 10> vk_commitments() const {
        std::array<G1Affine, 10> out{};
        ctx_.check(zkb_plonk_vk_commitments(pk_.get(), reinterpret_cast<uint64_t *>(out.data()), nullptr));
        return out;
    }
    // proof_system::prove: padded wires, the lookup table's entries, one value per public-input row, the 19 blinders in
    // the reference's draw order.  Returns Proof's 802 serialised bytes.
    std::array<uint8_t, 802> prove(const std::vector<Fr> &a, const std::vector<Fr> &b, const std::vector<Fr> &c, const std::vector<Fr> &table,
                                   const std::vector<Fr> &pi_values, const std::array<Fr, 19> &blinders) const {
        std::array<uint8_t, 802> out{};
        const Fr zero{{0, 0, 0, 0}};
        ctx_.check(zkb_plonk_prove(ctx_.raw(), pk_.get(), a.data()->l, b.data()->l, c.data()->l, table.empty() ? zero.l : table.data()->l,
                                   table.size(), pi_values.empty() ? zero.l : pi_values.data()->l, blinders.data()->l, out.data(), nullptr));
        return out;
    }

  private:
    PlonkKey(const Context &ctx, zkb_plonk_pk *pk) : ctx_(ctx) {
        zkb_ctx *raw = ctx.raw();
        pk_ = std::shared_ptr<zkb_plonk_pk>(pk, [raw](zkb_plonk_pk *p) { zkb_plonk_pk_destroy(raw, p); });
    }
    Context ctx_;                                               // keeps the context alive as long as the key
    std::shared_ptr<zkb_plonk_pk> pk_;
};

// Proof::verify: 0 = accepted, 1 / 2 = Error::ProofVerificationError { step }; throws on malformed input.  Host code, no GPU.
inline int verify(const VerifierKeyData &vk, const std::vector<Fr> &pub_inputs, const std::array<uint8_t, 802> &proof, const G2Affine &h,
                  const G2Affine &beta_h, Transcript t = Transcript::Merlin) {
    if (pub_inputs.size() != vk.pi_roots.size()) throw Error(ZKB_ERR_INVALID, "invalid length of public inputs");
    const Fr zero{{0, 0, 0, 0}};
    int rc = zkb_plonk_verify(vk.n, vk.pi_roots.empty() ? zero.l : vk.pi_roots.data()->l, vk.pi_roots.size(),
                              reinterpret_cast<const uint64_t *>(vk.commitments.data()), nullptr, pub_inputs.empty() ? zero.l : pub_inputs.data()->l,
                              proof.data(), h.w, beta_h.w, static_cast<int>(t));
    if (rc < 0) throw Error(rc, "zkb_plonk_verify: malformed verifier key, proof or G2 elements");
    return rc;
}

// deserialize_from_file::<VerifierKey> (vk) and the G2 half of the commitment verifier key (cvk)
inline VerifierKeyData read_vk_file(const std::string &path) {
    VerifierKeyData vk;
    size_t n_roots = 0;
    int inf[10];
    if (zkb_vk_file_read(path.c_str(), &vk.n, nullptr, 0, &n_roots, reinterpret_cast<uint64_t *>(vk.commitments.data()), inf) != ZKB_OK)
        throw Error(ZKB_ERR_INVALID, path + ": not a VerifierKey file");
    vk.pi_roots.resize(n_roots);
    Fr dummy;
    if (zkb_vk_file_read(path.c_str(), &vk.n, n_roots ? vk.pi_roots.data()->l : dummy.l, n_roots, &n_roots,
                         reinterpret_cast<uint64_t *>(vk.commitments.data()), inf) != ZKB_OK)
        throw Error(ZKB_ERR_INVALID, path + ": cannot re-read the VerifierKey file");
    return vk;
}
inline std::pair<G2Affine, G2Affine> read_cvk_file(const std::string &path) {
    std::pair<G2Affine, G2Affine> out{};
    if (zkb_cvk_file_read(path.c_str(), nullptr, nullptr, out.first.w, out.second.w) != ZKB_OK)
        throw Error(ZKB_ERR_INVALID, path + ": not a sonic_pc::VerifierKey file");
    return out;
}

}  // namespace zkb
