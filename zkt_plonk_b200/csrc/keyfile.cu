// keyfile.cu -- the reference CLI's on-disk keys (host code only).
//
// `zkt compile` writes ck / cvk / pk / (epk) / vk with ark-serialize 0.3 `serialize_unchecked`
// (bin/src/parser.rs:5-29, bin/src/main.rs:96-113) and `prove-withdraw` reads them back (main.rs:274-281).  The byte
// layout is what `#[derive(CanonicalSerialize)]` produces, field after field in declaration order:
//   usize / u64          8 bytes little endian
//   bool                 1 byte; Option<T> = bool + T when Some
//   Vec<T>, String       u64 length + the items (String: its UTF-8 bytes)
//   Fp256 / Fp384        32 / 48 bytes little endian, CANONICAL (not Montgomery) integer (Fr is 32 bytes on every curve; Fq is
//                        32 bytes on BN254 -- the CLI's curve -- and 48 on BLS12-381 / BLS12-377, the other curves plonk.rs:226-254
//                        instantiates: the same derive(CanonicalSerialize) layouts with wider base-field elements)
//   GroupAffine (unchecked = uncompressed)   x || y (FQB bytes each), SWFlags in the top bits of the last byte:
//                        bit 6 = infinity (arkworks stores the identity as (0, 1, true)); bit 7 unused here
//   Rc<T>, PhantomData   as T / nothing
// Types:
//   ck  = sonic_pc::CommitterKey<Bn254> { powers_of_g: Vec<G1Affine>, powers_of_gamma_g: Vec<G1Affine>,
//         shifted_powers_of_g: Option<Vec<G1Affine>>, shifted_powers_of_gamma_g: Option<BTreeMap<usize, Vec<G1Affine>>>,
//         enforced_degree_bounds: Option<Vec<usize>>, max_degree: usize }   [ark-poly-commit 0.3, un-vendored: recalled;
//         PC::trim(pp, 4n, 0, None) (plonk.rs:79-85) leaves 4n + 1 powers, 2 gamma powers and the three Options None]
//   pk  = ProverKey<Fr> { arith { q_m q_l q_r q_o q_c }, perm { sigma1 sigma2 sigma3 }, lookup { q_lookup q_table } }
//         (keys/mod.rs:29-40, arithmetic.rs:20-32, permutation.rs:20-31, lookup.rs:19-26), each a
//         LabeledPolynomial { label: String, polynomial: Rc<DensePolynomial { coeffs: Vec<Fr> }>,
//         degree_bound: Option<usize>, hiding_bound: Option<usize> } with the labels of setup.rs:93-102
//   vk  = VerifierKey { n: usize, pi_roots: Vec<Fr>, arith { q_m q_l q_r q_o q_c }, perm { sigma1 sigma2 sigma3 },
//         lookup { q_lookup q_table } } of kzg10::Commitment(G1Affine)   (keys/mod.rs:180-203)
//   cvk = sonic_pc::VerifierKey<Bn254> { g: G1Affine, gamma_g: G1Affine, h: G2Affine, beta_h: G2Affine, prepared_h,
//         prepared_beta_h: G2Prepared, degree_bounds_and_neg_powers_of_h: Option<..>, supported_degree, max_degree }
//         [ark-poly-commit 0.3: recalled].  Only the first four fields are read (12 FQB = 384 bytes on BN254: the verifier needs h and
//         beta_h, and derives its own line coefficients); G2Affine = x.c0 x.c1 y.c0 y.c1, flags in the last byte of y.c1.
// The in-memory side of every function is this library's usual form (Montgomery limbs, identity = (0, 0)).
// The epk file (13 coset tables, 1.7 GiB at n = 2^20) is never read: the key loader rebuilds those tables in HBM
// with 10 coset NTTs, which is faster than reading them from disk.
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <thread>
#include <vector>

#include "ctx.h"

using namespace zkb;
using host::Fe;
using host::Fq;
constexpr int QL = host::FQ_L;                // 64-bit words of a base-field element
constexpr size_t FQB = 8 * QL, PT_BYTES = 2 * FQB;   // bytes of a base-field element / of an uncompressed G1 point in a file
constexpr int AFF_W = 2 * QL;                 // words of an affine point at the C boundary

namespace {

const char *PK_LABELS[10] = {"q_m", "q_l", "q_r", "q_o", "q_c", "sigma1", "sigma2", "sigma3", "q_lookup", "q_table"};

struct File {
    FILE *f = nullptr;
    explicit File(const char *path, const char *mode) { f = path ? fopen(path, mode) : nullptr; }
    ~File() { if (f) fclose(f); }
    bool rd(void *dst, size_t n) { return n == 0 || fread(dst, 1, n, f) == n; }
    bool wr(const void *src, size_t n) { return n == 0 || fwrite(src, 1, n, f) == n; }
    bool rd_u64(uint64_t *v) { return rd(v, 8); }                 // little-endian hosts only (x86-64 / aarch64)
    bool wr_u64(uint64_t v) { return wr(&v, 8); }
    bool skip(uint64_t n) { return fseeko(f, (off_t)n, SEEK_CUR) == 0; }
    bool at_eof() { int c = fgetc(f); if (c == EOF) return true; ungetc(c, f); return false; }
};

template <int L> inline host::FeT<L> to_mont(const host::FeT<L> &canon, const host::ParamsT<L> &P) {
    host::FeT<L> r2;
    memcpy(r2.l, P.r2, 8 * L);
    return host::mul(canon, r2, P);
}
template <int L> inline host::FeT<L> from_mont(const host::FeT<L> &m, const host::ParamsT<L> &P) {
    host::FeT<L> one;
    memset(one.l, 0, sizeof one.l);
    one.l[0] = 1;
    return host::mul(m, one, P);
}

// PT_BYTES file bytes -> x || y Montgomery (identity -> zeros).  false: a coordinate is not below the modulus.
bool point_from_file(const uint8_t *src, uint64_t *xy) {
    Fq x, y;
    memcpy(x.l, src, FQB);
    memcpy(y.l, src + FQB, FQB);
    const bool inf = (y.l[QL - 1] >> 62) & 1;
    if (inf && ((y.l[QL - 1] >> 63) & 1)) return false;           // SWFlags::from_u8: both flag bits set is no valid flag
    y.l[QL - 1] &= ~(3ULL << 62);
    if (host::ge<QL>(x.l, host::FQ.p) || host::ge<QL>(y.l, host::FQ.p)) return false;
    if (inf) { memset(xy, 0, PT_BYTES); return true; }
    x = to_mont(x, host::FQ);
    y = to_mont(y, host::FQ);
    memcpy(xy, x.l, FQB);
    memcpy(xy + QL, y.l, FQB);
    return true;
}
void point_to_file(const uint64_t *xy, bool inf, uint8_t *dst) {
    bool zero = inf;
    if (!zero) { zero = true; for (int i = 0; i < AFF_W; ++i) if (xy[i]) zero = false; }
    if (zero) {
        memset(dst, 0, PT_BYTES);
        dst[FQB] = 1;                                             // (0, 1, true)
        dst[PT_BYTES - 1] |= 1 << 6;
        return;
    }
    Fq x, y;
    memcpy(x.l, xy, FQB);
    memcpy(y.l, xy + QL, FQB);
    x = from_mont(x, host::FQ);
    y = from_mont(y, host::FQ);
    memcpy(dst, x.l, FQB);
    memcpy(dst + FQB, y.l, FQB);
}

// `count` points, in parallel over a few host threads (2^20 points = 4 M host products).
bool points_from_file(const uint8_t *src, size_t count, uint64_t *xy) {
    unsigned hw = std::thread::hardware_concurrency();
    const size_t T = count < 4096 ? 1 : (hw ? (hw > 16 ? 16 : hw) : 4);
    std::vector<char> ok(T, 1);
    std::vector<std::thread> th;
    for (size_t t = 0; t < T; ++t) {
        const size_t lo = count * t / T, hi = count * (t + 1) / T;
        auto work = [=, &ok]() { for (size_t i = lo; i < hi; ++i) if (!point_from_file(src + PT_BYTES * i, xy + AFF_W * i)) ok[t] = 0; };
        if (T == 1) work(); else th.emplace_back(work);
    }
    for (auto &x : th) x.join();
    for (char c : ok) if (!c) return false;
    return true;
}

bool read_none(File &F) {                                          // an Option that must be None
    uint8_t b;
    return F.rd(&b, 1) && b == 0;
}

// header of a ck file: positions the stream at the first power; returns the counts
int ck_open(File &F, uint64_t *n_powers) {
    if (!F.f) return ZKB_ERR_INVALID;
    if (!F.rd_u64(n_powers) || *n_powers > (1ULL << 32)) return ZKB_ERR_INVALID;
    return ZKB_OK;
}

}  // namespace

extern "C" {

int zkb_ck_file_info(const char *path, size_t *n_powers, size_t *max_degree) {
    File F(path, "rb");
    uint64_t n, ng, md;
    int rc = ck_open(F, &n);
    if (rc) return rc;
    if (!F.skip(PT_BYTES * n) || !F.rd_u64(&ng) || ng > (1ULL << 32) || !F.skip(PT_BYTES * ng)) return ZKB_ERR_INVALID;
    // PC::trim without enforced degree bounds: the three Options are None (a key with degree bounds is not one the
    // reference writes; refuse it instead of guessing)
    if (!read_none(F) || !read_none(F) || !read_none(F) || !F.rd_u64(&md) || !F.at_eof()) return ZKB_ERR_INVALID;
    if (n_powers) *n_powers = (size_t)n;
    if (max_degree) *max_degree = (size_t)md;
    return ZKB_OK;
}

int zkb_ck_file_read(const char *path, size_t first, size_t count, uint64_t *xy_mont_out) {
    if (!xy_mont_out && count) return ZKB_ERR_INVALID;
    File F(path, "rb");
    uint64_t n;
    int rc = ck_open(F, &n);
    if (rc) return rc;
    if (first > n || count > n - first || !F.skip(PT_BYTES * (uint64_t)first)) return ZKB_ERR_INVALID;
    const size_t CH = 1 << 16;
    std::vector<uint8_t> buf(PT_BYTES * (count < CH ? count : CH));
    for (size_t done = 0; done < count;) {
        const size_t m = count - done < CH ? count - done : CH;
        if (!F.rd(buf.data(), PT_BYTES * m) || !points_from_file(buf.data(), m, xy_mont_out + AFF_W * done)) return ZKB_ERR_INVALID;
        done += m;
    }
    return ZKB_OK;
}

int zkb_ck_file_write(const char *path, const uint64_t *xy_mont, size_t n_powers, const uint64_t *gamma_xy_mont, size_t n_gamma,
                      size_t max_degree) {
    if ((!xy_mont && n_powers) || (!gamma_xy_mont && n_gamma)) return ZKB_ERR_INVALID;
    File F(path, "wb");
    if (!F.f) return ZKB_ERR_INVALID;
    uint8_t b[PT_BYTES];
    bool ok = F.wr_u64(n_powers);
    for (size_t i = 0; ok && i < n_powers; ++i) { point_to_file(xy_mont + AFF_W * i, false, b); ok = F.wr(b, PT_BYTES); }
    ok = ok && F.wr_u64(n_gamma);
    for (size_t i = 0; ok && i < n_gamma; ++i) { point_to_file(gamma_xy_mont + AFF_W * i, false, b); ok = F.wr(b, PT_BYTES); }
    const uint8_t none[3] = {0, 0, 0};
    ok = ok && F.wr(none, 3) && F.wr_u64(max_degree);
    return ok ? ZKB_OK : ZKB_ERR_INVALID;
}

// Reads the committer key's powers_of_g straight into the resident SRS (what main.rs:274 + the first commit do).
int zkb_srs_load_ck_file(zkb_ctx *ctx, const char *path, size_t max_points) {
    if (!ctx) return ZKB_ERR_INVALID;
    size_t n = 0, md = 0;
    int rc = zkb_ck_file_info(path, &n, &md);
    if (rc) ZKB_FAIL(ctx, rc, "zkb_srs_load_ck_file: not a sonic_pc::CommitterKey file written without degree bounds");
    if (max_points && max_points < n) n = max_points;
    if (n == 0) ZKB_FAIL(ctx, ZKB_ERR_INVALID, "zkb_srs_load_ck_file: the key holds no powers");
    std::vector<uint64_t> xy;
    try {
        xy.resize((size_t)AFF_W * n);                              // nothing may throw across the C boundary
    } catch (const std::bad_alloc &) {
        ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_srs_load_ck_file: not enough host memory for the committer key");
    }
    rc = zkb_ck_file_read(path, 0, n, xy.data());
    if (rc) ZKB_FAIL(ctx, rc, "zkb_srs_load_ck_file: a coordinate is not a canonical Fq element");
    return zkb_srs_load_g1(ctx, xy.data(), n);
}

// ---- cvk: g, gamma_g (G1) and h, beta_h (G2) from the head of the file
int zkb_cvk_file_read(const char *path, uint64_t *g_xy, uint64_t *gamma_g_xy, uint64_t *h_xy, uint64_t *beta_h_xy) {
    if (!h_xy || !beta_h_xy) return ZKB_ERR_INVALID;
    File F(path, "rb");
    if (!F.f) return ZKB_ERR_INVALID;
    uint8_t b[12 * FQB];                                          // g, gamma_g (2 FQB each), h, beta_h (4 FQB each)
    if (!F.rd(b, sizeof b)) return ZKB_ERR_INVALID;
    uint64_t tmp[AFF_W];
    if (!point_from_file(b, g_xy ? g_xy : tmp) || !point_from_file(b + PT_BYTES, gamma_g_xy ? gamma_g_xy : tmp)) return ZKB_ERR_INVALID;
    uint64_t *dst[2] = {h_xy, beta_h_xy};
    for (int k = 0; k < 2; ++k) {
        const uint8_t *src = b + 2 * PT_BYTES + 4 * FQB * k;
        Fq c[4];
        memcpy(c, src, 4 * FQB);
        const bool inf = (c[3].l[QL - 1] >> 62) & 1;
        c[3].l[QL - 1] &= ~(3ULL << 62);
        for (int j = 0; j < 4; ++j) {
            if (host::ge<QL>(c[j].l, host::FQ.p)) return ZKB_ERR_INVALID;
            c[j] = to_mont(c[j], host::FQ);
        }
        if (inf) memset(dst[k], 0, 4 * FQB); else memcpy(dst[k], c, 4 * FQB);
    }
    return ZKB_OK;
}

// ---- pk
static int pk_walk(File &F, uint64_t *const out[10], const size_t *caps, size_t lens[10]) {
    if (!F.f) return ZKB_ERR_INVALID;
    for (int k = 0; k < 10; ++k) {
        uint64_t ll, n;
        char label[16];
        if (!F.rd_u64(&ll) || ll != strlen(PK_LABELS[k]) || !F.rd(label, ll) || memcmp(label, PK_LABELS[k], ll) != 0) return ZKB_ERR_INVALID;
        if (!F.rd_u64(&n) || n > (1ULL << 28)) return ZKB_ERR_INVALID;
        lens[k] = (size_t)n;
        if (out) {
            if (n > caps[k] || (!out[k] && n)) return ZKB_ERR_INVALID;
            if (!F.rd(out[k], 32 * n)) return ZKB_ERR_INVALID;
            for (uint64_t i = 0; i < n; ++i) {
                Fe c;
                memcpy(c.l, out[k] + 4 * i, 32);
                if (host::ge(c.l, host::FR.p)) return ZKB_ERR_INVALID;
                c = to_mont(c, host::FR);
                memcpy(out[k] + 4 * i, c.l, 32);
            }
        } else if (!F.skip(32 * n)) {
            return ZKB_ERR_INVALID;
        }
        if (!read_none(F) || !read_none(F)) return ZKB_ERR_INVALID;   // degree_bound, hiding_bound: label_polynomial! sets neither
    }
    return F.at_eof() ? ZKB_OK : ZKB_ERR_INVALID;
}

int zkb_pk_file_info(const char *path, size_t lens[10]) {
    if (!lens) return ZKB_ERR_INVALID;
    File F(path, "rb");
    return pk_walk(F, nullptr, nullptr, lens);
}

int zkb_pk_file_read(const char *path, uint64_t *const coeffs_mont_out[10], const size_t caps[10], size_t lens[10]) {
    if (!coeffs_mont_out || !caps || !lens) return ZKB_ERR_INVALID;
    File F(path, "rb");
    return pk_walk(F, coeffs_mont_out, caps, lens);
}

int zkb_pk_file_write(const char *path, const uint64_t *const coeffs_mont[10], const size_t lens[10]) {
    if (!coeffs_mont || !lens) return ZKB_ERR_INVALID;
    File F(path, "wb");
    if (!F.f) return ZKB_ERR_INVALID;
    bool ok = true;
    for (int k = 0; ok && k < 10; ++k) {
        size_t n = lens[k];
        while (n && !(coeffs_mont[k][4 * (n - 1)] | coeffs_mont[k][4 * (n - 1) + 1] | coeffs_mont[k][4 * (n - 1) + 2] | coeffs_mont[k][4 * (n - 1) + 3]))
            --n;                                                   // DensePolynomial::from_coefficients_vec drops trailing zeros
        const uint64_t ll = strlen(PK_LABELS[k]);
        ok = F.wr_u64(ll) && F.wr(PK_LABELS[k], ll) && F.wr_u64(n);
        for (size_t i = 0; ok && i < n; ++i) {
            Fe c;
            memcpy(c.l, coeffs_mont[k] + 4 * i, 32);
            c = from_mont(c, host::FR);
            ok = F.wr(c.l, 32);
        }
        const uint8_t none[2] = {0, 0};
        ok = ok && F.wr(none, 2);
    }
    return ok ? ZKB_OK : ZKB_ERR_INVALID;
}

// ---- vk
int zkb_vk_file_read(const char *path, size_t *n, uint64_t *pi_roots_mont, size_t cap_roots, size_t *n_roots, uint64_t *commits_xy,
                     int is_inf[10]) {
    if (!n || !n_roots || !commits_xy) return ZKB_ERR_INVALID;
    File F(path, "rb");
    if (!F.f) return ZKB_ERR_INVALID;
    uint64_t nn, nr;
    if (!F.rd_u64(&nn) || !F.rd_u64(&nr) || nr > (1ULL << 28)) return ZKB_ERR_INVALID;
    *n = (size_t)nn;
    *n_roots = (size_t)nr;
    for (uint64_t i = 0; i < nr; ++i) {
        Fe c;
        if (!F.rd(c.l, 32) || host::ge(c.l, host::FR.p)) return ZKB_ERR_INVALID;
        if (pi_roots_mont && i < cap_roots) { c = to_mont(c, host::FR); memcpy(pi_roots_mont + 4 * i, c.l, 32); }
    }
    for (int k = 0; k < 10; ++k) {
        uint8_t b[PT_BYTES];
        if (!F.rd(b, PT_BYTES)) return ZKB_ERR_INVALID;
        if (is_inf) is_inf[k] = (b[PT_BYTES - 1] >> 6) & 1;
        if (!point_from_file(b, commits_xy + AFF_W * k)) return ZKB_ERR_INVALID;
    }
    return F.at_eof() ? ZKB_OK : ZKB_ERR_INVALID;
}

int zkb_vk_file_write(const char *path, size_t n, const uint64_t *pi_roots_mont, size_t n_roots, const uint64_t *commits_xy,
                      const int is_inf[10]) {
    if ((!pi_roots_mont && n_roots) || !commits_xy) return ZKB_ERR_INVALID;
    File F(path, "wb");
    if (!F.f) return ZKB_ERR_INVALID;
    bool ok = F.wr_u64(n) && F.wr_u64(n_roots);
    for (size_t i = 0; ok && i < n_roots; ++i) {
        Fe c;
        memcpy(c.l, pi_roots_mont + 4 * i, 32);
        c = from_mont(c, host::FR);
        ok = F.wr(c.l, 32);
    }
    for (int k = 0; ok && k < 10; ++k) {
        uint8_t b[PT_BYTES];
        point_to_file(commits_xy + AFF_W * k, is_inf && is_inf[k], b);
        ok = F.wr(b, PT_BYTES);
    }
    return ok ? ZKB_OK : ZKB_ERR_INVALID;
}

}  // extern "C"
