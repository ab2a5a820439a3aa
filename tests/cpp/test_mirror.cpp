// test_mirror.cpp -- drives include/zkb200.hpp (the C++ mirror of the reference's D / PC / free-function seams)
// on a GPU and checks every result bit for bit against the CPU oracle (oracle/zkb_oracle.c, test infrastructure).
// Built by __graft_entry__.build(); run by tests/test_gpu_cpp_mirror.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "zkb200.hpp"

extern "C" {
void zko_to_mont(int field, uint64_t *out, const uint64_t *in, size_t n);
void zko_from_mont(int field, uint64_t *out, const uint64_t *in, size_t n);
void zko_normalize(int field, uint64_t *data, size_t n);
int zko_ntt(uint64_t *data, unsigned log_n, int inverse, int coset, int threads);
int zko_msm_g1(const uint64_t *points, const uint64_t *scalars, size_t n, uint64_t out_xy[8], int *is_inf, int threads);
void zko_g1_mul(const uint64_t base_xy[8], const uint64_t *scalars, size_t n, uint64_t *out_xy);
void zko_z1_evals(unsigned log_n, const uint64_t *beta, const uint64_t *gamma, const uint64_t *a, const uint64_t *b,
                  const uint64_t *c, const uint64_t *s1, const uint64_t *s2, const uint64_t *s3, uint64_t *out);
void zko_z2_evals(unsigned log_n, const uint64_t *delta, const uint64_t *eps, const uint64_t *f, const uint64_t *t,
                  const uint64_t *h1, const uint64_t *h2, uint64_t *out);
void zko_epk_free_tables(unsigned log_n, uint64_t *x, uint64_t *zh, uint64_t *l1);
void zko_quotient_evals(unsigned log_n, const uint64_t *ch, const uint64_t *const *wit, const uint64_t *const *epk, uint64_t *out);
}

using namespace zkb;

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t next_u64() {   // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static std::vector<Fr> rand_fr_canonical(size_t n) {
    std::vector<Fr> v(n);
    for (auto &e : v) for (int k = 0; k < 4; ++k) e.l[k] = next_u64();
    zko_normalize(0, reinterpret_cast<uint64_t *>(v.data()), n);
    return v;
}
static std::vector<Fr> rand_fr(size_t n) {   // Montgomery form
    std::vector<Fr> c = rand_fr_canonical(n), m(n);
    zko_to_mont(0, reinterpret_cast<uint64_t *>(m.data()), reinterpret_cast<const uint64_t *>(c.data()), n);
    return m;
}
static uint64_t *W(std::vector<Fr> &v) { return reinterpret_cast<uint64_t *>(v.data()); }
[[maybe_unused]] static const uint64_t *W(const std::vector<Fr> &v) { return reinterpret_cast<const uint64_t *>(v.data()); }

static int failures = 0;
#define CHECK(cond, name)                                   \
    do {                                                    \
        if (cond) std::printf("ok   %s\n", name);           \
        else { std::printf("FAIL %s\n", name); ++failures; } \
    } while (0)

static std::vector<Fr> oracle_coset4(const DensePolynomial &p, unsigned log_n) {
    std::vector<Fr> buf(size_t(4) << log_n, Fr{{0, 0, 0, 0}});
    std::memcpy(buf.data(), p.data(), p.size() * 32);
    zko_ntt(W(buf), log_n + 2, 0, 1, 0);
    return buf;
}

int main() {
    Context ctx(0);
    // ---- D: GpuDomain
    {
        auto none = GpuDomain::create(ctx, (size_t(1) << 28) + 1);
        CHECK(!none.has_value(), "GpuDomain::create above 2^28 is None (InvalidEvalDomainSize)");
        auto dom = GpuDomain::create(ctx, 5000);
        CHECK(dom && dom->size() == 8192 && dom->log_size_of_group() == 13, "GpuDomain::create rounds up to a power of two");
        std::vector<Fr> coeffs = rand_fr(5000), padded = coeffs;
        padded.resize(8192, Fr{{0, 0, 0, 0}});
        const int modes[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
        bool all = true;
        for (auto &m : modes) {
            std::vector<Fr> got = coeffs, exp = padded;
            if (m[0] && m[1]) dom->coset_ifft_in_place(got); else if (m[0]) dom->ifft_in_place(got);
            else if (m[1]) dom->coset_fft_in_place(got); else dom->fft_in_place(got);
            zko_ntt(W(exp), 13, m[0], m[1], 0);
            all = all && got == exp;
        }
        CHECK(all, "fft / ifft / coset_fft / coset_ifft (zero padded) == oracle");
    }
    // ---- PC: GpuKZG10
    const size_t N = 700;
    std::vector<G1Affine> srs(N);
    {
        uint64_t gen_canon[8] = {1, 0, 0, 0, 2, 0, 0, 0}, gen[8];
        zko_to_mont(1, gen, gen_canon, 2);
        std::vector<Fr> k = rand_fr_canonical(N);
        zko_g1_mul(gen, W(k), N, reinterpret_cast<uint64_t *>(srs.data()));
    }
    for (int fixed = 0; fixed < 2; ++fixed) {
        GpuKZG10 kzg(ctx);
        kzg.trim(srs, fixed != 0);
        DensePolynomial p = rand_fr(N - 10);
        p[0] = p[1] = Fr{{0, 0, 0, 0}};                     // leading zeros are skipped, trailing ones dropped
        p.push_back(Fr{{0, 0, 0, 0}});
        G1Affine got = kzg.commit_one(p);
        std::vector<Fr> canon(p.size());
        zko_from_mont(0, W(canon), W(p), p.size());
        G1Affine exp{};
        int inf = 0;
        zko_msm_g1(reinterpret_cast<const uint64_t *>(srs.data()), W(canon), p.size(), reinterpret_cast<uint64_t *>(&exp), &inf, 0);
        CHECK(got == exp && !got.infinity(), fixed ? "GpuKZG10::commit (fixed-base tables) == oracle" : "GpuKZG10::commit == oracle");
        DensePolynomial zero(5, Fr{{0, 0, 0, 0}});
        CHECK(kzg.commit_one(zero).infinity(), "commit(0) is the identity");
        bool threw = false;
        try { DensePolynomial big = rand_fr(N + 1); kzg.commit_one(big); } catch (const Error &e) { threw = e.code == ZKB_ERR_NO_SRS; }
        CHECK(threw, "commit beyond the key raises TooManyCoefficients (ZKB_ERR_NO_SRS)");
        std::vector<std::array<uint64_t, 4>> sc(13);
        std::vector<Fr> scv = rand_fr_canonical(13);
        std::memcpy(sc.data(), scv.data(), 13 * 32);
        std::vector<G1Affine> pts(srs.begin(), srs.begin() + 13);
        G1Affine m = kzg.multi_scalar_mul(pts, sc), me{};
        zko_msm_g1(reinterpret_cast<const uint64_t *>(pts.data()), W(scv), 13, reinterpret_cast<uint64_t *>(&me), &inf, 0);
        CHECK(m == me, "HomomorphicCommitment::multi_scalar_mul (13 points, the verifier's shape) == oracle");
    }
    // ---- compute_z1_poly / compute_z2_poly / quotient_poly::compute
    {
        const unsigned log_n = 9;
        const size_t n = size_t(1) << log_n, n4 = 4 * n;
        auto dom = *GpuDomain::create(ctx, n);
        std::vector<Fr> col[6];
        for (auto &c : col) c = rand_fr(n);
        std::vector<Fr> ch = rand_fr(5);
        DensePolynomial z1 = compute_z1_poly(dom, ch[1], ch[2], col[0], col[1], col[2], col[3], col[4], col[5]);
        std::vector<Fr> exp(n);
        zko_z1_evals(log_n, ch[1].l, ch[2].l, W(col[0]), W(col[1]), W(col[2]), W(col[3]), W(col[4]), W(col[5]), W(exp));
        zko_ntt(W(exp), log_n, 1, 0, 0);
        truncate(exp);
        CHECK(z1 == exp, "compute_z1_poly == oracle (evaluations + iFFT)");
        DensePolynomial z2 = compute_z2_poly(dom, ch[3], ch[4], col[0], col[1], col[2], col[3]);
        exp.assign(n, Fr{{0, 0, 0, 0}});
        zko_z2_evals(log_n, ch[3].l, ch[4].l, W(col[0]), W(col[1]), W(col[2]), W(col[3]), W(exp));
        zko_ntt(W(exp), log_n, 1, 0, 0);
        truncate(exp);
        CHECK(z2 == exp, "compute_z2_poly == oracle");
        bool threw = false;
        try { compute_z1_poly(dom, ch[1], ch[2], col[0], col[1], col[2], col[3], col[4], std::vector<Fr>(n - 1)); } catch (const Error &) { threw = true; }
        CHECK(threw, "compute_z1_poly rejects a column of the wrong length (assert_eq! in the reference)");

        // quotient: witness polynomials of the prover's lengths, 10 key polynomials of length n
        const size_t lens[9] = {n + 3, n + 3, n + 2, n + 2, n + 2, n, n, n + 3, n + 2};   // z1 z2 a b c pi t h1 h2
        DensePolynomial wp[9], kp[10];
        for (int k = 0; k < 9; ++k) wp[k] = rand_fr(lens[k]);
        for (auto &p : kp) p = rand_fr(n);
        std::array<const DensePolynomial *, 10> kptr;
        for (int k = 0; k < 10; ++k) kptr[k] = &kp[k];
        ExtendedProverKey epk = extend_prover_key(dom, kptr);
        // argument order of quotient_poly::compute: z1, z2, a, b, c, pi, h1, h2, t
        DensePolynomial q = quotient_compute(dom, epk, ch[0], ch[1], ch[2], ch[3], ch[4], wp[0], wp[1], wp[2], wp[3], wp[4], wp[5],
                                             wp[7], wp[8], wp[6]);
        std::vector<std::vector<Fr>> ow, oe;
        for (int k = 0; k < 9; ++k) ow.push_back(oracle_coset4(wp[k], log_n));
        for (int k = 0; k < 10; ++k) oe.push_back(oracle_coset4(kp[k], log_n));
        std::vector<Fr> x(n4), zh(n4), l1(n4);
        zko_epk_free_tables(log_n, W(x), W(zh), W(l1));
        const uint64_t *wptr[9], *eptr[13];
        for (int k = 0; k < 9; ++k) wptr[k] = W(ow[k]);
        for (int k = 0; k < 10; ++k) eptr[k] = W(oe[k]);
        eptr[10] = W(x); eptr[11] = W(l1); eptr[12] = W(zh);
        std::vector<Fr> qe(n4);
        zko_quotient_evals(log_n, W(ch), wptr, eptr, W(qe));
        zko_ntt(W(qe), log_n + 2, 1, 1, 0);
        truncate(qe);
        CHECK(q == qe, "extend_prover_key + quotient_poly::compute == oracle (9 coset FFTs, fused kernel, coset iFFT)");
    }
    std::printf(failures ? "%d FAILED\n" : "all C++ mirror checks passed\n", failures);
    return failures ? 1 : 0;
}
