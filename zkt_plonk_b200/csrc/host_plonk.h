// host_plonk.h -- host-side pieces shared by the prover round driver (prover.cu) and the verifier (verify.cu):
// Keccak-f[1600], STROBE-128 / Merlin and the reference's two TranscriptProtocol implementations
// (plonk-core/src/transcript.rs:49-109, gadgets/src/transcript.rs:8-90), ark-serialize's field / point byte forms,
// and a few Fr shorthands.  Internal linkage: every including translation unit gets its own copy.
#pragma once
#include <string.h>

#include <vector>

#include "host_ff.h"

namespace {

using namespace zkb;
using zkb::host::Fe;
using zkb::host::Fq;
constexpr int AFF_W = 2 * host::FQ_L;          // 64-bit words of an affine point at the C boundary (8 / 12)
constexpr int FQB = 8 * host::FQ_L;           // bytes of a base-field element: 32 (BN254) or 48 (BLS12-381 / BLS12-377)
// Proof (proof.rs:112-154): 11 compressed commitments, 2 x (compressed witness + the None tag of random_v), 12 evaluations
constexpr size_t PROOF_BYTES = 11 * FQB + 2 * (FQB + 1) + 12 * 32;   // 802 on BN254, 1010 on the BLS12 curves

// ================================================================================================ Keccak / STROBE / Merlin
const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
const int KECCAK_ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};

inline uint64_t rol64(uint64_t v, int r) { return r ? (v << r) | (v >> (64 - r)) : v; }

void keccak_f1600(uint8_t st[200]) {
    uint64_t a[5][5];
    for (int x = 0; x < 5; ++x) for (int y = 0; y < 5; ++y) memcpy(&a[x][y], st + 8 * (x + 5 * y), 8);
    for (int rnd = 0; rnd < 24; ++rnd) {
        uint64_t c[5], d[5], b[5][5];
        for (int x = 0; x < 5; ++x) c[x] = a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4];
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int x = 0; x < 5; ++x) for (int y = 0; y < 5; ++y) a[x][y] ^= d[x];
        for (int x = 0; x < 5; ++x) for (int y = 0; y < 5; ++y) b[y][(2 * x + 3 * y) % 5] = rol64(a[x][y], KECCAK_ROT[x][y]);
        for (int x = 0; x < 5; ++x) for (int y = 0; y < 5; ++y) a[x][y] = b[x][y] ^ (~b[(x + 1) % 5][y] & b[(x + 2) % 5][y]);
        a[0][0] ^= KECCAK_RC[rnd];
    }
    for (int x = 0; x < 5; ++x) for (int y = 0; y < 5; ++y) memcpy(st + 8 * (x + 5 * y), &a[x][y], 8);
}

struct Strobe128 {
    static constexpr int R = 166;
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    uint8_t st[200];
    int pos = 0, pos_begin = 0, cur_flags = 0;
    explicit Strobe128(const char *label) {
        memset(st, 0, 200);
        const uint8_t head[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(st, head, 6);
        memcpy(st + 6, "STROBEv1.0.2", 12);
        keccak_f1600(st);
        meta_ad((const uint8_t *)label, strlen(label), false);
    }
    void run_f() {
        st[pos] ^= (uint8_t)pos_begin;
        st[pos + 1] ^= 0x04;
        st[R + 1] ^= 0x80;
        keccak_f1600(st);
        pos = 0;
        pos_begin = 0;
    }
    void absorb(const uint8_t *d, size_t n) {
        for (size_t i = 0; i < n; ++i) { st[pos] ^= d[i]; if (++pos == R) run_f(); }
    }
    void squeeze(uint8_t *d, size_t n) {
        for (size_t i = 0; i < n; ++i) { d[i] = st[pos]; st[pos] = 0; if (++pos == R) run_f(); }
    }
    void begin_op(int flags, bool more) {
        if (more) return;
        int old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {(uint8_t)old_begin, (uint8_t)flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
    }
    void meta_ad(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
    void ad(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
    void prf(uint8_t *d, size_t n) { begin_op(FLAG_I | FLAG_A | FLAG_C, false); squeeze(d, n); }
};

// canonical little-endian bytes of a Montgomery field element (ToBytes::write / CanonicalSerialize of Fp256 / Fp384): 8 L bytes
template <int L>
void fe_bytes(const host::FeT<L> &m, const host::ParamsT<L> &P, uint8_t *out) {
    host::FeT<L> one;
    memset(one.l, 0, sizeof one.l);
    one.l[0] = 1;
    host::FeT<L> c = host::mul(m, one, P);
    memcpy(out, c.l, 8 * L);
}

struct Pt {                                   // affine G1, Montgomery; inf = identity
    Fq x, y;
    bool inf;
};

// Keccak-256 (the pre-NIST padding 0x01 .. 0x80, rate 136) on the same permutation: sha3::Keccak256 of
// gadgets/src/transcript.rs:4
void keccak256(const uint8_t *data, size_t n, uint8_t out[32]) {
    uint8_t st[200];
    memset(st, 0, 200);
    const size_t rate = 136;
    while (n >= rate) {
        for (size_t i = 0; i < rate; ++i) st[i] ^= data[i];
        keccak_f1600(st);
        data += rate;
        n -= rate;
    }
    for (size_t i = 0; i < n; ++i) st[i] ^= data[i];
    st[n] ^= 0x01;
    st[rate - 1] ^= 0x80;
    keccak_f1600(st);
    memcpy(out, st, 32);
}

// canonical BIG-endian bytes (into_repr().to_bytes_be())
template <int L>
void fe_bytes_be(const host::FeT<L> &m, const host::ParamsT<L> &P, uint8_t *out) {
    uint8_t le[8 * L];
    fe_bytes(m, P, le);
    for (int i = 0; i < 8 * L; ++i) out[i] = le[8 * L - 1 - i];
}

// TranscriptProtocol for the reference's two transcripts.  kind 0: MerlinTranscript (plonk-core/src/transcript.rs:49-109,
// the default binary's); kind 1: EthereumTranscript (gadgets/src/transcript.rs:8-90: two chained Keccak-256 states,
// labels ignored, every item big-endian, challenges = digest mod 2^253), pinned by the reference's own known-answer
// test (gadgets/src/transcript.rs:100-127) through zkb_test_transcript.
struct Transcript {
    int kind;
    Strobe128 s;
    uint8_t st0[32], st1[32];
    uint32_t counter = 0;
    explicit Transcript(const char *label, int kind_ = 0) : kind(kind_), s("Merlin v1.0") {
        memset(st0, 0, 32);
        memset(st1, 0, 32);
        if (kind == 0) append_message("dom-sep", (const uint8_t *)label, strlen(label));
    }
    void eth_append(const uint8_t *item, size_t n) {              // append_bytes_without_label (transcript.rs:19-36)
        std::vector<uint8_t> d(65 + n);
        memcpy(d.data() + 1, st0, 32);
        memcpy(d.data() + 33, st1, 32);
        if (n) memcpy(d.data() + 65, item, n);
        uint8_t n0[32], n1[32];
        d[0] = 0;
        keccak256(d.data(), d.size(), n0);
        d[0] = 1;
        keccak256(d.data(), d.size(), n1);
        memcpy(st0, n0, 32);
        memcpy(st1, n1, 32);
    }
    void append_message(const char *label, const uint8_t *msg, size_t n) {    // Merlin only
        s.meta_ad((const uint8_t *)label, strlen(label), false);
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s.meta_ad(len, 4, true);
        s.ad(msg, n, false);
    }
    void append_u64(const char *label, uint64_t v) {
        uint8_t b[8];
        if (kind == 1) {
            for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(v >> (56 - 8 * i));
            eth_append(b, 8);
            return;
        }
        memcpy(b, &v, 8);
        append_message(label, b, 8);
    }
    void append_scalar(const char *label, const Fe &v) {
        uint8_t b[32];
        if (kind == 1) { fe_bytes_be(v, host::FR, b); eth_append(b, 32); return; }
        fe_bytes(v, host::FR, b);
        append_message(label, b, 32);
    }
    void append_scalars(const char *label, const Fe *v, size_t n) {
        if (kind == 1) { for (size_t i = 0; i < n; ++i) append_scalar(label, v[i]); return; }
        std::vector<uint8_t> b(32 * n);
        for (size_t i = 0; i < n; ++i) fe_bytes(v[i], host::FR, b.data() + 32 * i);
        append_message(label, b.data(), b.size());
    }
    void append_commitment(const char *label, const Pt &p) {
        if (kind == 1) {                                           // x then y, each its own item; arkworks' zero is (0, 1, true)
            uint8_t bx[FQB], by[FQB];                              // (the reference binds this transcript to Bn254: 32-byte items)
            memset(bx, 0, FQB);
            memset(by, 0, FQB);
            if (p.inf) by[FQB - 1] = 1;
            else { fe_bytes_be(p.x, host::FQ, bx); fe_bytes_be(p.y, host::FQ, by); }
            eth_append(bx, FQB);
            eth_append(by, FQB);
            return;
        }
        uint8_t b[2 * FQB + 1];                                    // GroupAffine::write: x || y || infinity (65 / 97 bytes)
        memset(b, 0, sizeof b);
        if (p.inf) { b[FQB] = 1; b[2 * FQB] = 1; }
        else { fe_bytes(p.x, host::FQ, b); fe_bytes(p.y, host::FQ, b + FQB); }
        append_message(label, b, sizeof b);
    }
    Fe challenge_scalar(const char *label) {
        uint8_t b[32];
        memset(b, 0, 32);
        if (kind == 1) {                                           // transcript.rs:76-89
            uint8_t d[69], dig[32];
            d[0] = 2;
            memcpy(d + 1, st0, 32);
            memcpy(d + 33, st1, 32);
            d[65] = (uint8_t)(counter >> 24); d[66] = (uint8_t)(counter >> 16); d[67] = (uint8_t)(counter >> 8); d[68] = (uint8_t)counter;
            ++counter;
            keccak256(d, 69, dig);
            for (int i = 0; i < 32; ++i) b[i] = dig[31 - i];
            b[31] &= 0x1f;
        } else {                                                   // 31 squeezed bytes -> from_random_bytes
            s.meta_ad((const uint8_t *)label, strlen(label), false);
            uint8_t len[4] = {31, 0, 0, 0};
            s.meta_ad(len, 4, true);
            s.prf(b, 31);
        }
        Fe canon, r2;
        memcpy(canon.l, b, 32);
        memcpy(r2.l, host::FR.r2, 32);
        return host::mul(canon, r2, host::FR);                      // canonical -> Montgomery
    }
};

// GroupAffine::serialize (ark-ec 0.3 / ark-serialize 0.3 SWFlags): x LE, bit 6 of the last byte = infinity,
// bit 7 = (y > -y)
void g1_compressed(const Pt &p, uint8_t *out /* FQB bytes */) {
    memset(out, 0, FQB);
    if (p.inf) { out[FQB - 1] |= 1 << 6; return; }
    fe_bytes(p.x, host::FQ, out);
    Fq one;
    memset(one.l, 0, sizeof one.l);
    one.l[0] = 1;
    Fq y = host::mul(p.y, one, host::FQ), ny;
    host::sub_raw<host::FQ_L>(ny.l, host::FQ.p, y.l);
    if (host::is_zero(y)) memset(ny.l, 0, sizeof ny.l);
    bool gt = false;
    for (int i = host::FQ_L - 1; i >= 0; --i) if (y.l[i] != ny.l[i]) { gt = y.l[i] > ny.l[i]; break; }
    if (gt) out[FQB - 1] |= 1 << 7;
}

// ================================================================================================ small helpers
inline Fe fe_from(const uint64_t *p) { Fe f; memcpy(f.l, p, 32); return f; }
inline Fq fq_from(const uint64_t *p) { Fq f; memcpy(f.l, p, FQB); return f; }
inline Fe FR_ONE() { return host::one(host::FR); }
inline Fe fadd(const Fe &a, const Fe &b) { return host::add(a, b, host::FR); }
inline Fe fsub(const Fe &a, const Fe &b) { return host::sub(a, b, host::FR); }
inline Fe fmul(const Fe &a, const Fe &b) { return host::mul(a, b, host::FR); }
inline Fe fneg(const Fe &a) { Fe z = {{0, 0, 0, 0}}; return host::sub(z, a, host::FR); }
template <int L> inline bool feq(const host::FeT<L> &a, const host::FeT<L> &b) { return memcmp(a.l, b.l, 8 * L) == 0; }

}  // namespace
