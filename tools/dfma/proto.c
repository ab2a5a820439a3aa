// CPU prototype of a Montgomery product built on FP64 fused multiply-adds (52-bit limbs, R = 2^260).
// Every floating-point operation below is exact by construction (see the comments), so the C `fma` under
// FE_TOWARDZERO reproduces what `fma.rz.f64` computes on the GPU.  Prints test vectors for tools/dfma/check.py.
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef union { double d; uint64_t u; } du;
static const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52, TWO52 = 0x1p52;
#define MASK52 ((1ULL << 52) - 1)
#define OFF_HI (0x467ULL << 52)
#define OFF_LO (0x433ULL << 52)

static uint64_t P52[5], PINV52;   // modulus limbs, -p^-1 mod 2^52

static inline uint64_t bits(double x) { du t; t.d = x; return t.u; }
static inline double from52(uint64_t x) { du t; t.u = x | OFF_LO; return t.d - TWO52; }   // exact for x < 2^52

static void mont_mul52(const uint64_t a[5], const uint64_t b[5], uint64_t r[5]) {
    double ad[5], bd[5], pd[5];
    for (int i = 0; i < 5; ++i) { ad[i] = from52(a[i]); bd[i] = from52(b[i]); pd[i] = from52(P52[i]); }
    const double pinv = from52(PINV52);
    uint64_t col[11];
    for (int k = 0; k < 11; ++k) {              // pre-subtract the exponent patterns of every term the column will receive
        int nlo = 0, nhi = 0;
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { if (i + j == k) ++nlo; if (i + j + 1 == k) ++nhi; }
        col[k] = 0 - 2 * ((uint64_t)nlo * OFF_LO + (uint64_t)nhi * OFF_HI);
    }
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            double hi = fma(ad[i], bd[j], C1);          // 2^104 + floor(ab / 2^52) * 2^52   (rz, ab < 2^104)
            double sub = C2 - hi;                        // (1 - h) * 2^52: exact
            double lo = fma(ad[i], bd[j], sub);          // 2^52 + (ab mod 2^52): exact
            col[i + j + 1] += bits(hi);
            col[i + j] += bits(lo);
        }
    for (int i = 0; i < 5; ++i) {
        double td = from52(col[i] & MASK52);
        double h2 = fma(td, pinv, C1), s2 = C2 - h2, l2 = fma(td, pinv, s2);
        double qd = l2 - TWO52;                          // q = t * pinv mod 2^52
        for (int j = 0; j < 5; ++j) {
            double hi = fma(qd, pd[j], C1), sub = C2 - hi, lo = fma(qd, pd[j], sub);
            col[i + j + 1] += bits(hi);
            col[i + j] += bits(lo);
        }
        col[i + 1] += col[i] >> 52;                      // low 52 bits of col[i] are zero now
    }
    uint64_t t[5], carry = 0;
    for (int k = 0; k < 5; ++k) { uint64_t v = col[5 + k] + carry; t[k] = v & MASK52; carry = v >> 52; }
    // t < 2p: one conditional subtraction
    uint64_t d[5]; int64_t borrow = 0;
    for (int k = 0; k < 5; ++k) { int64_t v = (int64_t)t[k] - (int64_t)P52[k] + borrow; d[k] = (uint64_t)v & MASK52; borrow = v >> 52; }
    for (int k = 0; k < 5; ++k) r[k] = borrow ? t[k] : d[k];
}

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(int argc, char **argv) {
    fesetround(FE_TOWARDZERO);
    // argv: p limbs (5 hex), pinv (hex), count
    for (int i = 0; i < 5; ++i) sscanf(argv[1 + i], "%lx", &P52[i]);
    sscanf(argv[6], "%lx", &PINV52);
    int count = atoi(argv[7]);
    for (int n = 0; n < count; ++n) {
        uint64_t a[5], b[5], r[5];
        for (int k = 0; k < 5; ++k) { a[k] = rnd() & MASK52; b[k] = rnd() & MASK52; }
        a[4] &= (1ULL << 45) - 1; b[4] &= (1ULL << 45) - 1;       // < 2^253 < p
        if (n == 0) for (int k = 0; k < 5; ++k) { a[k] = P52[k]; b[k] = P52[k]; }   // edge: p - 1
        if (n == 0) { a[0] -= 1; b[0] -= 1; }
        if (n == 1) memset(a, 0, sizeof a);
        mont_mul52(a, b, r);
        for (int k = 0; k < 5; ++k) printf("%lx ", a[k]);
        for (int k = 0; k < 5; ++k) printf("%lx ", b[k]);
        for (int k = 0; k < 5; ++k) printf("%lx ", r[k]);
        printf("\n");
    }
    return 0;
}
