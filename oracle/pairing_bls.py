"""Optimal ate pairing on BLS12-381 and BLS12-377 in plain Python integers.  TEST INFRASTRUCTURE ONLY.

The restated verifier needs PC::check = SonicKZG10::check, a product of two pairings (ark-poly-commit 0.3 over ark-ec 0.3 /
ark-bls12-381 / ark-bls12-377 0.3: crates.io dependencies of the reference, un-vendored); oracle/pairing.py does this for
BN254, this file for the two curves plonk-core/src/plonk.rs:226-254 also runs.  As there, the pairing is restated in its
simplest exact form -- any correct pairing gives the same accept / reject decisions:

  Fq2  = Fq[i] / (i^2 - beta)            beta = -1 (BLS12-381), -5 (BLS12-377)
  Fq12 = Fq[w] / (w^12 - A w^6 - B)      w^6 = xi = xi0 + i:  A = 2 xi0, B = beta - xi0^2   (flat: 12 coefficients in Fq)
  G2 on the twist E': y^2 = x^3 + b xi (M-type, BLS12-381) or b / xi (D-type, BLS12-377), affine Fq2 coordinates
  untwist: (x', y') -> (x' / w^2, y' / w^3) (M) or (x' w^2, y' w^3) (D); lines are multiplied by w^3 on the M-type twist
  (an element of Fq4, killed by the final exponentiation)
  Miller loop over |x| (x = -0xd201000000010000 / 0x8508c00000000001), conjugation for x < 0, final power (q^12 - 1) / r.

Pinned by tests/test_pairing_bls.py: bilinearity in both arguments, non-degeneracy, e(P, Q)^r = 1, the parameter identities
r = x^4 - x^2 + 1 and q = (x - 1)^2 r / 3 + x, and the standard G2 generator of BLS12-381.
"""

CURVES = {
    "bls12_381": dict(
        q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
        r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
        x=-0xd201000000010000, b=4, beta=-1, xi0=1, twist="M",
        g1=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
            0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
        # the standard generator of G2 (x = c0 + c1 i, y = c0 + c1 i)
        g2=((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
             0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
            (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
             0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))),
    "bls12_377": dict(
        q=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
        r=0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
        x=0x8508c00000000001, b=1, beta=-5, xi0=0, twist="D",
        g1=(0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
            0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6),
        # a point of order r on the twist, derived here (x = 2 + i, cofactor cleared): KZG10::setup draws h at random anyway
        g2=((0x6f72205595a839df693176b247c2fa251f7e02a29061e50540dc9e1c2bf1957bf1bab2288c257c2cb36b58f2418bc9,
             0x138c24b2b4e17888beed0a9802aac837cdea39890effe00072f754ecb0152dd6cb524f281298966dbaeca23d3e462b8),
            (0x16235fdea6c3faf2a83d3730f6ab2c033ef6c2739002946f7dc48e4688bca1af1c9b417d58220817e0dc644b5e7d916,
             0x707ac6cc7d192827fc54eb83267f3bed8511bd3c74f63a1ea75eabb66476769c8786f2af2a75166f33142379b4963c))),
}


class Pairing:
    def __init__(self, curve):
        c = CURVES[curve]
        self.curve, self.q, self.r, self.x, self.b, self.beta, self.xi0, self.twist = (
            curve, c["q"], c["r"], c["x"], c["b"], c["beta"], c["xi0"], c["twist"])
        self.g1, self.g2 = c["g1"], c["g2"]
        self.A, self.B = 2 * self.xi0 % self.q, (self.beta - self.xi0 * self.xi0) % self.q
        self.final_exp = (self.q ** 12 - 1) // self.r
        assert (self.q ** 12 - 1) % self.r == 0

    # ---- Fq2
    def f2_mul(self, a, b):
        q = self.q
        return ((a[0] * b[0] + self.beta * a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)

    def f2_sub(self, a, b):
        return ((a[0] - b[0]) % self.q, (a[1] - b[1]) % self.q)

    def f2_add(self, a, b):
        return ((a[0] + b[0]) % self.q, (a[1] + b[1]) % self.q)

    def f2_inv(self, a):
        q = self.q
        d = pow((a[0] * a[0] - self.beta * a[1] * a[1]) % q, -1, q)
        return (a[0] * d % q, -a[1] * d % q)

    def twist_b(self):
        xi = (self.xi0, 1)
        return self.f2_mul((self.b, 0), xi if self.twist == "M" else self.f2_inv(xi))

    def g2_on_curve(self, Q):
        if Q is None:
            return True
        x, y = Q
        return self.f2_sub(self.f2_mul(y, y), self.f2_add(self.f2_mul(self.f2_mul(x, x), x), self.twist_b())) == (0, 0)

    def g2_add(self, P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if self.f2_add(P[1], Q[1]) == (0, 0):
                return None
            lam = self.f2_mul(self.f2_mul((3, 0), self.f2_mul(P[0], P[0])), self.f2_inv(self.f2_add(P[1], P[1])))
        else:
            lam = self.f2_mul(self.f2_sub(Q[1], P[1]), self.f2_inv(self.f2_sub(Q[0], P[0])))
        x = self.f2_sub(self.f2_sub(self.f2_mul(lam, lam), P[0]), Q[0])
        return (x, self.f2_sub(self.f2_mul(lam, self.f2_sub(P[0], x)), P[1]))

    def g2_mul(self, k, P):
        R = None
        k %= self.r
        while k:
            if k & 1:
                R = self.g2_add(R, P)
            P = self.g2_add(P, P)
            k >>= 1
        return R

    # ---- G1 (affine over Fq)
    def g1_add(self, P, Q):
        q = self.q
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % q == 0:
                return None
            lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, q) % q
        else:
            lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, q) % q
        x = (lam * lam - P[0] - Q[0]) % q
        return (x, (lam * (P[0] - x) - P[1]) % q)

    def g1_mul(self, k, P):
        R = None
        k %= self.r
        while k:
            if k & 1:
                R = self.g1_add(R, P)
            P = self.g1_add(P, P)
            k >>= 1
        return R

    # ---- Fq12 = Fq[w] / (w^12 - A w^6 - B), 12 coefficients
    def f12_one(self):
        return [1] + [0] * 11

    def f12_mul(self, a, b):
        q = self.q
        t = [0] * 23
        for i, ai in enumerate(a):
            if ai:
                for j, bj in enumerate(b):
                    t[i + j] += ai * bj
        for k in range(22, 11, -1):
            v = t[k] % q
            if v:
                t[k - 6] += self.A * v
                t[k - 12] += self.B * v
        return [v % q for v in t[:12]]

    def f12_pow(self, a, e):
        r = self.f12_one()
        while e:
            if e & 1:
                r = self.f12_mul(r, a)
            a = self.f12_mul(a, a)
            e >>= 1
        return r

    def embed(self, a):
        """Fq2 -> (coefficient at w^0, coefficient at w^6): a0 + a1 i with i = w^6 - xi0"""
        return ((a[0] - self.xi0 * a[1]) % self.q, a[1] % self.q)

    def line(self, lam, xr, yr, P):
        """the line through the untwisted (xr, yr) with slope lam (on the twist), evaluated at P in E(Fq), as an Fq12 element"""
        q = self.q
        l = [0] * 12
        c = self.embed(self.f2_sub(yr, self.f2_mul(lam, xr)))
        s = self.embed(lam)
        if self.twist == "D":                                      # -yP + lam xP w + (yr - lam xr) w^3
            l[0] = -P[1] % q
            l[1], l[7] = s[0] * P[0] % q, s[1] * P[0] % q
            l[3], l[9] = c
        else:                                                      # times w^3: -yP w^3 + lam xP w^2 + (yr - lam xr)
            l[3] = -P[1] % q
            l[2], l[8] = s[0] * P[0] % q, s[1] * P[0] % q
            l[0], l[6] = c
        return l

    def miller_loop(self, P, Q):
        if P is None or Q is None:
            return self.f12_one()
        f, T = self.f12_one(), Q
        bits = bin(abs(self.x))[3:]
        for bit in bits:
            lam = self.f2_mul(self.f2_mul((3, 0), self.f2_mul(T[0], T[0])), self.f2_inv(self.f2_add(T[1], T[1])))
            f = self.f12_mul(self.f12_mul(f, f), self.line(lam, T[0], T[1], P))
            T = self.g2_add(T, T)
            if bit == "1":
                lam = self.f2_mul(self.f2_sub(Q[1], T[1]), self.f2_inv(self.f2_sub(Q[0], T[0])))
                f = self.f12_mul(f, self.line(lam, T[0], T[1], P))
                T = self.g2_add(T, Q)
        if self.x < 0:                                             # f_{-|x|} = 1 / f_{|x|} up to factors the final power kills:
            f = [v if k % 2 == 0 else -v % self.q for k, v in enumerate(f)]   # the q^6 Frobenius w -> -w inverts a unitary element
        return f

    def pairing(self, P, Q):
        return self.f12_pow(self.miller_loop(P, Q), self.final_exp)

    def pairing_product_is_one(self, pairs):
        f = self.f12_one()
        for P, Q in pairs:
            f = self.f12_mul(f, self.miller_loop(P, Q))
        return self.f12_pow(f, self.final_exp) == self.f12_one()
