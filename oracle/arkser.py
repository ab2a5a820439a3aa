"""ark-serialize 0.3 `serialize_unchecked` of the reference's key types, in plain Python integers.  TEST INFRASTRUCTURE ONLY.

Independent restatement of what `#[derive(CanonicalSerialize)]` emits for the files `zkt compile` writes
(bin/src/parser.rs:16-29, bin/src/main.rs:106-112), used to cross-check csrc/keyfile.cu byte for byte:
  ProverKey<Fr>              plonk-core/src/proof_system/keys/mod.rs:29-40 (+ arithmetic.rs:20-32, permutation.rs:20-31,
                             lookup.rs:19-26); LabeledPolynomial / DensePolynomial from ark-poly-commit / ark-poly 0.3
  VerifierKey<Fr, KZG10>     keys/mod.rs:180-203
  sonic_pc::CommitterKey     ark-poly-commit 0.3 (un-vendored: field list recalled), as PC::trim(pp, 4n, 0, None) leaves it
PARITY UNPINNED against the Rust binary (no cargo here, no key file in the reference repository): the layout follows the
derive rules (fields in declaration order; usize as u64 LE; Vec / String = u64 length + items; Option = bool byte + value;
Fp256 = 32 B LE canonical; GroupAffine unchecked = x || y with the infinity flag in bit 6 of the last byte, identity
stored as (0, 1)).  Works on canonical Python ints; points are (x, y) tuples or None.
"""
import struct

Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ_BYTES = 32            # base-field element: 32 bytes on BN254 (the CLI's curve), 48 on BLS12-381 / BLS12-377 (use_curve)
PK_ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "sigma1", "sigma2", "sigma3", "q_lookup", "q_table")


def use_curve(name):
    """the same derive(CanonicalSerialize) layouts on another curve the reference instantiates (plonk.rs:226-254)"""
    global Q, R, FQ_BYTES
    from oracle import pyref
    R, Q = pyref._CURVES[name][0], pyref._CURVES[name][1]
    FQ_BYTES = 8 * ((Q.bit_length() + 63) // 64)


def u64(v):
    return struct.pack("<Q", v)


def fq(v):
    return int(v).to_bytes(FQ_BYTES, "little")


def fe(v):
    return int(v).to_bytes(32, "little")


def g1(pt):
    if pt is None:
        b = bytearray(fq(0) + fq(1))
        b[-1] |= 1 << 6
        return bytes(b)
    return fq(pt[0]) + fq(pt[1])


def vec(items, enc):
    return u64(len(items)) + b"".join(enc(x) for x in items)


def option(value, enc):
    return b"\x00" if value is None else b"\x01" + enc(value)


def labeled_polynomial(label, coeffs):
    coeffs = list(coeffs)
    while coeffs and coeffs[-1] == 0:                     # DensePolynomial::from_coefficients_vec
        coeffs.pop()
    return vec(label.encode(), lambda c: bytes([c])) + vec(coeffs, fe) + option(None, u64) + option(None, u64)


def prover_key(polys):
    """polys: {name: [canonical coefficient ints]}."""
    return b"".join(labeled_polynomial(name, polys[name]) for name in PK_ORDER)


def verifier_key(n, pi_roots, commits):
    """commits: {name: (x, y) | None}."""
    return u64(n) + vec(list(pi_roots), fe) + b"".join(g1(commits[name]) for name in PK_ORDER)


def committer_key(powers, gamma_powers, max_degree):
    return vec(list(powers), g1) + vec(list(gamma_powers), g1) + b"\x00\x00\x00" + u64(max_degree)


def g2(pt):
    """G2Affine, uncompressed: x.c0 x.c1 y.c0 y.c1, flags in the last byte."""
    if pt is None:
        b = bytearray(fq(0) + fq(0) + fq(1) + fq(0))
        b[-1] |= 1 << 6
        return bytes(b)
    (x0, x1), (y0, y1) = pt
    return fq(x0) + fq(x1) + fq(y0) + fq(y1)


def sonic_verifier_key(g, gamma_g, h, beta_h, supported_degree, max_degree, n_ell=3):
    """cvk as sonic_pc::VerifierKey derives it; the two G2Prepared fields (Vec of Fq2 triples + infinity flag) are filled
    with placeholder coefficients: no reader in this repository looks past beta_h."""
    prepared = vec([((1, 2), (3, 4), (5, 6))] * n_ell, lambda t: b"".join(fq(c) for pair in t for c in pair)) + b"\x00"
    return g1(g) + g1(gamma_g) + g2(h) + g2(beta_h) + prepared + prepared + option(None, u64) + u64(supported_degree) + u64(max_degree)


# ---- readers (for files written by the product)
class _Rd:
    def __init__(self, data):
        self.d, self.o = data, 0

    def take(self, n):
        assert self.o + n <= len(self.d), "truncated"
        out = self.d[self.o: self.o + n]
        self.o += n
        return out

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def fe(self, mod):
        v = int.from_bytes(self.take(32), "little")
        assert v < mod
        return v

    def g1(self):
        x = int.from_bytes(self.take(FQ_BYTES), "little")
        y = int.from_bytes(self.take(FQ_BYTES), "little")
        top = 8 * FQ_BYTES - 2
        inf = (y >> top) & 1
        y &= (1 << top) - 1
        return None if inf else (x, y)


def parse_prover_key(data):
    r, out = _Rd(data), {}
    for name in PK_ORDER:
        assert r.take(r.u64()) == name.encode()
        out[name] = [r.fe(R) for _ in range(r.u64())]
        assert r.take(2) == b"\x00\x00"
    assert r.o == len(data)
    return out


def parse_verifier_key(data):
    r = _Rd(data)
    n = r.u64()
    roots = [r.fe(R) for _ in range(r.u64())]
    commits = {name: r.g1() for name in PK_ORDER}
    assert r.o == len(data)
    return n, roots, commits


def parse_committer_key(data):
    r = _Rd(data)
    powers = [r.g1() for _ in range(r.u64())]
    gamma = [r.g1() for _ in range(r.u64())]
    assert r.take(3) == b"\x00\x00\x00"
    md = r.u64()
    assert r.o == len(data)
    return powers, gamma, md
