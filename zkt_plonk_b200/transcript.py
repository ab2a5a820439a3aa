"""Merlin transcript (merlin 3.0 over STROBE-128 / Keccak-f[1600]) and the reference's `TranscriptProtocol` on top;
also the reference's alternative `EthereumTranscript` (gadgets/src/transcript.rs:8-90, Keccak-256 over the same
permutation), which is pinned by the reference's own known-answer test (gadgets/src/transcript.rs:100-127).

Host-side, bytes-scale (a few dozen absorb/squeeze calls per proof), so it stays in Python.  Restates
plonk-core/src/transcript.rs:49-109 (`MerlinTranscript`): scalars are appended as `ToBytes::write` output
(32-byte little-endian canonical integers), commitments as x || y || infinity (65 bytes), challenges are 31
squeezed bytes read as a little-endian integer (`F::from_random_bytes`).  merlin and strobe are crates.io
dependencies of the reference (plonk-core/Cargo.toml:30) that are not vendored; the construction below follows
the published STROBE "lite" used by merlin and is pinned by merlin's own published test vector
(tests/test_transcript.py).
"""
from . import field

_MASK = (1 << 64) - 1
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(v, r):
    r %= 64
    return ((v << r) | (v >> (64 - r))) & _MASK if r else v


def keccak_f1600(state):
    """state: bytearray(200), permuted in place."""
    a = [[int.from_bytes(state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8] = (a[x][y] & _MASK).to_bytes(8, "little")


_R = 166
_FLAG_I, _FLAG_A, _FLAG_C, _FLAG_T, _FLAG_M, _FLAG_K = 1, 2, 4, 8, 16, 32


class Strobe128:
    def __init__(self, protocol_label):
        st = bytearray(200)
        st[0:6] = bytes([1, _R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[_R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos, self.pos_begin = 0, 0

    def _absorb(self, data):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == _R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == _R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert not (flags & _FLAG_T)
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if (flags & (_FLAG_C | _FLAG_K)) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(_FLAG_M | _FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(_FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin_op(_FLAG_I | _FLAG_A | _FLAG_C, more)
        return self._squeeze(n)


class Merlin:
    """merlin::Transcript."""

    def __init__(self, label):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, message):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def append_u64(self, label, x):
        self.append_message(label, int(x).to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(int(n).to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


def fr_bytes(x):
    """Fr ToBytes::write / CanonicalSerialize: 32 bytes little endian, canonical."""
    return int(x % field.R_MOD).to_bytes(32, "little")


def g1_write_bytes(pt):
    """GroupAffine ToBytes::write: x || y || infinity (65 bytes on BN254, 97 on the BLS12 curves).  pt: (x, y) canonical ints or
    None (identity, which arkworks stores as (0, 1, true))."""
    nb = field.FQ_BYTES
    if pt is None:
        return (0).to_bytes(nb, "little") + (1).to_bytes(nb, "little") + b"\x01"
    return int(pt[0]).to_bytes(nb, "little") + int(pt[1]).to_bytes(nb, "little") + b"\x00"


class MerlinTranscript:
    """TranscriptProtocol<Fr, kzg10::Commitment<Bn254>> for MerlinTranscript (transcript.rs:49-109)."""

    def __init__(self, label):
        self.t = Merlin(label.encode() if isinstance(label, str) else label)

    def append_u64(self, label, item):
        self.t.append_u64(label.encode(), item)

    def append_scalar(self, label, item):
        self.t.append_message(label.encode(), fr_bytes(item))

    def append_scalars(self, label, items):
        self.t.append_message(label.encode(), b"".join(fr_bytes(x) for x in items))

    def append_commitment(self, label, pt):
        self.t.append_message(label.encode(), g1_write_bytes(pt))

    def append_commitments(self, label, pts):
        self.t.append_message(label.encode(), b"".join(g1_write_bytes(p) for p in pts))

    def challenge_scalar(self, label):
        num_bytes = (field.R_MOD.bit_length() + 7) // 8 - 1  # (F::size_in_bits() + 7) / 8 - 1 = 31 on all three curves
        return int.from_bytes(self.t.challenge_bytes(label.encode(), num_bytes), "little")   # from_random_bytes


def keccak256(data):
    """sha3::Keccak256 (original Keccak padding 0x01 .. 0x80, rate 136), as gadgets/src/transcript.rs:4 imports it."""
    rate, st = 136, bytearray(200)
    data = bytes(data)
    off = 0
    while len(data) - off >= rate:
        for i in range(rate):
            st[i] ^= data[off + i]
        keccak_f1600(st)
        off += rate
    for i, byte in enumerate(data[off:]):
        st[i] ^= byte
    st[len(data) - off] ^= 0x01
    st[rate - 1] ^= 0x80
    keccak_f1600(st)
    return bytes(st[:32])


class EthereumTranscript:
    """TranscriptProtocol<Fr, KZG10Commitment<Bn254>> for EthereumTranscript (gadgets/src/transcript.rs:8-90): labels are
    ignored, two 32-byte states are re-hashed with domain bytes 0 / 1 on every item, items are big-endian, a
    challenge is Keccak-256(2 || state_0 || state_1 || counter_be32) read big-endian with the top three bits cleared."""

    def __init__(self, label=None):
        self.state_0, self.state_1, self.counter = bytes(32), bytes(32), 0

    def _append(self, item):
        body = self.state_0 + self.state_1 + bytes(item)
        self.state_0, self.state_1 = keccak256(b"\x00" + body), keccak256(b"\x01" + body)

    def append_u64(self, label, item):
        self._append(int(item).to_bytes(8, "big"))

    def append_scalar(self, label, item):
        self._append(int(item % field.R_MOD).to_bytes(32, "big"))

    def append_scalars(self, label, items):
        for x in items:
            self.append_scalar(label, x)

    def append_commitment(self, label, pt):
        x, y = (0, 1) if pt is None else pt               # arkworks' zero is (0, 1, true)
        self._append(int(x).to_bytes(32, "big"))
        self._append(int(y).to_bytes(32, "big"))

    def append_commitments(self, label, pts):
        for p in pts:
            self.append_commitment(label, p)

    def challenge_scalar(self, label):
        digest = keccak256(b"\x02" + self.state_0 + self.state_1 + self.counter.to_bytes(4, "big"))
        self.counter += 1
        return int.from_bytes(digest, "big") & ((1 << 253) - 1)   # query.reverse(); query[31] &= 0x1f; Fr::read


TRANSCRIPTS = {"merlin": (0, MerlinTranscript), "ethereum": (1, EthereumTranscript)}
