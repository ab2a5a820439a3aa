"""CPU-only: the two transcripts of the reference, in both host implementations (Python mirror, C++ round driver).

`EthereumTranscript` is checked against the ONE byte-level known-answer test the reference repository holds
(gadgets/src/transcript.rs:100-127): three challenges after append_u64(1), append_scalar(2), append_commitment((3, 4)).
Both transcripts run on the same Keccak-f[1600] restatement, so this vector also pins the permutation under the
Merlin transcript (whose framing is pinned by merlin's published vector in tests/test_prover_cpu.py)."""
import ctypes
import hashlib

import numpy as np

from zkt_plonk_b200 import _lib, field, prover
from zkt_plonk_b200.transcript import EthereumTranscript, MerlinTranscript, keccak256

# gadgets/src/transcript.rs:107-126
REF_KAT = ["0f9d11cec4f06b0d18060cde3db4196495ddfbb096108951446fc8a1d45f4b59",
           "0f4dccb919a5dba2dd010a562ba45b4551291f5e565706536e78b24ac8b5c64d",
           "1b5bf46adfcd1dd4f9ac7166586cf83f261192bc4b83fdda30ddee22f9054c1f"]
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def test_keccak256_known_answers():
    """Public Keccak-256 vectors (empty string, "abc") and the rate boundary against SHA3's sibling padding."""
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    for n in (135, 136, 137, 272, 300):                     # multi-block absorb: same sponge as hashlib's SHA3-256 except
        msg = bytes(range(256)) * 2                          # for the domain byte, so check the block handling differs
        assert keccak256(msg[:n]) != hashlib.sha3_256(msg[:n]).digest() and len(keccak256(msg[:n])) == 32


def test_ethereum_transcript_reference_kat_python():
    t = EthereumTranscript("test")
    t.append_u64("a", 1)
    assert t.challenge_scalar("a").to_bytes(32, "big").hex() == REF_KAT[0]
    t.append_scalar("b", 2)
    assert t.challenge_scalar("b").to_bytes(32, "big").hex() == REF_KAT[1]
    t.append_commitment("c", (3, 4))
    assert t.challenge_scalar("c").to_bytes(32, "big").hex() == REF_KAT[2]


def _run_native(kind, script):
    """script: list of ("u64", v) / ("scalar", int) / ("commit", (x, y) or None) / ("challenge",)."""
    lib = _lib.lib()
    ops = np.zeros(len(script), dtype=np.uint8)
    args = np.zeros((len(script), 9), dtype=np.uint64)
    R = 1 << 256
    n_ch = 0
    for i, item in enumerate(script):
        if item[0] == "u64":
            ops[i], args[i, 0] = 0, item[1]
        elif item[0] == "scalar":
            ops[i] = 1
            args[i, :4] = prover.ints_to_mont_array([item[1]])[0]
        elif item[0] == "commit":
            ops[i] = 2
            if item[1] is None:
                args[i, 8] = 1
            else:
                for j, v in enumerate(item[1]):
                    m = v * R % Q
                    args[i, 4 * j: 4 * j + 4] = [(m >> (64 * k)) & (2**64 - 1) for k in range(4)]
        else:
            ops[i] = 3
            n_ch += 1
    out = np.zeros((max(n_ch, 1), 32), dtype=np.uint8)
    rc = lib.zkb_test_transcript(kind, ops.ctypes.data_as(ctypes.c_void_p), len(script), args.ctypes.data_as(ctypes.c_void_p),
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return [int.from_bytes(out[k].tobytes(), "little") for k in range(n_ch)]


def test_ethereum_transcript_reference_kat_native():
    got = _run_native(1, [("u64", 1), ("challenge",), ("scalar", 2), ("challenge",), ("commit", (3, 4)), ("challenge",)])
    assert [g.to_bytes(32, "big").hex() for g in got] == REF_KAT


def test_native_transcripts_match_python_on_a_longer_script():
    """Same script through the C++ and the Python transcripts, both kinds: u64s, scalars near the modulus, points, the
    identity commitment (arkworks' (0, 1, true)), and challenges interleaved."""
    script = [("u64", 1 << 40), ("commit", (1, 2)), ("challenge",), ("scalar", field.R_MOD - 1), ("scalar", 0), ("commit", None),
              ("challenge",), ("challenge",), ("commit", (Q - 1, Q - 2)), ("u64", 0), ("challenge",)]
    for kind, cls in ((0, MerlinTranscript), (1, EthereumTranscript)):
        t, exp = cls("test"), []
        for item in script:
            if item[0] == "u64":
                t.append_u64("a", item[1])
            elif item[0] == "scalar":
                t.append_scalar("b", item[1])
            elif item[0] == "commit":
                t.append_commitment("c", item[1])
            else:
                exp.append(t.challenge_scalar("a"))
        assert _run_native(kind, script) == exp, kind
