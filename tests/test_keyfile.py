"""CPU-only: the reference CLI's key files (ck / pk / vk, ark-serialize 0.3 serialize_unchecked; bin/src/parser.rs:5-29)
through the host-side C ABI (csrc/keyfile.cu) against the independent Python restatement of the derive rules
(oracle/arkser.py): both directions byte for byte, plus the malformed inputs a loader must refuse."""
import random

import numpy as np
import pytest

from oracle import arkser, plonk_ref, pyref
from zkt_plonk_b200 import _lib, keyfile, prover, synthetic

P = prover.P
Q = arkser.Q
TAU = 0x1D9E5F1B2C3A49587766554433221100FFEEDDCCBBAA99887766554433221101 % P


def _pts_to_array(pts):
    out = np.zeros((len(pts), 8), dtype=np.uint64)
    for i, pt in enumerate(pts):
        if pt is None:
            continue
        for j, v in enumerate(pt):
            m = (v << 256) % Q
            out[i, 4 * j: 4 * j + 4] = [(m >> (64 * k)) & (2**64 - 1) for k in range(4)]
    return out


def _array_to_pts(arr):
    return [prover.point_to_ints(row, not row.any()) for row in arr]


@pytest.fixture(scope="module")
def keys():
    """A small circuit's keys from the oracle backend: polynomials as canonical ints, commitments as points."""
    circ = synthetic.make_circuit(4, seed=2, table_size=4)
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    pk, vk = prover.setup(be, circ)
    polys = {name: prover.mont_array_to_ints(pk.polys[name].data[: pk.polys[name].len]) for name in arkser.PK_ORDER}
    return circ, pk, vk, polys


def test_prover_key_file_both_directions(tmp_path, keys):
    circ, pk, vk, polys = keys
    ref_bytes = arkser.prover_key(polys)
    # reference-format bytes -> C++ reader
    path = tmp_path / "pk"
    path.write_bytes(ref_bytes)
    got = keyfile.pk_read(path)
    for name in arkser.PK_ORDER:
        assert prover.mont_array_to_ints(got[name]) == polys[name], name
    # C++ writer -> identical bytes (also when the caller hands over zero-padded coefficient arrays)
    out = tmp_path / "pk2"
    padded = {name: np.concatenate([got[name], np.zeros((3, 4), dtype=np.uint64)]) for name in arkser.PK_ORDER}
    keyfile.pk_write(out, padded)
    assert out.read_bytes() == ref_bytes
    assert arkser.parse_prover_key(out.read_bytes()) == polys
    # q_table's polynomial of this circuit is the mask of lookup/table.rs:42-48
    assert len(polys["q_table"]) <= circ.n and len(ref_bytes) == sum(len(v) for v in polys.values()) * 32 + sum(
        8 + len(n) + 8 + 2 for n in arkser.PK_ORDER)


def test_verifier_key_file_both_directions(tmp_path, keys):
    circ, pk, vk, polys = keys
    ref_bytes = arkser.verifier_key(vk.n, vk.pi_roots, vk.commits)
    path = tmp_path / "vk"
    path.write_bytes(ref_bytes)
    n, roots, xy, inf = keyfile.vk_read(path)
    assert n == vk.n and prover.mont_array_to_ints(roots) == vk.pi_roots
    assert dict(zip(arkser.PK_ORDER, _array_to_pts(xy))) == vk.commits and not any(inf)
    out = tmp_path / "vk2"
    keyfile.vk_write(out, n, roots, xy, inf)
    assert out.read_bytes() == ref_bytes
    # an identity commitment (an all-zero selector commits to the identity): (0, 1) + flag on disk, (0, 0) in memory
    commits = dict(vk.commits)
    commits["q_c"] = None
    path.write_bytes(arkser.verifier_key(vk.n, [], commits))
    n, roots, xy, inf = keyfile.vk_read(path)
    assert roots.shape[0] == 0 and inf[4] and not xy[4].any() and sum(inf) == 1
    keyfile.vk_write(out, n, roots, xy)                     # the flag is implied by (0, 0)
    assert out.read_bytes() == path.read_bytes()


def test_committer_key_file(tmp_path):
    rnd = random.Random(4)
    pts = [pyref.g1_mul(rnd.randrange(1, P), pyref.G1_GEN) for _ in range(37)] + [None]
    gamma = pts[:2]
    ref_bytes = arkser.committer_key(pts, gamma, 1 << 20)
    path = tmp_path / "ck"
    path.write_bytes(ref_bytes)
    assert keyfile.ck_info(path) == (38, 1 << 20)
    assert _array_to_pts(keyfile.ck_read(path)) == pts
    assert _array_to_pts(keyfile.ck_read(path, 5, 9)) == pts[5:14]
    assert keyfile.ck_read(path, 38, 0).shape == (0, 8)
    with pytest.raises(_lib.ZkbError):
        keyfile.ck_read(path, 30, 9)                        # past the end
    out = tmp_path / "ck2"
    keyfile.ck_write(out, _pts_to_array(pts), _pts_to_array(gamma), 1 << 20)
    assert out.read_bytes() == ref_bytes
    assert arkser.parse_committer_key(out.read_bytes()) == (pts, gamma, 1 << 20)
    # many points: the threaded conversion path
    many = [pts[i % 37] for i in range(5000)]
    path.write_bytes(arkser.committer_key(many, gamma, 4999))
    assert _array_to_pts(keyfile.ck_read(path, 4000, 1000)) == many[4000:]


def test_malformed_files_are_refused(tmp_path, keys):
    circ, pk, vk, polys = keys
    good_pk = arkser.prover_key(polys)
    good_vk = arkser.verifier_key(vk.n, vk.pi_roots, vk.commits)
    path = tmp_path / "bad"

    def bad_pk(data):
        path.write_bytes(data)
        with pytest.raises(_lib.ZkbError):
            keyfile.pk_read(path)

    def bad_vk(data):
        path.write_bytes(data)
        with pytest.raises(_lib.ZkbError):
            keyfile.vk_read(path)

    bad_pk(good_pk[:-1])                                    # truncated
    bad_pk(good_pk + b"\x00")                               # trailing bytes
    bad_pk(good_pk.replace(b"q_l", b"q_x", 1))              # wrong label
    bad_pk(b"")
    first = 8 + 3 + 8                                       # label length, "q_m", coefficient count
    bad_pk(good_pk[:first] + b"\xff" * 32 + good_pk[first + 32:])          # coefficient >= r
    with_bound = arkser.vec(b"q_m", lambda c: bytes([c])) + arkser.vec(polys["q_m"], arkser.fe) + b"\x01" + arkser.u64(5) + b"\x00"
    bad_pk(with_bound + good_pk[len(arkser.labeled_polynomial("q_m", polys["q_m"])):])   # Some(degree_bound): not a plonk key
    bad_vk(good_vk[:-5])
    bad_vk(good_vk + b"\x00")
    bad_vk(good_vk[:-32] + b"\xff" * 31 + b"\x3f")          # y >= q once the flag bits are masked
    with pytest.raises(_lib.ZkbError):
        keyfile.pk_read(tmp_path / "missing")
    with pytest.raises(_lib.ZkbError):
        keyfile.ck_info(tmp_path / "missing")
    # a committer key with enforced degree bounds is not what `compile` writes
    pts = [pyref.G1_GEN]
    path.write_bytes(arkser.vec(pts, arkser.g1) + arkser.vec(pts, arkser.g1) + b"\x01" + arkser.vec(pts, arkser.g1) + b"\x00\x00" + arkser.u64(1))
    with pytest.raises(_lib.ZkbError):
        keyfile.ck_info(path)


def test_cvk_file_head_feeds_the_verifier(tmp_path, keys):
    """The G2 half of the commitment verifier key, read from a cvk-shaped file, verifies a proof made on the same SRS."""
    from oracle import pairing as pr
    from zkt_plonk_b200 import verifier
    circ, pk, vk, polys = keys
    h, beta_h = plonk_ref.make_cvk(TAU)
    gamma_g = pyref.g1_mul(77, pyref.G1_GEN)
    path = tmp_path / "cvk"
    path.write_bytes(arkser.sonic_verifier_key(pyref.G1_GEN, gamma_g, h, beta_h, 4 * circ.n, 1 << 20))
    g_arr, gg_arr, h_arr, bh_arr = keyfile.cvk_read(path)
    assert _array_to_pts(np.stack([g_arr, gg_arr])) == [pyref.G1_GEN, gamma_g]
    assert np.array_equal(h_arr, verifier.g2_array(h)) and np.array_equal(bh_arr, verifier.g2_array(beta_h))
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    raw = prover.prove(be, pk, vk, circ, list(range(3, 22))).to_bytes()
    assert verifier.verify(vk, raw, list(circ.pi.values()), (h, beta_h)) == 0
    assert verifier.verify(vk, raw, list(circ.pi.values()), (h_arr, bh_arr)) == 0          # straight from the file
    path.write_bytes(path.read_bytes()[:300])
    with pytest.raises(_lib.ZkbError):
        keyfile.cvk_read(path)


def test_readers_survive_mutated_files(tmp_path, keys):
    """Byte flips, truncations and length-field inflation of valid files: every reader returns an error or a parse --
    it never crashes, hangs or allocates by an attacker-chosen length (lengths are bounded before any allocation)."""
    circ, pk, vk, polys = keys
    rnd = random.Random(99)
    pts = [pyref.g1_mul(k + 1, pyref.G1_GEN) for k in range(6)]
    h, beta_h = plonk_ref.make_cvk(TAU)
    files = {
        "pk": (arkser.prover_key(polys), keyfile.pk_read),
        "vk": (arkser.verifier_key(vk.n, vk.pi_roots, vk.commits), keyfile.vk_read),
        "ck": (arkser.committer_key(pts, pts[:2], 5), keyfile.ck_read),
        "cvk": (arkser.sonic_verifier_key(pts[0], pts[1], h, beta_h, 64, 64), keyfile.cvk_read),
    }
    path = tmp_path / "mut"
    outcomes = {"ok": 0, "refused": 0}
    for name, (good, reader) in files.items():
        for trial in range(120):
            data = bytearray(good)
            kind = trial % 4
            if kind == 0:                                               # flip a few bytes
                for _ in range(rnd.randrange(1, 4)):
                    data[rnd.randrange(len(data))] ^= 1 << rnd.randrange(8)
            elif kind == 1:                                             # truncate
                data = data[: rnd.randrange(len(data))]
            elif kind == 2:                                             # huge length field somewhere plausible
                off = rnd.choice([0, 8, 11, 19]) if len(data) > 27 else 0
                data[off: off + 8] = (rnd.choice([1 << 40, (1 << 64) - 1, 1 << 31])).to_bytes(8, "little")
            else:                                                       # garbage tail
                data += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 40)))
            path.write_bytes(bytes(data))
            try:
                reader(path)
                outcomes["ok"] += 1
            except _lib.ZkbError:
                outcomes["refused"] += 1
    assert outcomes["refused"] > 100 and outcomes["ok"] + outcomes["refused"] == 480
