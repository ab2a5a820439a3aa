"""CPU-only: the BLS12-381 / BLS12-377 builds (the reference's test_full runs on both: plonk-core/src/plonk.rs:226-254).

Pins the per-curve C oracle (oracle/zkb_oracle.c compiled with -DZKO_CURVE=1/2) against the mathematical definitions in
Python big integers -- field operations, the O(n^2) DFT with the root arkworks derives, affine double-and-add -- and against
published constants (ark-bls12-381 / ark-bls12-377 0.3 TWO_ADIC_ROOT_OF_UNITY, the standard G1 generators); checks the
generated sources are current and that every build of the product library exports the whole C ABI.
"""
import ctypes
import importlib.util
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import cref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("gen_curves", os.path.join(ROOT, "tools", "gen_curves.py"))
gen_curves = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen_curves)
CURVES = {c["name"]: c for c in gen_curves.CURVES}
BLS = ["bls12_381", "bls12_377"]


def limbs(vals, nw):
    a = np.zeros((len(vals), nw), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(nw):
            a[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return a


def ints(a, nw):
    a = np.ascontiguousarray(a).reshape(-1, nw)
    return [sum(int(a[i, j]) << (64 * j) for j in range(nw)) for i in range(a.shape[0])]


def test_generated_sources_are_current(tmp_path):
    """curve_params.h, ff_wide.cuh and zko_curve_params.h are what their generators produce."""
    csrc = os.path.join(ROOT, "zkt_plonk_b200", "csrc")
    assert open(os.path.join(csrc, "curve_params.h")).read() == gen_curves.gen_params()
    assert open(os.path.join(csrc, "ff_wide.cuh")).read() == gen_curves.gen_wide()
    assert open(os.path.join(ROOT, "oracle", "zko_curve_params.h")).read() == gen_curves.gen_oracle_params()


@pytest.mark.parametrize("curve", BLS)
def test_published_constants(curve):
    c = CURVES[curve]
    gen_curves.check(c)                       # generator on the curve and of order r, GENERATOR a non-residue, GENERATOR^T == root
    r = c["r"]
    assert pow(c["root"], 1 << c["s"], r) == 1 and pow(c["root"], 1 << (c["s"] - 1), r) == r - 1
    assert {"bls12_381": (32, 7, 255, 381), "bls12_377": (47, 22, 253, 377)}[curve] == (c["s"], c["fr_gen"], r.bit_length(), c["q"].bit_length())
    o = cref.oracle(curve)
    assert (o.nq, o.fr_bits) == (6, r.bit_length())
    one = ints(o.to_mont(cref.FQ, limbs([1], 6)), 6)[0]
    assert one == (1 << 384) % c["q"]


@pytest.mark.parametrize("curve", BLS)
def test_field_ops_vs_python(curve):
    o, c = cref.oracle(curve), CURVES[curve]
    for field, p, nw in ((cref.FR, c["r"], 4), (cref.FQ, c["q"], 6)):
        a, b = o.rand_fe(field, 200, 1), o.rand_fe(field, 200, 2)
        a[0], b[0] = 0, 0
        a[1] = limbs([p - 1], nw)[0]
        b[1] = a[1]
        a[2] = limbs([1], nw)[0]
        ai, bi = ints(a, nw), ints(b, nw)
        assert all(x < p for x in ai + bi)
        am, bm = o.to_mont(field, a), o.to_mont(field, b)
        assert ints(o.from_mont(field, am), nw) == ai
        assert ints(o.from_mont(field, o.binop(field, 0, am, bm)), nw) == [x * y % p for x, y in zip(ai, bi)]
        assert ints(o.from_mont(field, o.binop(field, 1, am, bm)), nw) == [(x + y) % p for x, y in zip(ai, bi)]
        assert ints(o.from_mont(field, o.binop(field, 2, am, bm)), nw) == [(x - y) % p for x, y in zip(ai, bi)]
        assert ints(o.from_mont(field, o.binop(field, 3, am)), nw) == [x * x % p for x in ai]
        assert ints(o.from_mont(field, o.binop(field, 4, am[1:])), nw) == [pow(x, -1, p) for x in ai[1:]]


@pytest.mark.parametrize("curve", BLS)
@pytest.mark.parametrize("log_n", [0, 1, 3, 6])
def test_ntt_vs_definition(curve, log_n):
    """A_k = sum_j a_j w^(jk) with w = TWO_ADIC_ROOT^(2^(s - log_n)); coset: a_j scaled by g^j first (g = GENERATOR); inverses undo."""
    o, c = cref.oracle(curve), CURVES[curve]
    r, n = c["r"], 1 << log_n
    w = pow(c["root"], 1 << (c["s"] - log_n), r)
    g = c["fr_gen"]
    x = o.rand_fe(cref.FR, n, 10 + log_n)
    xi = ints(x, 4)
    xm = o.to_mont(cref.FR, x)
    fwd = [sum(xi[j] * pow(w, j * k, r) for j in range(n)) % r for k in range(n)]
    cfwd = [sum(xi[j] * pow(g, j, r) * pow(w, j * k, r) for j in range(n)) % r for k in range(n)]
    assert ints(o.from_mont(cref.FR, o.ntt(xm, log_n)), 4) == fwd
    assert ints(o.from_mont(cref.FR, o.ntt(xm, log_n, coset=True)), 4) == cfwd
    assert np.array_equal(o.ntt(o.ntt(xm, log_n), log_n, inverse=True), xm)
    assert np.array_equal(o.ntt(o.ntt(xm, log_n, coset=True), log_n, inverse=True, coset=True), xm)


@pytest.mark.parametrize("curve", BLS)
def test_g1_and_msm_vs_affine_definition(curve):
    o, c = cref.oracle(curve), CURVES[curve]
    q, r = c["q"], c["r"]
    G = c["gen"]
    mq = lambda pts: o.to_mont(cref.FQ, limbs([v for P in pts for v in (P if P else (0, 0))], 6)).reshape(len(pts), 12)
    gm = mq([G])[0].copy()
    assert o.g1_on_curve(gm)
    rnd = random.Random(5)
    ks = [0, 1, 2, r - 1, r - 2] + [rnd.randrange(r) for _ in range(7)]
    got = o.g1_mul(gm, limbs(ks, 4))
    exp = [gen_curves.ec_mul(k, G, q) for k in ks]
    assert np.array_equal(got, mq(exp))
    # MSM over points k_i G with the edge scalars the reference meets (0, 1, r - 1), the point at infinity and a repeated base
    n = 70
    pk = [rnd.randrange(1, r) for _ in range(n)]
    pts = o.g1_mul(gm, limbs(pk, 4))
    pts[3] = 0
    pk[3] = 0
    pts[5] = pts[4]
    pk[5] = pk[4]
    sc = [rnd.randrange(r) for _ in range(n)]
    sc[0], sc[1], sc[2], sc[7] = 0, 1, r - 1, 1
    out, inf = o.msm_g1(pts, limbs(sc, 4))
    exp_k = sum(s * k for s, k in zip(sc, pk)) % r
    assert not inf and np.array_equal(out, mq([gen_curves.ec_mul(exp_k, G, q)])[0])
    out, inf = o.msm_g1(pts[:2], limbs([r - 1, 0], 4))
    assert np.array_equal(out, mq([gen_curves.ec_mul((r - 1) * pk[0] % r, G, q)])[0])
    out, inf = o.msm_g1(pts[:4], limbs([0, 0, 0, 0], 4))
    assert inf and not out.any()


@pytest.mark.parametrize("curve", ["bn254"] + BLS)
def test_every_build_exports_the_whole_abi(curve):
    """No compute calls (no GPU here): each curve's shared object loads, exports every symbol include/zkb200.h declares and
    reports its curve; the key-file readers are there on every curve (a missing file is ZKB_ERR_INVALID, not UNSUPPORTED)."""
    from zkt_plonk_b200 import _lib
    lib = _lib.lib(curve)
    missing = [s for s in _lib.declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    cid, frw, fqw, bits, has_prover = _lib.curve_info(curve)
    assert (cid, frw) == (_lib.CURVE_IDS[curve], 4)
    assert (fqw, bits, has_prover) == {"bn254": (4, 254, 1), "bls12_381": (6, 255, 1), "bls12_377": (6, 253, 1)}[curve]
    assert lib.zkb_plonk_proof_bytes() == (802 if curve == "bn254" else 1010)
    gen = np.zeros(2 * fqw, dtype=np.uint64)
    assert lib.zkb_g1_generator(ctypes.c_void_p(gen.ctypes.data)) == 0
    o = cref.oracle(curve)
    assert o.g1_on_curve(gen)
    exp = CURVES[curve]["gen"]
    assert ints(o.from_mont(cref.FQ, gen.reshape(2, fqw)), fqw) == list(exp)
    n = ctypes.c_size_t(0)
    assert lib.zkb_ck_file_info(b"/nonexistent", ctypes.byref(n), ctypes.byref(n)) == _lib.ZKB_ERR_INVALID


# ------------------------------------------------------------------------------------------------ the whole protocol on BLS12
@pytest.fixture
def on_curve(request):
    """run a test with the Python mirror (field / pyref) switched to a curve, BN254 again afterwards"""
    from oracle import pyref
    from zkt_plonk_b200 import field
    field.use_curve(request.param)
    pyref.use_curve(request.param)
    yield request.param
    field.use_curve("bn254")
    pyref.use_curve("bn254")


@pytest.mark.parametrize("on_curve", BLS, indirect=True)
def test_prove_verify_roundtrip_on_oracle_backend(on_curve):
    """The reference's own acceptance test on these curves (plonk.rs:226-254 test_full on Bls12_381 / Bls12_377): the round
    schedule (zkt_plonk_b200.prover) over the per-curve oracle gives a proof of 11 * 48 + 2 * 49 + 12 * 32 = 1010 bytes that the
    restated verifier accepts (PC::check through the synthetic SRS's trapdoor and as the reference's product of pairings on the
    curve's own pairing, oracle/pairing_bls.py), that parses back to itself, and tampered proofs are rejected."""
    from oracle import plonk_ref
    from zkt_plonk_b200 import field, prover, synthetic
    P = field.R_MOD
    tau = 0x1D9E5F1B2C3A49587766554433221100FFEEDDCCBBAA99887766554433221101 % P
    circ = synthetic.make_circuit(5, seed=5, table_size=4)
    assert synthetic.check_gates(circ)
    srs = plonk_ref.make_srs_host(circ.n + 8, tau)
    assert srs.shape == (circ.n + 8, 12)
    be = plonk_ref.OracleBackend(srs)
    pk, vk = prover.setup(be, circ)
    rnd = random.Random(7)
    blinders = [rnd.randrange(P) for _ in range(19)]
    proof = prover.prove(be, pk, vk, circ, blinders)
    raw = proof.to_bytes()
    assert len(raw) == 11 * 48 + 2 * 49 + 12 * 32
    pub = list(circ.pi.values())
    assert plonk_ref.verify(vk, proof, pub, tau) == 0
    assert plonk_ref.verify(vk, proof, pub, tau + 1) == 1
    # the reference's own PC::check: a product of two pairings per opening on this curve's pairing, no trapdoor involved
    cvk = plonk_ref.make_cvk(tau)
    assert plonk_ref.verify(vk, proof, pub, cvk=cvk) == 0
    assert plonk_ref.verify(vk, proof, pub, cvk=plonk_ref.make_cvk(tau + 1)) == 1
    # the library's verifier of this curve's build (csrc/verify.cu: host code, its own pairing) agrees
    from zkt_plonk_b200 import verifier
    lib_cvk = verifier.make_cvk(tau)
    assert verifier.verify(vk, raw, pub, lib_cvk) == 0
    assert verifier.verify(vk, raw, pub, verifier.make_cvk(tau + 1)) == 1
    assert verifier.verify(vk, raw, [(pub[0] + 1) % P] + pub[1:], lib_cvk) != 0
    flipped = bytearray(raw)
    flipped[-1] ^= 1                                               # the last evaluation (h2)
    assert verifier.verify(vk, bytes(flipped), pub, lib_cvk) != 0
    swapped = raw[:12 * 48 + 1] + raw[11 * 48: 12 * 48] + raw[13 * 48 + 1:]       # the second witness replaced by the first
    assert len(swapped) == len(raw) and verifier.verify(vk, swapped, pub, lib_cvk) == 2
    back = prover.proof_from_bytes(raw)
    assert back.to_bytes() == raw and back.commits == proof.commits and back.aw == proof.aw and back.evals == proof.evals
    assert prover.prove(be, pk, vk, circ, blinders).to_bytes() == raw
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.evals["a"] = (bad.evals["a"] + 1) % P
    assert plonk_ref.verify(vk, bad, pub, tau) != 0
    assert plonk_ref.verify(vk, proof, [(pub[0] + 1) % P] + pub[1:], tau) != 0
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.commits["z1"] = proof.commits["z2"]
    assert plonk_ref.verify(vk, bad, pub, tau) != 0
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.evals["h1_next"] = (bad.evals["h1_next"] + 1) % P                  # an evaluation of the second opening
    assert plonk_ref.verify(vk, bad, pub, tau) != 0
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.aw, dict(proof.evals))   # the second witness replaced
    assert plonk_ref.verify(vk, bad, pub, tau) == 2


@pytest.mark.parametrize("on_curve", BLS, indirect=True)
def test_key_files_on_the_bls12_curves(on_curve, tmp_path):
    """The reference's key types are generic over the pairing engine (keys/mod.rs:29-40,180-203 derive CanonicalSerialize), so
    on Bls12_381 / Bls12_377 the files are the same layouts with 48-byte base-field elements: this curve's build of
    csrc/keyfile.cu against the independent Python serialiser (oracle/arkser.py) in both directions, byte for byte; the G2 head
    of a cvk file feeds this curve's pairing verifier; malformed points are refused."""
    from oracle import arkser, plonk_ref, pyref
    from zkt_plonk_b200 import _lib, field, keyfile, prover, synthetic, verifier
    arkser.use_curve(on_curve)
    try:
        P, Q, w = field.R_MOD, field.Q_MOD, field.FQ_WORDS
        assert (arkser.FQ_BYTES, w) == (48, 6)
        tau = 0x1D9E5F1B2C3A49587766554433221100FFEEDDCCBBAA99887766554433221101 % P
        circ = synthetic.make_circuit(4, seed=2, table_size=4)
        be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, tau))
        pk, vk = prover.setup(be, circ)
        polys = {name: prover.mont_array_to_ints(pk.polys[name].data[: pk.polys[name].len]) for name in arkser.PK_ORDER}

        def to_array(pts):
            out = np.zeros((len(pts), 2 * w), dtype=np.uint64)
            for i, pt in enumerate(pts):
                if pt is not None:
                    for j, v in enumerate(pt):
                        out[i, w * j: w * j + w] = field.int_to_limbs((v << (64 * w)) % Q, w)
            return out

        def to_pts(arr):
            return [prover.point_to_ints(row, not row.any()) for row in arr]

        # pk: scalar-field coefficients only -- the same bytes on every curve with a 32-byte Fr
        ref_pk = arkser.prover_key(polys)
        (tmp_path / "pk").write_bytes(ref_pk)
        got = keyfile.pk_read(tmp_path / "pk")
        assert all(prover.mont_array_to_ints(got[name]) == polys[name] for name in arkser.PK_ORDER)
        keyfile.pk_write(tmp_path / "pk2", got)
        assert (tmp_path / "pk2").read_bytes() == ref_pk
        # vk: ten 96-byte commitments
        ref_vk = arkser.verifier_key(vk.n, vk.pi_roots, vk.commits)
        assert len(ref_vk) == 8 + 8 + 32 * len(vk.pi_roots) + 10 * 96
        (tmp_path / "vk").write_bytes(ref_vk)
        n, roots, xy, inf = keyfile.vk_read(tmp_path / "vk")
        assert n == vk.n and prover.mont_array_to_ints(roots) == vk.pi_roots and xy.shape == (10, 12)
        assert dict(zip(arkser.PK_ORDER, to_pts(xy))) == vk.commits and not any(inf)
        keyfile.vk_write(tmp_path / "vk2", n, roots, xy, inf)
        assert (tmp_path / "vk2").read_bytes() == ref_vk
        commits = dict(vk.commits)
        commits["q_c"] = None                                     # identity: (0, 1) + flag in the last of the 96 bytes
        (tmp_path / "vk").write_bytes(arkser.verifier_key(vk.n, [], commits))
        n, roots, xy, inf = keyfile.vk_read(tmp_path / "vk")
        assert inf[4] and not xy[4].any() and sum(inf) == 1
        keyfile.vk_write(tmp_path / "vk2", n, roots, xy)
        assert (tmp_path / "vk2").read_bytes() == (tmp_path / "vk").read_bytes()
        bad = bytearray(ref_vk)
        bad[-96: -48] = Q.to_bytes(48, "little")                  # x == q: not a field element
        (tmp_path / "vk").write_bytes(bytes(bad))
        with pytest.raises(_lib.ZkbError):
            keyfile.vk_read(tmp_path / "vk")
        (tmp_path / "vk").write_bytes(ref_vk[:-48])               # a BN254-sized tail
        with pytest.raises(_lib.ZkbError):
            keyfile.vk_read(tmp_path / "vk")
        # ck
        rnd = random.Random(4)
        pts = [pyref.g1_mul(rnd.randrange(1, P), pyref.G1_GEN) for _ in range(11)] + [None]
        ref_ck = arkser.committer_key(pts, pts[:2], 1 << 20)
        (tmp_path / "ck").write_bytes(ref_ck)
        assert keyfile.ck_info(tmp_path / "ck") == (12, 1 << 20)
        assert to_pts(keyfile.ck_read(tmp_path / "ck")) == pts and to_pts(keyfile.ck_read(tmp_path / "ck", 3, 5)) == pts[3:8]
        keyfile.ck_write(tmp_path / "ck2", to_array(pts), to_array(pts[:2]), 1 << 20)
        assert (tmp_path / "ck2").read_bytes() == ref_ck
        assert arkser.parse_committer_key(ref_ck) == (pts, pts[:2], 1 << 20)
        # cvk: h, beta_h from the file verify a proof made on the same SRS (this curve's pairing, csrc/verify.cu)
        h, beta_h = plonk_ref.make_cvk(tau)
        gamma_g = pyref.g1_mul(77, pyref.G1_GEN)
        (tmp_path / "cvk").write_bytes(arkser.sonic_verifier_key(pyref.G1_GEN, gamma_g, h, beta_h, 4 * circ.n, 1 << 20))
        g_arr, gg_arr, h_arr, bh_arr = keyfile.cvk_read(tmp_path / "cvk")
        assert to_pts(np.stack([g_arr, gg_arr])) == [pyref.G1_GEN, gamma_g]
        assert np.array_equal(h_arr, verifier.g2_array(h)) and np.array_equal(bh_arr, verifier.g2_array(beta_h))
        raw = prover.prove(be, pk, vk, circ, list(range(3, 22))).to_bytes()
        pub = list(circ.pi.values())
        assert verifier.verify(vk, raw, pub, (h_arr, bh_arr)) == 0
        assert verifier.verify(vk, raw, pub, (h_arr, h_arr)) != 0
        # mutated files (byte flips, truncations, inflated length fields, garbage tails): an error or a parse, never a crash
        files = {"pk": (ref_pk, keyfile.pk_read), "vk": (ref_vk, keyfile.vk_read), "ck": (ref_ck, keyfile.ck_read),
                 "cvk": ((tmp_path / "cvk").read_bytes(), keyfile.cvk_read)}
        refused = 0
        for name, (good, reader) in files.items():
            for trial in range(60):
                data = bytearray(good)
                kind = trial % 4
                if kind == 0:
                    for _ in range(rnd.randrange(1, 4)):
                        data[rnd.randrange(len(data))] ^= 1 << rnd.randrange(8)
                elif kind == 1:
                    data = data[: rnd.randrange(len(data))]
                elif kind == 2:
                    off = rnd.choice([0, 8, 11, 19])
                    data[off: off + 8] = rnd.choice([1 << 40, (1 << 64) - 1, 1 << 31]).to_bytes(8, "little")
                else:
                    data += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 40)))
                (tmp_path / "mut").write_bytes(bytes(data))
                try:
                    reader(tmp_path / "mut")
                except _lib.ZkbError:
                    refused += 1
        assert refused > 50
    finally:
        arkser.use_curve("bn254")
