"""CPU-only: the host-side verifier of the C ABI (csrc/verify.cu: zkb_plonk_verify, zkb_pairing) against the Python
restatements (oracle/pairing.py, oracle/plonk_ref.py) -- the pairing coefficient by coefficient, the verifier on accepted
proofs (both transcripts), on every kind of tampering the reference's checks catch, and on malformed inputs."""
import random

import numpy as np
import pytest

from oracle import pairing as pr
from oracle import plonk_ref, pyref
from zkt_plonk_b200 import _lib, prover, synthetic, verifier

P = prover.P
TAU = 0x1D9E5F1B2C3A49587766554433221100FFEEDDCCBBAA99887766554433221101 % P


def test_pairing_matches_the_python_restatement():
    rnd = random.Random(8)
    for _ in range(2):
        a, b = rnd.randrange(1, P), rnd.randrange(1, P)
        p1, q2 = pyref.g1_mul(a, pyref.G1_GEN), pr.g2_mul(b, pr.G2_GEN)
        assert verifier.pairing(p1, q2) == pr.pairing(q2, p1)
    assert verifier.pairing(None, pr.G2_GEN) == pr.F12_ONE == verifier.pairing(pyref.G1_GEN, None)
    with pytest.raises(_lib.ZkbError):
        verifier.pairing((1, 3), pr.G2_GEN)                              # not on y^2 = x^3 + 3
    with pytest.raises(_lib.ZkbError):
        verifier.pairing(pyref.G1_GEN, ((1, 2), (3, 4)))                 # not on the twist


def test_pairing_product_check():
    a = 0x123456789ABCDEF
    g1, g2 = pyref.G1_GEN, pr.G2_GEN
    good = [(pyref.g1_mul(a, g1), g2), (pyref.g1_neg(g1), pr.g2_mul(a, g2))]
    assert verifier.pairing_product_is_one(good)
    assert not verifier.pairing_product_is_one([(pyref.g1_mul(a + 1, g1), g2), good[1]])
    assert verifier.pairing_product_is_one([(None, g2)])


@pytest.fixture(scope="module")
def proved():
    circ = synthetic.make_circuit(5, seed=17, table_size=8)
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    pk, vk = prover.setup(be, circ)
    blinders = list(range(31, 50))
    return circ, vk, {name: prover.prove(be, pk, vk, circ, blinders, transcript=name).to_bytes() for name in ("merlin", "ethereum")}


def test_verifier_accepts_and_rejects_like_the_restated_one(proved):
    circ, vk, proofs = proved
    pub = list(circ.pi.values())
    cvk = plonk_ref.make_cvk(TAU)
    for name, raw in proofs.items():
        assert verifier.verify(vk, raw, pub, cvk, transcript=name) == 0
        assert plonk_ref.verify(vk, prover.proof_from_bytes(raw), pub, cvk=cvk, transcript=name) == 0
    raw = proofs["merlin"]
    assert verifier.verify(vk, raw, pub, cvk, transcript="ethereum") != 0            # wrong transcript
    assert verifier.verify(vk, raw, pub, plonk_ref.make_cvk(TAU + 1)) == 1            # wrong SRS
    assert verifier.verify(vk, raw, [(pub[0] + 1) % P] + pub[1:], cvk) != 0           # wrong public input
    # tampering: same decisions (including WHICH opening fails) as the Python verifier
    proof = prover.proof_from_bytes(raw)
    cases = []
    for key in ("a", "h1_next", "t_next", "z2_next", "q_lookup"):
        bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
        bad.evals[key] = (bad.evals[key] + 1) % P
        cases.append(bad)
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.commits["z1"] = proof.commits["z2"]
    cases.append(bad)
    bad = prover.Proof(dict(proof.commits), proof.saw, proof.aw, dict(proof.evals))      # openings swapped
    cases.append(bad)
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.commits["q_hi"] = None                                                            # identity commitment
    cases.append(bad)
    for bad in cases:
        got = verifier.verify(vk, bad.to_bytes(), pub, cvk)
        assert got != 0 and got == plonk_ref.verify(vk, bad, pub, cvk=cvk)


def test_verifier_refuses_malformed_input(proved):
    circ, vk, proofs = proved
    pub = list(circ.pi.values())
    cvk = plonk_ref.make_cvk(TAU)
    raw = bytearray(proofs["merlin"])
    x_bad = next(x for x in range(1, 50) if pow((x ** 3 + 3) % pr.Q, (pr.Q - 1) // 2, pr.Q) != 1)   # x^3 + 3 is a non-residue
    bad = bytearray(raw)
    bad[0:32] = x_bad.to_bytes(32, "little")
    with pytest.raises(_lib.ZkbError):
        verifier.verify(vk, bytes(bad), pub, cvk)
    bad = bytearray(raw)
    bad[384] = 1                                                                      # Option tag of kzg10::Proof::random_v
    with pytest.raises(_lib.ZkbError):
        verifier.verify(vk, bytes(bad), pub, cvk)
    bad = bytearray(raw)
    bad[418:450] = b"\xff" * 32                                                       # evaluation >= r
    with pytest.raises(_lib.ZkbError):
        verifier.verify(vk, bytes(bad), pub, cvk)
    with pytest.raises(_lib.ZkbError):
        verifier.verify(vk, bytes(raw), pub, (((1, 2), (3, 4)), cvk[1]))              # h not on the twist
    # ark-ec 0.3 SWFlags::from_u8: infinity + positive-y both set is not a flag; the x field must be canonical even under
    # the infinity flag -- the accepted byte set must equal arkworks' (no malleated encodings of the same point)
    bad = bytearray(raw)
    bad[31] |= 0xC0
    with pytest.raises(_lib.ZkbError):
        verifier.verify(vk, bytes(bad), pub, cvk)
    bad = bytearray(raw)
    bad[0:32] = b"\xff" * 31 + b"\x7f"                                                # infinity flag, x field >= q
    with pytest.raises(_lib.ZkbError):
        verifier.verify(vk, bytes(bad), pub, cvk)


def test_g2_mul_against_the_oracle():
    """zkb_g2_mul (setup-time beta_h = beta * h of a synthetic SRS; bench.py verifies its timed proofs with it) against the
    Python restatement of G2 arithmetic, incl. k = 0, 1, r - 1, r."""
    from oracle import pairing
    for k in (0, 1, 2, 12345678901234567890123, P - 1, P, TAU):
        got = verifier.g2_mul(k)
        exp = pairing.g2_mul(k % P, pairing.G2_GEN) if k % P else None
        assert np.array_equal(got, verifier.g2_array(exp)), k
    h, bh = verifier.make_cvk(TAU)
    eh, ebh = plonk_ref.make_cvk(TAU)
    assert np.array_equal(h, verifier.g2_array(eh)) and np.array_equal(bh, verifier.g2_array(ebh))


def test_cpp_mirror_reads_key_files_and_verifies(tmp_path, proved):
    """include/zkb200.hpp's host-only half (zkb::read_vk_file, zkb::read_cvk_file, zkb::verify) in a C++ program
    (tests/cpp/test_mirror_host.cpp): files written by the Python restatement, a proof from the oracle backend."""
    import os
    import subprocess
    from oracle import arkser
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp"), "-s", "test_mirror_host"])
    circ, vk, proofs = proved
    h, beta_h = plonk_ref.make_cvk(TAU)
    (tmp_path / "vk").write_bytes(arkser.verifier_key(vk.n, vk.pi_roots, vk.commits))
    (tmp_path / "cvk").write_bytes(arkser.sonic_verifier_key(pyref.G1_GEN, pyref.G1_GEN, h, beta_h, 4 * circ.n, 4 * circ.n))
    (tmp_path / "proof").write_bytes(proofs["merlin"])
    (tmp_path / "pub").write_bytes(prover.ints_to_mont_array(list(circ.pi.values())).tobytes())
    out = subprocess.run([os.path.join(root, "tests", "cpp", "test_mirror_host")] + [str(tmp_path / f) for f in ("vk", "cvk", "proof", "pub")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"n={vk.n} roots={len(vk.pi_roots)} verify=0" in out.stdout
