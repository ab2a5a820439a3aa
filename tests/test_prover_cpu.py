"""CPU-only: the prover round schedule (zkt_plonk_b200.prover) run over the oracle backend produces proofs that the
restated verifier accepts -- the reference's own acceptance criterion (plonk.rs:191-254 test_full) -- and rejects
tampered ones; also pins the Merlin transcript against merlin's published test vector."""
import random

import numpy as np

import pytest

from oracle import plonk_ref
from zkt_plonk_b200 import prover, synthetic
from zkt_plonk_b200.transcript import Merlin, MerlinTranscript

P = prover.P
TAU = 0x1D9E5F1B2C3A49587766554433221100FFEEDDCCBBAA99887766554433221101 % P


def test_merlin_published_vector():
    """merlin 3.0 `equivalence_simple`: pins STROBE-128 / Keccak-f and the framing the transcript relies on."""
    t = Merlin(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_challenge_scalar_is_31_bytes_le():
    a, b = MerlinTranscript("x"), Merlin(b"x")
    v = a.challenge_scalar("beta")
    assert v == int.from_bytes(b.challenge_bytes(b"beta", 31), "little") and v < 1 << 248


def test_combine_split_reference_golden_vector():
    """multiset.rs:271-329 (test_combine_split): the reference's own expected halves for t = {0..6},
    f = {3,6,0,5,4,3,2,0,0,1,2} -- both host implementations (Python lists, numpy limb rows)."""
    t, f = [0, 1, 2, 3, 4, 5, 6], [3, 6, 0, 5, 4, 3, 2, 0, 0, 1, 2]
    evens, odds = [0, 0, 1, 2, 2, 3, 4, 5, 6], [0, 0, 1, 2, 3, 3, 4, 5, 6]
    assert prover.combine_split(t, f) == (evens, odds)
    a1, a2 = prover.combine_split_arrays(prover.ints_to_mont_array(t), prover.ints_to_mont_array(f))
    assert prover.mont_array_to_ints(a1) == evens and prover.mont_array_to_ints(a2) == odds
    with pytest.raises(ValueError):
        prover.combine_split([1, 2], [3])                      # ElementNotIndexedInTable (multiset.rs:121)


def test_vectorised_host_plumbing_matches_the_reference_shaped_one():
    rnd = random.Random(1)
    n = 256
    table = list(dict.fromkeys(rnd.randrange(P) for _ in range(40)))
    t = prover.table_multiset(table, 64, n)
    q = [rnd.choice([0, 1, 1, 5]) for _ in range(n)]
    c = [rnd.choice(table + [0]) for _ in range(n)]
    c = [ci if qi in (0, 1) else 0 for qi, ci in zip(q, c)]     # keep q*c inside the table
    f = [a * b % P for a, b in zip(q, c)]
    f_arr = prover.lookup_f_array(prover.ints_to_mont_array(q), prover.ints_to_mont_array(c))
    assert prover.mont_array_to_ints(f_arr) == f
    t_arr = prover.table_multiset_array(table, 64, n)
    assert prover.mont_array_to_ints(t_arr) == t
    h1, h2 = prover.combine_split(t, f)
    a1, a2 = prover.combine_split_arrays(t_arr, f_arr)
    assert prover.mont_array_to_ints(a1) == h1 and prover.mont_array_to_ints(a2) == h2
    with pytest.raises(ValueError):
        prover.combine_split_arrays(t_arr, prover.ints_to_mont_array([12345]))
    # zero handling: zero inside the table (before other entries), no zero in t at all, dense non-zero f
    for t2, f2 in (([5, 0, 7, 0, 9, 9], [0, 0, 7, 5, 5, 9]), ([3, 4, 5, 6], [4, 4, 6, 3]), ([1, 2, 0, 0], [2, 2, 2, 1])):
        e1, e2 = prover.combine_split(t2, f2)
        g1, g2 = prover.combine_split_arrays(prover.ints_to_mont_array(t2), prover.ints_to_mont_array(f2))
        assert prover.mont_array_to_ints(g1) == e1 and prover.mont_array_to_ints(g2) == e2
    with pytest.raises(ValueError):
        prover.combine_split_arrays(prover.ints_to_mont_array([3, 4]), prover.ints_to_mont_array([0]))


@pytest.mark.parametrize("log_n", [4, 6])
def test_prove_verify_roundtrip_on_oracle_backend(log_n):
    circ = synthetic.make_circuit(log_n, seed=log_n, table_size=4)
    assert synthetic.check_gates(circ)
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    pk, vk = prover.setup(be, circ)
    rnd = random.Random(7)
    blinders = [rnd.randrange(P) for _ in range(19)]
    proof = prover.prove(be, pk, vk, circ, blinders)
    raw = proof.to_bytes()
    assert len(raw) == 11 * 32 + 2 * 33 + 12 * 32
    pub = list(circ.pi.values())
    assert plonk_ref.verify(vk, proof, pub, TAU) == 0
    # the reference's own PC::check: a product of two pairings per opening, no trapdoor involved
    cvk = plonk_ref.make_cvk(TAU)
    assert plonk_ref.verify(vk, proof, pub, cvk=cvk) == 0
    assert plonk_ref.verify(vk, proof, pub, cvk=plonk_ref.make_cvk(TAU + 1)) == 1
    # same inputs, same bytes; different blinders, different proof that still verifies
    assert prover.prove(be, pk, vk, circ, blinders).to_bytes() == raw
    other = prover.prove(be, pk, vk, circ, [rnd.randrange(P) for _ in range(19)])
    assert other.to_bytes() != raw and plonk_ref.verify(vk, other, pub, TAU) == 0
    # tampering: an evaluation, a public input, a commitment
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.evals["a"] = (bad.evals["a"] + 1) % P
    assert plonk_ref.verify(vk, bad, pub, TAU) != 0 and plonk_ref.verify(vk, bad, pub, cvk=cvk) != 0
    assert plonk_ref.verify(vk, proof, [(pub[0] + 1) % P] + pub[1:], TAU) != 0
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.commits["z1"] = proof.commits["z2"]
    assert plonk_ref.verify(vk, bad, pub, TAU) != 0 and plonk_ref.verify(vk, bad, pub, cvk=cvk) != 0
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.evals["h1_next"] = (bad.evals["h1_next"] + 1) % P                  # only the second opening sees this one
    assert plonk_ref.verify(vk, bad, pub, cvk=cvk) != 0


@pytest.mark.parametrize("length", [0, 1, 2, 3, 6])
def test_oracle_add_blinders_follows_the_reference_for_short_polynomials(length):
    """prove.rs:472-483: extend by the k blinders FIRST, then coeffs[i] -= b_i for every i < k -- also when len < k."""
    rnd = random.Random(length)
    coeffs = [rnd.randrange(P) for _ in range(length)]
    bl = [rnd.randrange(P) for _ in range(3)]
    rust = coeffs + bl
    for i, b in enumerate(bl):
        rust[i] = (rust[i] - b) % P
    data = np.zeros((length + 4, 4), dtype=np.uint64)
    if length:
        data[:length] = prover.ints_to_mont_array(coeffs)
    po = prover.Poly(data, length)
    plonk_ref.OracleBackend(np.zeros((1, 8), dtype=np.uint64)).add_blinders(po, bl)
    assert po.len == length + 3 and prover.mont_array_to_ints(po.data[: po.len]) == rust
    if length == 0:
        assert rust == [0, 0, 0]


def test_degenerate_circuits_behave_like_the_reference():
    """A polynomial shorter than n is blinded at its CURRENT length (prove.rs:472-483), which is not + b(X)(X^n - 1), so
    the blinded polynomial no longer agrees with its evaluations on the domain and the division by Z_H is not exact.
    "no_lookup" (empty table, no lookup rows: h1 = h2 = 0 -> [b-b, ..] = 0, but z2 = 1 -> [1-b0, b0-b1, b1-b2, b2]): the
    remainder has low degree, the quotient still fits 3n + 6 coefficients, and the reference emits a proof that its own
    verifier rejects -- the restated prover must emit the same bytes, not fail.  "const_wire" (a = 5 on every row): the
    interpolated quotient fills all 4n coefficients, q_hi exceeds the committer key and the reference fails in PC::commit;
    the restated prover reports it instead of writing past its buffers."""
    circ = synthetic.make_edge_circuit(4, "no_lookup", seed=2)
    assert synthetic.check_gates(circ)
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    pk, vk = prover.setup(be, circ)
    proof = prover.prove(be, pk, vk, circ, list(range(3, 22)))
    assert proof.commits["h1"] is None and proof.commits["h2"] is None and proof.commits["t"] is None   # zero polynomials
    assert plonk_ref.verify(vk, proof, list(circ.pi.values()), TAU) != 0
    circ = synthetic.make_edge_circuit(4, "const_wire", seed=2)
    assert synthetic.check_gates(circ)
    pk, vk = prover.setup(be, circ)
    with pytest.raises(ValueError, match="quotient longer"):
        prover.prove(be, pk, vk, circ, list(range(3, 22)))


def test_prove_verify_with_the_ethereum_transcript():
    """`T = EthereumTranscript` (bin feature "ethereum-transcript"): same schedule, other challenges; a proof made with
    one transcript is rejected under the other."""
    circ = synthetic.make_circuit(4, seed=11, table_size=4)
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    pk, vk = prover.setup(be, circ)
    blinders = list(range(7, 26))
    pub = list(circ.pi.values())
    eth = prover.prove(be, pk, vk, circ, blinders, transcript="ethereum")
    mer = prover.prove(be, pk, vk, circ, blinders)
    assert eth.to_bytes() != mer.to_bytes()
    assert plonk_ref.verify(vk, eth, pub, TAU, transcript="ethereum") == 0
    assert plonk_ref.verify(vk, eth, pub, TAU) != 0 and plonk_ref.verify(vk, mer, pub, TAU, transcript="ethereum") != 0


def test_unsatisfied_witness_is_rejected():
    circ = synthetic.make_circuit(5, seed=3, table_size=4)
    c = prover.mont_array_to_ints(circ.c)
    row = max(i for i in range(circ.n) if c[i])            # corrupt one gate output
    c[row] = (c[row] + 1) % P
    circ.c = prover.ints_to_mont_array(c)
    assert not synthetic.check_gates(circ)
    be = plonk_ref.OracleBackend(plonk_ref.make_srs_host(circ.n + 8, TAU))
    pk, vk = prover.setup(be, circ)
    try:
        proof = prover.prove(be, pk, vk, circ, list(range(1, 20)))
    except (AssertionError, ValueError, ZeroDivisionError):
        return                                              # e.g. the lookup value left the table
    assert plonk_ref.verify(vk, proof, list(circ.pi.values()), TAU) != 0


def test_sparse_combine_split_matches_dense_across_reuse():
    """The C++ round driver's combine_split only touches the table and the lookup rows and keeps its staging columns
    zero elsewhere (csrc/prover.cu).  Against the dense numpy restatement of multiset.rs:103-146, over a sequence of
    calls that reuse the same staging: tables with a zero entry first / in the middle / last, shrinking tables,
    lookups of the zero entry, and an element that is not in the table."""
    import ctypes
    from zkt_plonk_b200 import _lib
    lib = _lib.lib()
    fn = lib.zkb_test_combine_split
    fn.restype = ctypes.c_int
    vp = ctypes.c_void_p
    fn.argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, vp, vp, ctypes.c_size_t, vp, vp, vp, vp]
    n = 256
    rng = np.random.default_rng(3)
    rows = np.sort(rng.choice(n, size=40, replace=False)).astype(np.uint32)
    h1, h2 = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 4), dtype=np.uint64)
    dirty = (ctypes.c_size_t * 4)(0, n, 0, n)
    lens = (ctypes.c_size_t * 2)()
    base = [int(v) for v in rng.integers(1, 1 << 60, size=30)]

    def run(table, picks):
        t = prover.table_multiset_array(table, 64, n)
        f = np.zeros((n, 4), dtype=np.uint64)
        f[rows] = prover.ints_to_mont_array(picks)
        tab = prover.ints_to_mont_array(table) if table else np.zeros((1, 4), dtype=np.uint64)
        rc = fn(tab.ctypes.data, len(table), n, f.ctypes.data, rows.ctypes.data, len(rows), h1.ctypes.data, h2.ctypes.data,
                ctypes.addressof(dirty), ctypes.addressof(lens))
        return rc, t, f

    cases = [base, base[:3] + [0] + base[3:], [0] + base, base + [0], base[:7], base[:7][::-1], [0, base[0]], base]
    for k, table in enumerate(cases):
        pool = table if k % 2 == 0 else [e for e in table if e] or table
        picks = [pool[int(j) % len(pool)] for j in rng.integers(0, 1 << 30, size=len(rows))]
        if 0 in table:
            picks[0] = 0                                           # a lookup of the zero entry itself
        rc, t, f = run(table, picks)
        assert rc == 0
        e1, e2 = prover.combine_split_arrays(t, f)
        assert (lens[0], lens[1]) == (e1.shape[0], e2.shape[0]) == (n, n)
        assert np.array_equal(h1, e1) and np.array_equal(h2, e2), f"case {k}"
    rc, _, _ = run(base[:5], [base[9]] * len(rows))               # ElementNotIndexedInTable
    assert rc != 0
    rc, t, f = run(base, [base[2]] * len(rows))                   # and the staging is still usable afterwards
    e1, e2 = prover.combine_split_arrays(t, f)
    assert rc == 0 and np.array_equal(h1, e1) and np.array_equal(h2, e2)
