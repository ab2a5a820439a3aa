"""GPU: the full five-round prover on the CUDA backend gives byte-identical proofs to the oracle backend on the same
circuit, witness, SRS and blinders, and the restated verifier accepts them (north-star acceptance bar)."""
import random

import numpy as np
import pytest

from oracle import cref, plonk_ref
from tests.util import gen_xy, to_dev, to_host
from zkt_plonk_b200 import prover, synthetic

pytestmark = pytest.mark.gpu
P = prover.P
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P


def gpu_srs(ctx, n_points):
    """[tau^i] G built in HBM; returns (device tensor, host array)."""
    import torch
    powers, x = [], 1
    for _ in range(n_points):
        powers.append(x)
        x = x * TAU % P
    k = cref.ints_to_limbs(powers)
    out = torch.empty((n_points, 8), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(gen_xy(), to_dev(k), n_points, out)
    torch.cuda.synchronize()
    return out, to_host(out)


@pytest.mark.parametrize("log_n,fixed_base", [(5, False), (8, False), (10, True)])
def test_gpu_proof_is_byte_identical_and_verifies(ctx, log_n, fixed_base):
    import zkt_plonk_b200 as z
    circ = synthetic.make_circuit(log_n, seed=40 + log_n, table_size=min(64, (1 << log_n) // 4))
    assert synthetic.check_gates(circ)
    d_srs, h_srs = gpu_srs(ctx, circ.n + 8)
    kzg = z.GpuKZG10(ctx)
    kzg.load_committer_key(d_srs)
    if fixed_base:
        ctx.srs_precompute(0)
    gbe, obe = prover.GpuBackend(kzg), plonk_ref.OracleBackend(h_srs)
    rnd = random.Random(99)
    blinders = [rnd.randrange(P) for _ in range(19)]
    gpk, gvk = prover.setup(gbe, circ)
    opk, ovk = prover.setup(obe, circ)
    assert gvk.commits == ovk.commits and gvk.pi_roots == ovk.pi_roots
    gproof = prover.prove(gbe, gpk, gvk, circ, blinders)
    oproof = prover.prove(obe, opk, ovk, circ, blinders)
    assert gproof.to_bytes() == oproof.to_bytes()
    assert plonk_ref.verify(gvk, gproof, list(circ.pi.values()), TAU) == 0
    assert plonk_ref.verify(gvk, gproof, list(circ.pi.values()), cvk=plonk_ref.make_cvk(TAU)) == 0     # PC::check by pairings
    ctx.srs_precompute(-1)


def test_degenerate_circuits_in_the_native_driver(ctx):
    """Same shapes as tests/test_prover_cpu.py::test_degenerate_circuits_behave_like_the_reference.  "no_lookup": h1 and h2
    are zero polynomials before blinding (add_blinders_to_poly with len = 0 < k) and z2 has one coefficient; the native
    driver must emit the bytes the oracle-backend schedule emits.  "const_wire": the quotient is not a polynomial of
    3n + 6 coefficients; zkb_plonk_prove must return an error (the reference fails in PC::commit), not overrun a buffer."""
    circ = synthetic.make_edge_circuit(5, "no_lookup", seed=3)
    d_srs, h_srs = gpu_srs(ctx, circ.n + 8)
    ctx.srs_load(d_srs)
    blinders = list(range(5, 24))
    native = prover.NativeProver(ctx, circ)
    raw = native.prove_bytes(blinders)
    native.close()
    obe = plonk_ref.OracleBackend(h_srs)
    opk, ovk = prover.setup(obe, circ)
    assert raw == prover.prove(obe, opk, ovk, circ, blinders).to_bytes()
    circ = synthetic.make_edge_circuit(5, "const_wire", seed=3)
    native = prover.NativeProver(ctx, circ)
    with pytest.raises(Exception, match="quotient longer"):
        native.prove_bytes(blinders)
    native.close()


@pytest.mark.parametrize("log_n", [6, 12])
def test_wires_gathered_on_the_device_give_the_same_proof(ctx, log_n):
    """zkb_plonk_pk_set_wiring + zkb_plonk_prove_vars (ProvingComposer::wire_evals on the device, prove.rs:49-55; SURVEY.md
    8f-1): the key keeps w_l / w_r / w_o, the proof uploads the variable assignment only.  Same 802 bytes as the call that
    takes the three wire vectors; bad inputs are refused."""
    import ctypes
    from zkt_plonk_b200._lib import ZkbError
    circ = synthetic.make_circuit(log_n, seed=80 + log_n, table_size=min(64, (1 << log_n) // 4))
    assert circ.var_values.shape[0] < 3 * circ.n
    for k, w in enumerate((circ.a, circ.b, circ.c)):
        assert np.array_equal(circ.var_values[circ.wiring[k]], w)
    d_srs, _ = gpu_srs(ctx, circ.n + 8)
    ctx.srs_load(d_srs)
    native = prover.NativeProver(ctx, circ)
    blinders = list(range(300, 319))
    want = native.prove_bytes(blinders)
    with pytest.raises(ZkbError):
        native.prove_bytes(blinders, from_vars=True)              # no wiring yet
    native.set_wiring()
    assert native.prove_bytes(blinders, from_vars=True) == want
    assert native.prove_bytes(blinders, from_vars=True) == want   # and again (the assignment buffer is reused)
    assert native.prove_bytes(blinders) == want                   # the wire-vector entry still works on the same key
    good = circ.var_values
    try:
        circ.var_values = good[: good.shape[0] - 1]               # the wiring refers to a variable beyond n_vars
        with pytest.raises(ZkbError):
            native.prove_bytes(blinders, from_vars=True)
        bad = good.copy()
        bad[0, 0] = 1                                              # Variable::Zero must be zero
        circ.var_values = bad
        with pytest.raises(ZkbError):
            native.prove_bytes(blinders, from_vars=True)
    finally:
        circ.var_values = good
    assert native.prove_bytes(blinders, from_vars=True) == want
    native.close()


def test_gpu_proof_2_14_verifies(ctx):
    """Larger circuit: too slow for the Python-int oracle prover, so acceptance by the verifier is the check."""
    import zkt_plonk_b200 as z
    log_n = 14
    circ = synthetic.make_circuit(log_n, seed=5)
    d_srs, _ = gpu_srs(ctx, circ.n + 8)
    kzg = z.GpuKZG10(ctx)
    kzg.load_committer_key(d_srs)
    ctx.srs_precompute(0)
    be = prover.GpuBackend(kzg)
    pk, vk = prover.setup(be, circ)
    proof = prover.prove(be, pk, vk, circ, list(range(100, 119)))
    assert plonk_ref.verify(vk, proof, list(circ.pi.values()), TAU) == 0
    bad = prover.Proof(dict(proof.commits), proof.aw, proof.saw, dict(proof.evals))
    bad.evals["t_next"] = (bad.evals["t_next"] + 1) % P
    assert plonk_ref.verify(vk, bad, list(circ.pi.values()), TAU) != 0
    ctx.srs_precompute(-1)


@pytest.mark.parametrize("log_n,fixed_base", [(6, False), (10, True)])
def test_native_cpp_driver_matches_python_and_oracle(ctx, log_n, fixed_base):
    """zkb_plonk_setup / zkb_plonk_prove (the C++ round driver a Rust FFI crate would call) produce the same 802
    bytes as the Python schedule on the GPU backend and on the oracle backend, and the verifier accepts them."""
    import zkt_plonk_b200 as z
    circ = synthetic.make_circuit(log_n, seed=70 + log_n, table_size=min(64, (1 << log_n) // 4))
    d_srs, h_srs = gpu_srs(ctx, circ.n + 8)
    kzg = z.GpuKZG10(ctx)
    kzg.load_committer_key(d_srs)
    if fixed_base:
        ctx.srs_precompute(0)
    rnd = random.Random(5)
    blinders = [rnd.randrange(P) for _ in range(19)]
    native = prover.NativeProver(ctx, circ)
    raw, tm = native.prove_bytes(blinders, timings=True)
    assert raw == native.prove_bytes(blinders)
    obe = plonk_ref.OracleBackend(h_srs)
    opk, ovk = prover.setup(obe, circ)
    assert native.vk().commits == ovk.commits
    assert raw == prover.prove(obe, opk, ovk, circ, blinders).to_bytes()
    gbe = prover.GpuBackend(kzg)
    gpk, gvk = prover.setup(gbe, circ)
    assert raw == prover.prove(gbe, gpk, gvk, circ, blinders).to_bytes()
    proof = prover.proof_from_bytes(raw)
    assert proof.to_bytes() == raw
    assert plonk_ref.verify(native.vk(), proof, list(circ.pi.values()), TAU) == 0
    from zkt_plonk_b200 import verifier                                   # zkb_plonk_verify: PC::check by pairings, host side
    assert verifier.verify(native.vk(), raw, list(circ.pi.values()), plonk_ref.make_cvk(TAU)) == 0
    assert tm["total_ms"] > 0
    native.close()
    ctx.srs_precompute(-1)


def test_native_driver_with_the_ethereum_transcript(ctx):
    """zkb_plonk_pk_set_transcript(1): the C++ driver with `T = EthereumTranscript` (gadgets/src/transcript.rs:8-90) gives
    the bytes of the Python schedule over the oracle backend with the same transcript; switching back restores Merlin."""
    import zkt_plonk_b200 as z
    circ = synthetic.make_circuit(6, seed=21, table_size=16)
    d_srs, h_srs = gpu_srs(ctx, circ.n + 8)
    kzg = z.GpuKZG10(ctx)
    kzg.load_committer_key(d_srs)
    blinders = list(range(500, 519))
    native = prover.NativeProver(ctx, circ)
    merlin_raw = native.prove_bytes(blinders)
    native.set_transcript("ethereum")
    raw = native.prove_bytes(blinders)
    obe = plonk_ref.OracleBackend(h_srs)
    opk, ovk = prover.setup(obe, circ)
    assert raw == prover.prove(obe, opk, ovk, circ, blinders, transcript="ethereum").to_bytes() and raw != merlin_raw
    assert plonk_ref.verify(native.vk(), prover.proof_from_bytes(raw), list(circ.pi.values()), TAU, transcript="ethereum") == 0
    native.set_transcript("merlin")
    assert native.prove_bytes(blinders) == merlin_raw
    native.close()


def test_native_driver_from_the_reference_cli_key_files(ctx, tmp_path):
    """ck / pk / vk files in the reference CLI's format (ark-serialize unchecked).  The files the GPU key
    writes are byte-identical to the independent Python serialisation of the oracle-backend keys; a key loaded back
    from files (committer key included) proves byte-identically to the key set up from the composer's columns."""
    import zkt_plonk_b200 as z
    from oracle import arkser
    circ = synthetic.make_circuit(7, seed=33, table_size=16)
    n_powers = 4 * circ.n + 1                                              # PC::trim(pp, 4n, 0, None): plonk.rs:79-85
    d_srs, h_srs = gpu_srs(ctx, n_powers)
    kzg = z.GpuKZG10(ctx)
    kzg.load_committer_key(d_srs)
    blinders = list(range(900, 919))
    native = prover.NativeProver(ctx, circ)
    raw = native.prove_bytes(blinders)
    pk_path, vk_path, ck_path = tmp_path / "pk", tmp_path / "vk", tmp_path / "ck"
    native.save_keys(pk_path, vk_path)
    obe = plonk_ref.OracleBackend(h_srs)
    opk, ovk = prover.setup(obe, circ)
    polys = {name: prover.mont_array_to_ints(opk.polys[name].data[: opk.polys[name].len]) for name in arkser.PK_ORDER}
    assert pk_path.read_bytes() == arkser.prover_key(polys)
    assert vk_path.read_bytes() == arkser.verifier_key(ovk.n, ovk.pi_roots, ovk.commits)
    pts = [prover.point_to_ints(row, not row.any()) for row in h_srs]
    ck_path.write_bytes(arkser.committer_key(pts, pts[:2], n_powers - 1))
    native.close()
    ctx.srs_load(d_srs[:8].contiguous())                                   # forget the key, then read it from the file
    kzg.load_committer_key_file(ck_path)
    assert ctx.srs_size() == n_powers
    loaded = prover.NativeProver(ctx, circ, key_files=(pk_path, vk_path))
    assert loaded.vk().commits == ovk.commits and loaded.vk().pi_roots == ovk.pi_roots
    assert loaded.prove_bytes(blinders) == raw
    out_pk, out_vk = tmp_path / "pk2", tmp_path / "vk2"
    loaded.save_keys(out_pk, out_vk)
    assert out_pk.read_bytes() == pk_path.read_bytes() and out_vk.read_bytes() == vk_path.read_bytes()
    loaded.close()
    # without the vk file's commitments (zkb_plonk_pk_from_polys with vk_xy = NULL) they are recomputed
    recomputed = prover.NativeProver(ctx, circ, key_polys=z.keyfile.pk_read(pk_path))
    assert recomputed.vk().commits == ovk.commits and recomputed.prove_bytes(blinders) == raw
    recomputed.close()
    bad = tmp_path / "bad"
    bad.write_bytes(pk_path.read_bytes()[:-7])
    with pytest.raises(Exception, match="ProverKey"):
        prover.NativeProver(ctx, circ, key_files=(bad, vk_path))


def test_native_driver_reuses_key_across_witnesses(ctx):
    """One key object, several witnesses / tables in a row: the C++ driver keeps its pinned lookup staging zero outside
    the regions a proof writes (sparse combine_split), so a proof must not see what the previous one left behind --
    including a table that holds a zero entry (the zero bucket then sits in the middle of h1 / h2) and a shorter table."""
    import copy
    import zkt_plonk_b200 as z
    log_n = 8
    circ = synthetic.make_circuit(log_n, seed=77, table_size=64)
    d_srs, _ = gpu_srs(ctx, circ.n + 8)
    kzg = z.GpuKZG10(ctx)
    kzg.load_committer_key(d_srs)
    gbe = prover.GpuBackend(kzg)
    gpk, gvk = prover.setup(gbe, circ)
    native = prover.NativeProver(ctx, circ)
    blinders = list(range(300, 319))
    lookup_rows = [i for i, q in enumerate(prover.mont_array_to_ints(circ.selectors["q_lookup"])) if q]
    assert lookup_rows

    def with_table(table):                                                 # same witness (it must stay satisfiable), other table
        v = copy.copy(circ)
        v.table = list(table)
        return v

    c0 = prover.mont_array_to_ints(circ.c)
    used = {c0[i] for i in lookup_rows}
    only_used = [e for e in circ.table if e in used]                       # shorter table
    # (a zero entry is only a valid Plookup table when it comes last, where the padding zeros follow it)
    variants = [circ, with_table(circ.table + [0]), with_table(circ.table[::-1]), with_table(only_used),
                with_table(only_used[::-1] + [0]), circ]
    for v in variants:
        native.circuit = v
        raw = native.prove_bytes(blinders)
        assert raw == prover.prove(gbe, gpk, gvk, v, blinders).to_bytes()
        assert plonk_ref.verify(gvk, prover.proof_from_bytes(raw), list(circ.pi.values()), TAU) == 0
    native.circuit = with_table(only_used[1:])                             # a looked-up output that is not in the table
    with pytest.raises(Exception, match="ElementNotIndexedInTable"):
        native.prove_bytes(blinders)
    native.circuit = circ
    assert native.prove_bytes(blinders) == prover.prove(gbe, gpk, gvk, circ, blinders).to_bytes()
    native.close()


@pytest.mark.parametrize("log_n,table_len,dup", [(6, 9, False), (10, 300, True), (13, 1024, True)])
def test_lookup_multisets_on_the_device_match_combine_split(ctx, log_n, table_len, dup):
    """csrc/lookup.cu against the reference-shaped Python combine_split (multiset.rs:103-146): t = table || zeros,
    f = q_lookup * c, h1 / h2 alternate over the buckets in order of first appearance in t -- dense lookups (60 % of the rows),
    selectors other than 0 / 1, zero inside the table, duplicated table entries; then the two error reports."""
    import torch
    rnd = random.Random(log_n)
    n = 1 << log_n
    table = [rnd.randrange(1, P) for _ in range(table_len)]
    table[2] = 0                                                 # a zero inside the table, before other entries
    if dup:
        table[7] = table[3]
        table[-1] = table[0]
    t = table + [0] * (n - table_len)
    q = [rnd.choice([0, 1, 1, 1, 5]) if rnd.random() < 0.6 else 0 for _ in range(n)]
    c = []
    for qi in q:
        if qi == 0:
            c.append(rnd.randrange(P))                           # no lookup gate: c is free, f is zero
        else:
            v = rnd.choice(table)
            c.append(v * pow(qi, -1, P) % P)                     # q * c lands in the table
    f = [a * b % P for a, b in zip(q, c)]
    e1, e2 = prover.combine_split(t, f)
    dev = lambda v: to_dev(prover.ints_to_mont_array(v))
    outs = [torch.empty((n, 4), dtype=torch.int64, device="cuda") for _ in range(4)]
    st = ctx.lookup_multisets_dev(log_n, prover.ints_to_mont_array(table), dev(q), dev(c), *outs)
    assert st == 0
    got = [prover.mont_array_to_ints(to_host(o)) for o in outs]
    assert got[0] == t and got[1] == f and got[2] == e1 and got[3] == e2
    # an element the table does not hold (multiset.rs:121) / a zero in f that a full, zero-free table does not hold
    c2 = list(c)
    k = next(i for i, qi in enumerate(q) if qi == 1)
    c2[k] = (max(table) + 12345) % P
    assert ctx.lookup_multisets_dev(log_n, prover.ints_to_mont_array(table), dev(q), dev(c2), *outs) & 1
    full = [rnd.randrange(1, P) for _ in range(n)]
    assert ctx.lookup_multisets_dev(log_n, prover.ints_to_mont_array(full), dev([0] * n), dev(c), *outs) & 1


@pytest.mark.parametrize("log_n,lookup_frac", [(6, 0.01), (10, 0.6), (12, 0.9)])
def test_native_driver_lookup_paths_give_the_same_proof(ctx, log_n, lookup_frac):
    """zkb_plonk_pk_set_lookup_mode: the sparse host path, the device path and the automatic choice (device above n / 8 lookup
    rows) emit the same bytes -- the bytes of the oracle-backend schedule -- also from the variable assignment; a witness whose
    lookup value is not in the table is refused on both paths."""
    import zkt_plonk_b200 as z
    circ = synthetic.make_circuit(log_n, seed=70 + log_n, table_size=min(64, (1 << log_n) // 4), lookup_frac=lookup_frac)
    assert synthetic.check_gates(circ)
    d_srs, h_srs = gpu_srs(ctx, circ.n + 8)
    ctx.srs_load(d_srs)
    rnd = random.Random(5)
    blinders = [rnd.randrange(P) for _ in range(19)]
    obe = plonk_ref.OracleBackend(h_srs)
    opk, ovk = prover.setup(obe, circ)
    want = prover.prove(obe, opk, ovk, circ, blinders).to_bytes()
    native = prover.NativeProver(ctx, circ)
    try:
        for mode in (1, 2, 0):
            native.set_lookup_mode(mode)
            assert native.prove_bytes(blinders) == want, mode
        if circ.wiring is not None:
            native.set_wiring()
            native.set_lookup_mode(2)
            assert native.prove_bytes(blinders, from_vars=True) == want
        # break one lookup row's output wire
        rows = [i for i in range(circ.n) if prover.mont_array_to_ints(circ.selectors["q_lookup"][i:i + 1])[0]]
        bad = synthetic.make_circuit(log_n, seed=70 + log_n, table_size=min(64, (1 << log_n) // 4), lookup_frac=lookup_frac)
        bad.c = bad.c.copy()
        bad.c[rows[0]] = prover.ints_to_mont_array([max(circ.table) + 7])[0]
        nb = prover.NativeProver(ctx, bad)
        try:
            for mode in (1, 2):
                nb.set_lookup_mode(mode)
                with pytest.raises(z.ZkbError) as e:
                    nb.prove_bytes(blinders)
                assert e.value.code == -1 and "ElementNotIndexedInTable" in str(e.value)
        finally:
            nb.close()
    finally:
        native.close()
