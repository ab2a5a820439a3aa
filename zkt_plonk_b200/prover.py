"""prover.py -- the five prover rounds of zkt-plonk (plonk-core/src/proof_system/prove.rs:59-470) driven over the
sm_100a kernels, every polynomial resident in HBM between rounds.

`setup`  mirrors proof_system/setup.rs:42-166 (10 iFFTs + 10 commitments + key extension),
`prove`  mirrors proof_system/prove.rs:59-470 round by round (same transcript schedule, same blinder order,
         same trailing-zero truncation semantics of DensePolynomial), and `Proof.to_bytes` follows the derived
         CanonicalSerialize of proof_system/proof.rs:106-155 (11 compressed G1, 2 x (w + Option tag), 12 Fr).

Only challenges (32 B), commitments (64 B), evaluations and the 19 blinders cross PCIe.  The heavy steps are calls
into a *backend*: `GpuBackend` (this file, libzkb200.so) is the product; tests run the same schedule over an
oracle backend (oracle/plonk_ref.py) and check the resulting proofs byte for byte and against a restated verifier.
Host-side witness plumbing (wire gathers, f = q_lookup * c, combine_split) stays on the host, as SURVEY.md 8f says.
"""
import numpy as np

from . import field
from .transcript import TRANSCRIPTS, MerlinTranscript, fr_bytes

# The moduli and encodings are the selected curve's (field.use_curve / field.curve): BN254 by default; BLS12-381 / BLS12-377
# (what plonk.rs:226-254 instantiates) run through this Python round schedule and through NativeProver (the C++ driver of that
# curve's build of the library).


def __getattr__(name):
    """`prover.P` / `prover.Q`: r and q of the selected curve (BN254's unless field.use_curve chose another)."""
    if name == "P":
        return field.R_MOD
    if name == "Q":
        return field.Q_MOD
    raise AttributeError(name)


# ------------------------------------------------------------------------------------------------ conversions
def fr_to_limbs(x):
    """canonical int -> (4,) uint64 Montgomery limbs."""
    return np.array(field.int_to_limbs(field.to_mont(x)), dtype=np.uint64)


def limbs_to_fr(l):
    return field.limbs_to_int(l) * field.RINV_R % field.R_MOD


def ints_to_mont_array(vals):
    """list/array of canonical ints -> (n, 4) uint64 Montgomery limbs (vectorised over Python ints)."""
    v = np.array([int(x) * field.MONT_R % field.R_MOD for x in vals], dtype=object)
    out = np.empty((len(v), 4), dtype=np.uint64)
    m = (1 << 64) - 1
    for k in range(4):
        out[:, k] = ((v >> (64 * k)) & m).astype(np.uint64)
    return out


def mont_array_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    v = a[:, 0].astype(object)
    for k in range(1, 4):
        v = v + (a[:, k].astype(object) << (64 * k))
    return [int(x) * field.RINV_R % field.R_MOD for x in v]


def point_to_ints(xy, is_inf):
    """(2 * FQ_WORDS,) uint64 Montgomery affine -> (x, y) canonical ints, or None for the identity."""
    if is_inf:
        return None
    w = field.FQ_WORDS
    return (field.limbs_to_int(xy[:w]) * field.RINV_Q % field.Q_MOD, field.limbs_to_int(xy[w:2 * w]) * field.RINV_Q % field.Q_MOD)


def g1_compressed(pt):
    """GroupAffine::serialize (ark-ec 0.3 / ark-serialize 0.3 SWFlags): x little endian (32 bytes on BN254, 48 on the BLS12
    curves) with flags in the top two bits of the last byte: bit 6 = infinity, bit 7 = (y > -y)."""
    nb = field.FQ_BYTES
    if pt is None:
        b = bytearray(nb)
        b[nb - 1] |= 1 << 6
        return bytes(b)
    x, y = pt
    b = bytearray(int(x).to_bytes(nb, "little"))
    if y > (field.Q_MOD - y) % field.Q_MOD:
        b[nb - 1] |= 1 << 7
    return bytes(b)


# ------------------------------------------------------------------------------------------------ data classes
class Poly:
    """Coefficients in a backend buffer with spare capacity; `len` follows DensePolynomial's truncation."""
    __slots__ = ("data", "len")

    def __init__(self, data, length):
        self.data, self.len = data, length


class Circuit:
    """A synthesised circuit in evaluation form (what SetupComposer / ProvingComposer hold after pad_to(n)).

    selectors: dict q_m,q_l,q_r,q_o,q_c,q_lookup -> (n,4) Montgomery; sigma: (sigma1,sigma2,sigma3) evals (n,4);
    table: list of canonical ints (<= table_size entries); witness a,b,c: (n,4) Montgomery; pi: {row: int}."""

    def __init__(self, log_n, selectors, sigma, table, table_size, a, b, c, pi, wiring=None, var_values=None):
        self.log_n, self.n = log_n, 1 << log_n
        self.selectors, self.sigma, self.table, self.table_size = selectors, sigma, table, table_size
        self.a, self.b, self.c, self.pi = a, b, c, dict(sorted(pi.items()))
        # optional: the composer's wire maps w_l / w_r / w_o ((3, n) uint32 variable indices, 0 = Variable::Zero) and the
        # variable assignment ((n_vars, 4) Montgomery, row 0 zero): a = var_values[wiring[0]] etc. (prove.rs:49-55)
        self.wiring, self.var_values = wiring, var_values


class ProverKey:
    def __init__(self, polys, evals, epk, n, log_n):
        self.polys, self.evals, self.epk, self.n, self.log_n = polys, evals, epk, n, log_n


class VerifierKey:
    ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "sigma1", "sigma2", "sigma3", "q_lookup", "q_table")

    def __init__(self, n, pi_roots, commits):
        self.n, self.pi_roots, self.commits = n, pi_roots, commits

    def seed_transcript(self, tr):
        """keys/mod.rs:260-275."""
        tr.append_u64("circuit_size", self.n)
        for name in self.ORDER:
            tr.append_commitment(name + "_commit", self.commits[name])


class Proof:
    COMMITS = ("a", "b", "c", "t", "h1", "h2", "z1", "z2", "q_lo", "q_mid", "q_hi")
    EVALS = ("a", "b", "c", "sigma1", "sigma2", "z1_next", "q_lookup", "t", "t_next", "z2_next", "h1_next", "h2")

    def __init__(self, commits, aw, saw, evals):
        self.commits, self.aw, self.saw, self.evals = commits, aw, saw, evals

    def to_bytes(self):
        out = b"".join(g1_compressed(self.commits[k]) for k in self.COMMITS)
        out += g1_compressed(self.aw) + b"\x00"            # kzg10::Proof { w, random_v: None }
        out += g1_compressed(self.saw) + b"\x00"
        out += b"".join(fr_bytes(self.evals[k]) for k in self.EVALS)
        return out


# ------------------------------------------------------------------------------------------------ host plumbing
def table_multiset(table, table_size, n):
    """LookupTable::into_multiset (lookup/table.rs:52-61): table entries then zeros up to n."""
    assert n > table_size and len(table) <= table_size
    return list(table) + [0] * (n - len(table))


def table_masks(table_size, n):
    """LookupTable::masks (lookup/table.rs:42-48): q_table evaluations."""
    assert n > table_size
    return [0] * table_size + [1] * (n - table_size)


def combine_split(t, f):
    """MultiSet::combine_split (lookup/multiset.rs:103-146) on canonical ints."""
    counters = {}
    for e in t:
        counters[e] = counters.get(e, 0) + 1
    for e in f:
        if e not in counters:
            raise ValueError("ElementNotIndexedInTable")
        counters[e] += 1
    evens, odds, parity = [], [], False
    for e, cnt in counters.items():
        half = cnt // 2
        evens.extend([e] * half)
        odds.extend([e] * half)
        if cnt % 2 == 1:
            if parity:
                odds.append(e)
                parity = False
            else:
                evens.append(e)
                parity = True
    return evens, odds


# ---- the same plumbing on (n, 4) uint64 Montgomery limb arrays (numpy, no Python-int loop over n) -----------------
_ONE_MONT = np.array(field.int_to_limbs(field.to_mont(1)), dtype=np.uint64)


def _row_nonzero(a):
    """(n, 4) limb rows -> bool mask of non-zero field elements (column ops: much faster than any(axis=1))."""
    return (a[:, 0] | a[:, 1] | a[:, 2] | a[:, 3]) != 0


def table_multiset_array(table, table_size, n):
    assert n > table_size and len(table) <= table_size
    out = np.zeros((n, 4), dtype=np.uint64)
    if len(table):
        out[: len(table)] = ints_to_mont_array(table)
    return out


def lookup_f_array(q_lookup, c):
    """f_i = q_lookup_i * c_i (prove.rs:157-161).  Selector values 0 and 1 are handled as masks; anything else
    falls back to exact integer arithmetic for those rows only."""
    q_lookup, c = np.asarray(q_lookup, dtype=np.uint64), np.asarray(c, dtype=np.uint64)
    is_zero = ~_row_nonzero(q_lookup)
    is_one = ((q_lookup[:, 0] == _ONE_MONT[0]) & (q_lookup[:, 1] == _ONE_MONT[1]) & (q_lookup[:, 2] == _ONE_MONT[2])
              & (q_lookup[:, 3] == _ONE_MONT[3]))
    f = np.zeros_like(c)
    ones = np.flatnonzero(is_one)
    f[ones] = c[ones]
    other = np.flatnonzero(~(is_zero | is_one))
    if other.size:
        qi, ci = mont_array_to_ints(q_lookup[other]), mont_array_to_ints(c[other])
        f[other] = ints_to_mont_array([a * b % field.R_MOD for a, b in zip(qi, ci)])
    return np.ascontiguousarray(f)


def combine_split_arrays(t, f):
    """MultiSet::combine_split (lookup/multiset.rs:103-146): buckets in order of first appearance in t, every
    element of f must be in t, halves alternate on odd counts.  Montgomery limb rows are unique per field element,
    so they serve as keys directly.  Zero rows (table padding, non-lookup gates: the bulk of both multisets) are
    counted separately so that only the non-zero rows are sorted."""
    t, f = np.ascontiguousarray(t, dtype=np.uint64), np.ascontiguousarray(f, dtype=np.uint64)
    n_t = t.shape[0]
    nz_t, nz_f = _row_nonzero(t), _row_nonzero(f)
    zeros_t, zeros_f = n_t - int(nz_t.sum()), f.shape[0] - int(nz_f.sum())
    if zeros_f and not zeros_t:
        raise ValueError("ElementNotIndexedInTable")
    pos = np.concatenate([np.flatnonzero(nz_t), n_t + np.flatnonzero(nz_f)])     # original positions in t ++ f
    rows = np.concatenate([t[nz_t], f[nz_f]])
    if rows.shape[0]:
        _, first, counts = np.unique(rows.view(np.dtype((np.void, 32))).reshape(-1), return_index=True, return_counts=True)
        first_pos = pos[first]
        if (first_pos >= n_t).any():
            raise ValueError("ElementNotIndexedInTable")
        keys = rows[first]
    else:
        first_pos, counts, keys = np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros((0, 4), dtype=np.uint64)
    if zeros_t:                                                    # the zero bucket sits where 0 first appears in t
        first_pos = np.concatenate([first_pos, [int(np.argmin(nz_t))]])
        counts = np.concatenate([counts, [zeros_t + zeros_f]])
        keys = np.concatenate([keys, np.zeros((1, 4), dtype=np.uint64)])
    order = np.argsort(first_pos, kind="stable")                   # IndexMap insertion order
    counts, keys = counts[order], keys[order]
    half, odd = counts // 2, (counts % 2).astype(bool)
    rank = np.cumsum(odd) - 1                                       # k-th odd bucket: even k -> evens, odd k -> odds
    to_even = odd & (rank % 2 == 0)
    evens = np.repeat(keys, half + to_even, axis=0)
    odds = np.repeat(keys, half + (odd & ~to_even), axis=0)
    return np.ascontiguousarray(evens), np.ascontiguousarray(odds)


# ------------------------------------------------------------------------------------------------ GPU backend
class GpuBackend:
    """Heavy steps on the device through the C ABI (zkt_plonk_b200.Context / GpuKZG10)."""

    def __init__(self, kzg):
        import torch
        self.torch = torch
        self.kzg, self.ctx = kzg, kzg.ctx
        self.dev = torch.device("cuda", self.ctx.device)

    def sync(self):
        self.torch.cuda.synchronize(self.dev)

    # -- storage
    def from_host(self, a, cap=None):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        cap = a.shape[0] if cap is None else cap
        buf = self.torch.zeros((cap, 4), dtype=self.torch.int64, device=self.dev)
        if a.shape[0]:
            buf[: a.shape[0]] = self.torch.from_numpy(a.view(np.int64)).to(self.dev)
        return buf

    def to_host(self, buf, length):
        return buf[:length].cpu().numpy().view(np.uint64)

    def slice_copy(self, buf, lo, hi, cap):
        out = self.torch.zeros((cap, 4), dtype=self.torch.int64, device=self.dev)
        out[: hi - lo] = buf[lo:hi]
        return out

    def get(self, buf, i):
        return limbs_to_fr(buf[i].cpu().numpy().view(np.uint64))

    def put(self, buf, i, value):
        buf[i] = self.torch.from_numpy(fr_to_limbs(value).view(np.int64)).to(self.dev)

    # -- transforms / commitments
    def ifft(self, evals, log_n, cap):
        n = 1 << log_n
        out = self.torch.zeros((cap, 4), dtype=self.torch.int64, device=self.dev)
        out[:n] = evals[:n]
        self.ctx.ntt_dev(out, log_n, True, False)
        return out

    def effective_len(self, buf, n):
        return self.ctx.poly_effective_len_dev(buf, n)

    def add_blinders(self, poly, blinders):
        b = np.stack([fr_to_limbs(x) for x in blinders])
        self.ctx.poly_add_blinders_dev(poly.data, poly.len, b)
        poly.len += len(blinders)

    def commit(self, poly):
        if poly.len == 0:
            return None
        xy, inf = self.kzg.commit_dev(poly.data, poly.len)
        return point_to_ints(xy, inf)

    def commit_many(self, polys):
        """One PC::commit call for several polynomials: the MSMs are pipelined on the device."""
        live = [p for p in polys if p.len]
        res = iter(self.kzg.commit_many_dev([p.data for p in live], [p.len for p in live])) if live else iter(())
        out = []
        for p in polys:
            if p.len:
                xy, inf = next(res)
                out.append(point_to_ints(xy, inf))
            else:
                out.append(None)
        return out

    # -- argument math
    def z1_poly(self, log_n, beta, gamma, a, b, c, s1, s2, s3, cap):
        n = 1 << log_n
        z = self.torch.zeros((cap, 4), dtype=self.torch.int64, device=self.dev)
        self.ctx.z1_evals_dev(log_n, fr_to_limbs(beta), fr_to_limbs(gamma), a, b, c, s1, s2, s3, z)
        if self.ctx.grand_product_failed():
            raise ZeroDivisionError("compute_z1_poly: zero denominator")
        self.ctx.ntt_dev(z, log_n, True, False)
        return z

    def z2_poly(self, log_n, delta, eps, f, t, h1, h2, cap):
        z = self.torch.zeros((cap, 4), dtype=self.torch.int64, device=self.dev)
        self.ctx.z2_evals_dev(log_n, fr_to_limbs(delta), fr_to_limbs(eps), f, t, h1, h2, z)
        if self.ctx.grand_product_failed():
            raise ZeroDivisionError("compute_z2_poly: zero denominator")
        self.ctx.ntt_dev(z, log_n, True, False)
        return z

    def extend_prover_key(self, log_n, polys):
        """keys/mod.rs:78-146: 10 coset tables on 4n + l_1_coset (x_coset / zh_coset are computed in the kernel)."""
        from .prover_ops import EPK_ORDER
        n4 = 4 << log_n
        epk = {}
        for name in EPK_ORDER[:-1]:
            p = polys[name]
            buf = self.torch.zeros((n4, 4), dtype=self.torch.int64, device=self.dev)
            buf[: p.len] = p.data[: p.len]
            self.ctx.ntt_dev(buf, log_n + 2, False, True, length=p.len)
            epk[name] = buf
        epk["l1"] = self.ctx.l1_coset_dev(log_n, self.torch.empty((n4, 4), dtype=self.torch.int64, device=self.dev))
        return epk

    def quotient(self, log_n, epk, ch, polys):
        """quotient_poly::compute: 9 coset FFTs, the fused kernel, one coset iFFT; returns 4n coefficients."""
        from .prover_ops import EPK_ORDER, WIT_ORDER
        n4 = 4 << log_n
        wit = []
        for name in WIT_ORDER:
            p = polys[name]
            buf = self.torch.zeros((n4, 4), dtype=self.torch.int64, device=self.dev)
            buf[: p.len] = p.data[: p.len]
            self.ctx.ntt_dev(buf, log_n + 2, False, True, length=p.len)
            wit.append(buf)
        chal = np.stack([fr_to_limbs(x) for x in ch])
        out = self.torch.empty((n4, 4), dtype=self.torch.int64, device=self.dev)
        self.ctx.quotient_evals_dev(log_n, chal, wit, [epk[k] for k in EPK_ORDER], out)
        del wit
        self.ctx.ntt_dev(out, log_n + 2, True, True)
        return out

    def evaluate(self, poly, z):
        return limbs_to_fr(self.ctx.poly_eval_dev(poly.data, poly.len, fr_to_limbs(z)))

    def lincomb(self, polys, scalars, cap=None):
        m = max(p.len for p in polys)
        out = self.torch.zeros((cap or m, 4), dtype=self.torch.int64, device=self.dev)
        sc = np.stack([fr_to_limbs(s) for s in scalars])
        self.ctx.poly_lincomb_dev([p.data for p in polys], [p.len for p in polys], sc, out, m)
        return Poly(out, m)

    def divide_linear(self, poly, z):
        quot = self.torch.zeros((max(poly.len - 1, 1), 4), dtype=self.torch.int64, device=self.dev)
        ev = self.ctx.poly_divide_linear_dev(poly.data, poly.len, fr_to_limbs(z), quot)
        return Poly(quot, max(poly.len - 1, 0)), limbs_to_fr(ev)


# ------------------------------------------------------------------------------------------------ setup / prove
def _commit_all(be, polys):
    """PC::commit(ck, [p_0, p_1, ...], None): batched when the backend can pipeline, else one by one."""
    if hasattr(be, "commit_many"):
        return be.commit_many(polys)
    return [be.commit(p) for p in polys]


def _poly_from_evals(be, evals_host, log_n, cap):
    """poly_from_evals: iFFT then DensePolynomial truncation."""
    n = 1 << log_n
    buf = be.ifft(be.from_host(evals_host), log_n, cap)
    return Poly(buf, be.effective_len(buf, n))


def setup(be, circuit):
    """proof_system/setup.rs:42-166 with extend = true.  Returns (ProverKey, VerifierKey)."""
    n, log_n = circuit.n, circuit.log_n
    sel = circuit.selectors
    polys = {name: _poly_from_evals(be, sel[name], log_n, n) for name in ("q_m", "q_l", "q_r", "q_o", "q_c", "q_lookup")}
    for k, name in enumerate(("sigma1", "sigma2", "sigma3")):
        polys[name] = _poly_from_evals(be, circuit.sigma[k], log_n, n)
    polys["q_table"] = _poly_from_evals(be, ints_to_mont_array(table_masks(circuit.table_size, n)), log_n, n)
    commits = dict(zip(VerifierKey.ORDER, _commit_all(be, [polys[name] for name in VerifierKey.ORDER])))
    w = field.root_of_unity(log_n)
    pi_roots = [pow(w, pos, field.R_MOD) for pos in circuit.pi.keys()]
    vk = VerifierKey(n, pi_roots, commits)
    evals = {"sigma1": be.from_host(circuit.sigma[0]), "sigma2": be.from_host(circuit.sigma[1]),
             "sigma3": be.from_host(circuit.sigma[2]), "q_lookup": sel["q_lookup"]}
    epk = be.extend_prover_key(log_n, polys)
    return ProverKey(polys, evals, epk, n, log_n), vk


def prove(be, pk, vk, circuit, blinders, transcript=None, timings=None):
    """proof_system::prove (prove.rs:59-470).  blinders: the 19 field elements the reference draws with F::rand, in
    draw order a(2) b(2) c(2) h1(3) h2(2) z1(3) z2(3) b0 b1.  Returns a Proof."""
    import time
    n, log_n = pk.n, pk.log_n
    assert len(blinders) == 19
    clock = [time.perf_counter()]

    def tick(name):
        """timings[name] += wall time since the previous tick (device work drained first)."""
        if timings is None:
            return
        if hasattr(be, "sync"):
            be.sync()
        now = time.perf_counter()
        timings[name] = timings.get(name, 0.0) + (now - clock[0]) * 1e3
        clock[0] = now

    bl = iter(blinders)
    take = lambda k: [next(bl) for _ in range(k)]
    tr = transcript
    if tr is None or isinstance(tr, str):                      # the `T: TranscriptProtocol` parameter (plonk.rs:39-46)
        tr = TRANSCRIPTS[tr or "merlin"][1]("ZKT Plonk")      # plonk.rs:107
        vk.seed_transcript(tr)
    cap = n + 8
    w_n = field.root_of_unity(log_n)

    # ---- round 1: wires
    tr.append_scalars("pi", list(circuit.pi.values()))          # prove.rs:110
    a_ev, b_ev, c_ev = be.from_host(circuit.a), be.from_host(circuit.b), be.from_host(circuit.c)
    tick("h2d_wires_ms")
    wires = {}
    for name, ev in (("a", a_ev), ("b", b_ev), ("c", c_ev)):
        buf = be.ifft(ev, log_n, cap)
        p = Poly(buf, be.effective_len(buf, n))
        be.add_blinders(p, take(2))
        wires[name] = p
    commits = dict(zip(("a", "b", "c"), _commit_all(be, [wires[k] for k in ("a", "b", "c")])))
    for k in ("a", "b", "c"):
        tr.append_commitment(k + "_commit", commits[k])
    tick("round1_wires_ms")

    # ---- round 2: lookup multisets (host side: prove.rs:145-167, multiset.rs:103-146)
    t_arr = table_multiset_array(circuit.table, circuit.table_size, n)
    f_arr = lookup_f_array(pk.evals["q_lookup"], circuit.c)
    h1_arr, h2_arr = combine_split_arrays(t_arr, f_arr)
    assert h1_arr.shape[0] == n and h2_arr.shape[0] == n
    t_ev, f_ev, h1_ev, h2_ev = (be.from_host(x) for x in (t_arr, f_arr, h1_arr, h2_arr))
    tick("host_lookup_plumbing_ms")
    polys = dict(wires)
    for name, ev, k in (("t", t_ev, 0), ("h1", h1_ev, 3), ("h2", h2_ev, 2)):
        buf = be.ifft(ev, log_n, cap)
        p = Poly(buf, be.effective_len(buf, n))
        if k:
            be.add_blinders(p, take(k))
        polys[name] = p
    commits.update(zip(("t", "h1", "h2"), _commit_all(be, [polys[k] for k in ("t", "h1", "h2")])))
    for k in ("t", "h1", "h2"):
        tr.append_commitment(k + "_commit", commits[k])

    beta, gamma = tr.challenge_scalar("beta"), tr.challenge_scalar("gamma")
    delta, epsilon = tr.challenge_scalar("delta"), tr.challenge_scalar("epsilon")
    assert len({beta, gamma, delta, epsilon}) == 4, "challenges must be different"
    tick("round2_lookup_ms")

    # ---- round 3: grand products
    buf = be.z1_poly(log_n, beta, gamma, a_ev, b_ev, c_ev, pk.evals["sigma1"], pk.evals["sigma2"], pk.evals["sigma3"], cap)
    polys["z1"] = Poly(buf, be.effective_len(buf, n))
    be.add_blinders(polys["z1"], take(3))
    buf = be.z2_poly(log_n, delta, epsilon, f_ev, t_ev, h1_ev, h2_ev, cap)
    polys["z2"] = Poly(buf, be.effective_len(buf, n))
    be.add_blinders(polys["z2"], take(3))
    del a_ev, b_ev, c_ev, f_ev, t_ev, h1_ev, h2_ev
    commits.update(zip(("z1", "z2"), _commit_all(be, [polys[k] for k in ("z1", "z2")])))
    for k in ("z1", "z2"):
        tr.append_commitment(k + "_commit", commits[k])

    tick("round3_grand_products_ms")

    # ---- round 4: quotient
    pi_evals = np.zeros((n, 4), dtype=np.uint64)                # PublicInputs::as_evals (pi.rs:75-82)
    if circuit.pi:
        pi_evals[list(circuit.pi.keys())] = ints_to_mont_array(list(circuit.pi.values()))
    polys["pi"] = _poly_from_evals(be, pi_evals, log_n, n)
    alpha = tr.challenge_scalar("alpha")
    q_buf = be.quotient(log_n, pk.epk, (alpha, beta, gamma, delta, epsilon), polys)
    q_len = be.effective_len(q_buf, 4 * n)
    assert q_len >= 2 * (n + 2), "quotient shorter than 2(n+2): the reference's slice would panic"
    if q_len - 2 * (n + 2) > n + 8:                           # same guard and reason as csrc/prover.cu
        raise ValueError("quotient longer than 3n + 6 coefficients: the division by Z_H was not exact "
                         "(prove.rs:287-292; PC::commit would fail upstream)")
    parts = []
    for lo, hi in ((0, n + 2), (n + 2, 2 * (n + 2)), (2 * (n + 2), q_len)):
        buf = be.slice_copy(q_buf, lo, hi, n + 8)
        parts.append(Poly(buf, be.effective_len(buf, hi - lo)))
    del q_buf
    q_lo, q_mid, q_hi = parts
    b0, b1 = take(2)
    be.put(q_lo.data, q_lo.len, b0); q_lo.len += 1            # q_lo.coeffs.push(b0)
    be.put(q_mid.data, 0, (be.get(q_mid.data, 0) - b0) % field.R_MOD)   # q_mid.coeffs[0] -= b0
    be.put(q_mid.data, q_mid.len, b1); q_mid.len += 1
    be.put(q_hi.data, 0, (be.get(q_hi.data, 0) - b1) % field.R_MOD)
    polys.update(q_lo=q_lo, q_mid=q_mid, q_hi=q_hi)
    commits.update(zip(("q_lo", "q_mid", "q_hi"), _commit_all(be, [polys[k] for k in ("q_lo", "q_mid", "q_hi")])))
    for k in ("q_lo", "q_mid", "q_hi"):
        tr.append_commitment(k + "_commit", commits[k])
    xi = tr.challenge_scalar("xi")
    tick("round4_quotient_ms")

    # ---- round 5: linearisation (linearization_poly.rs:19-121) and openings
    shifted = xi * w_n % field.R_MOD
    zh = (pow(xi, n, field.R_MOD) - 1) % field.R_MOD
    l1 = zh * pow(n * (xi - 1) % field.R_MOD, -1, field.R_MOD) % field.R_MOD
    kp = pk.polys
    ev = {"a": be.evaluate(polys["a"], xi), "b": be.evaluate(polys["b"], xi), "c": be.evaluate(polys["c"], xi),
          "sigma1": be.evaluate(kp["sigma1"], xi), "sigma2": be.evaluate(kp["sigma2"], xi),
          "z1_next": be.evaluate(polys["z1"], shifted), "q_lookup": be.evaluate(kp["q_lookup"], xi),
          "t": be.evaluate(polys["t"], xi), "t_next": be.evaluate(polys["t"], shifted),
          "z2_next": be.evaluate(polys["z2"], shifted), "h1_next": be.evaluate(polys["h1"], shifted),
          "h2": be.evaluate(polys["h2"], xi)}
    a_, b_, c_ = ev["a"], ev["b"], ev["c"]
    al2, al3 = alpha * alpha % field.R_MOD, pow(alpha, 3, field.R_MOD)
    al4, al5 = pow(alpha, 4, field.R_MOD), pow(alpha, 5, field.R_MOD)
    bxi = beta * xi % field.R_MOD
    opd = (1 + delta) % field.R_MOD
    eopd = epsilon * opd % field.R_MOD
    s_z1 = (alpha * (bxi + a_ + gamma) % field.R_MOD * (bxi * field.K1 + b_ + gamma) % field.R_MOD * (bxi * field.K2 + c_ + gamma) + l1 * al2) % field.R_MOD
    s_sigma3 = (-alpha * beta % field.R_MOD * ev["z1_next"] % field.R_MOD * (beta * ev["sigma1"] + a_ + gamma) % field.R_MOD * (beta * ev["sigma2"] + b_ + gamma)) % field.R_MOD
    s_z2 = (al3 * opd % field.R_MOD * (epsilon + ev["q_lookup"] * c_) % field.R_MOD * (eopd + ev["t"] + delta * ev["t_next"]) + al4 * l1) % field.R_MOD
    s_h1 = (-al3 * ev["z2_next"] % field.R_MOD * (eopd + ev["h2"] + delta * ev["h1_next"])) % field.R_MOD
    s_qtable = al5 * ev["t"] % field.R_MOD
    xn2 = (zh + 1) * xi % field.R_MOD * xi % field.R_MOD                           # xi^(n+2)
    terms = [(kp["q_m"], a_ * b_ % field.R_MOD), (kp["q_l"], a_), (kp["q_r"], b_), (kp["q_o"], c_), (kp["q_c"], 1),
             (polys["z1"], s_z1), (kp["sigma3"], s_sigma3), (polys["z2"], s_z2), (polys["h1"], s_h1),
             (kp["q_table"], s_qtable), (q_lo, (-zh) % field.R_MOD), (q_mid, (-zh * xn2) % field.R_MOD), (q_hi, (-zh * xn2 % field.R_MOD * xn2) % field.R_MOD)]
    r_poly = be.lincomb([t[0] for t in terms], [t[1] for t in terms])
    r_poly.len = be.effective_len(r_poly.data, r_poly.len)

    for k in Proof.EVALS:
        tr.append_scalar(k + "_eval", ev[k])
    eta = tr.challenge_scalar("eta")

    def witness_of(plist, point):
        comb = be.lincomb(plist, [pow(eta, i, field.R_MOD) for i in range(len(plist))])
        wit, _ = be.divide_linear(comb, point)
        wit.len = be.effective_len(wit.data, wit.len)
        return wit

    w1 = witness_of([r_poly, polys["a"], polys["b"], polys["c"], kp["sigma1"], kp["sigma2"], kp["q_lookup"], polys["t"], polys["h2"]], xi)
    w2 = witness_of([polys["z1"], polys["z2"], polys["t"], polys["h1"]], shifted)
    aw, saw = _commit_all(be, [w1, w2])                          # the two openings are independent MSMs
    tick("round5_linearisation_openings_ms")
    return Proof(commits, aw, saw, ev)


# ------------------------------------------------------------------------------------------------ C++ round driver
class NativeProver:
    """The same setup / prove behind the two C-ABI calls zkb_plonk_setup / zkb_plonk_prove (csrc/prover.cu): what a
    Rust FFI crate would call with the composer's vectors.  The committer key must be resident in `ctx`."""

    def __init__(self, ctx, circuit, key_files=None, key_polys=None):
        """key_files = (pk_path, vk_path): load the reference CLI's ProverKey / VerifierKey files (zkb_plonk_load_keys)
        instead of running setup on the circuit's selector columns (zkb_plonk_setup).  key_polys = {name: (len, 4)
        Montgomery coefficients}: the ProverKey's polynomials already in memory (zkb_plonk_pk_from_polys; the verifier
        key's commitments are recomputed)."""
        import ctypes
        self.ctx, self.circuit = ctx, circuit
        lib = ctx._lib
        h = ctypes.c_void_p()
        if key_polys is not None:
            from .keyfile import PK_ORDER
            bufs = [np.ascontiguousarray(key_polys[name], dtype=np.uint64).reshape(-1, 4) for name in PK_ORDER]
            keep = [b if b.shape[0] else np.zeros((1, 4), dtype=np.uint64) for b in bufs]
            ptrs = (ctypes.c_void_p * 10)(*[b.ctypes.data for b in keep])
            lens = (ctypes.c_size_t * 10)(*[b.shape[0] for b in bufs])
            pos = list(circuit.pi.keys())
            PP = (ctypes.c_size_t * max(len(pos), 1))(*pos)
            ctx._check(lib.zkb_plonk_pk_from_polys(ctx._h, circuit.log_n, ptrs, lens, circuit.table_size, PP, len(pos), None, None,
                                                   ctypes.byref(h)))
            self._pk = h
            return
        if key_files is not None:
            import os
            ctx._check(lib.zkb_plonk_load_keys(ctx._h, os.fsencode(key_files[0]), os.fsencode(key_files[1]), circuit.table_size,
                                               ctypes.byref(h)))
            self._pk = h
            return
        sel = [np.ascontiguousarray(circuit.selectors[k], dtype=np.uint64) for k in ("q_m", "q_l", "q_r", "q_o", "q_c", "q_lookup")]
        sig = [np.ascontiguousarray(x, dtype=np.uint64) for x in circuit.sigma]
        self._keep = sel + sig
        S = (ctypes.c_void_p * 6)(*[x.ctypes.data for x in sel])
        G = (ctypes.c_void_p * 3)(*[x.ctypes.data for x in sig])
        pos = list(circuit.pi.keys())
        PP = (ctypes.c_size_t * max(len(pos), 1))(*pos)
        ctx._check(lib.zkb_plonk_setup(ctx._h, circuit.log_n, S, G, circuit.table_size, PP, len(pos), ctypes.byref(h)))
        self._pk = h

    def save_keys(self, pk_path, vk_path):
        """Write the ProverKey / VerifierKey files `compile` writes (bin/src/main.rs:106-112)."""
        import os
        enc = lambda p: None if p is None else os.fsencode(p)
        self.ctx._check(self.ctx._lib.zkb_plonk_save_keys(self.ctx._h, self._pk, enc(pk_path), enc(vk_path)))

    def set_lookup_mode(self, mode):
        """Where round 2's witness plumbing runs (zkb_plonk_pk_set_lookup_mode): 0 = on the device when more than n / 8 rows are
        lookup gates (default), 1 = sparse on a host thread, 2 = on the device.  The proof bytes do not depend on it."""
        self.ctx._check(self.ctx._lib.zkb_plonk_pk_set_lookup_mode(self._pk, int(mode)))

    def set_transcript(self, name):
        """"merlin" (default) or "ethereum": which TranscriptProtocol later proofs use (zkb_plonk_pk_set_transcript)."""
        self.ctx._check(self.ctx._lib.zkb_plonk_pk_set_transcript(self._pk, TRANSCRIPTS[name][0]))

    def vk(self):
        import ctypes
        xy = np.zeros((10, self.ctx.aff_words), dtype=np.uint64)
        inf = (ctypes.c_int * 10)()
        self.ctx._check(self.ctx._lib.zkb_plonk_vk_commitments(self._pk, xy.ctypes.data_as(ctypes.c_void_p), inf))
        commits = {name: point_to_ints(xy[k], bool(inf[k])) for k, name in enumerate(VerifierKey.ORDER)}
        w = field.root_of_unity(self.circuit.log_n)
        return VerifierKey(self.circuit.n, [pow(w, p, field.R_MOD) for p in self.circuit.pi.keys()], commits)

    def set_wiring(self, wiring=None):
        """zkb_plonk_pk_set_wiring: the key keeps the circuit's wire maps; prove_bytes(..., from_vars=True) then uploads the
        variable assignment instead of the three wire vectors (ProvingComposer::wire_evals on the device)."""
        import ctypes
        w = np.ascontiguousarray(self.circuit.wiring if wiring is None else wiring, dtype=np.uint32)
        assert w.shape == (3, self.circuit.n)
        vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        self.ctx._check(self.ctx._lib.zkb_plonk_pk_set_wiring(self.ctx._h, self._pk, vp(w[0]), vp(w[1]), vp(w[2])))

    def prove_bytes(self, blinders, timings=False, from_vars=False):
        import ctypes
        c = self.circuit
        if from_vars:
            return self._prove_vars(blinders, timings)
        a, b, cc = (np.ascontiguousarray(x, dtype=np.uint64) for x in (c.a, c.b, c.c))
        table, pi = self._table_pi()
        bl = ints_to_mont_array(blinders)
        out = np.zeros(self.ctx.proof_bytes(), dtype=np.uint8)
        tm = (ctypes.c_float * 8)() if timings else None
        vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        self.ctx._check(self.ctx._lib.zkb_plonk_prove(self.ctx._h, self._pk, vp(a), vp(b), vp(cc), vp(table), len(c.table), vp(pi),
                                                      vp(bl), vp(out), tm))
        raw = out.tobytes()
        if timings:
            names = ("h2d_wires_ms", "round1_wires_ms", "host_lookup_plumbing_ms", "round2_lookup_ms", "round3_grand_products_ms",
                     "round4_quotient_ms", "round5_linearisation_openings_ms", "total_ms")
            return raw, dict(zip(names, [float(x) for x in tm]))
        return raw

    def _table_pi(self):
        """The lookup table and the public inputs as Montgomery limbs (what the C ABI takes).  The circuit holds them as Python
        integers; converting 1024 table entries costs ~1.8 ms of interpreter time, so it is done once per prover, not per proof
        (a caller on the reference's side hands its field elements over as they are).  Assign a new list to change the table."""
        c = self.circuit
        key = (id(c.table), len(c.table), tuple(c.pi.items()) if c.pi else ())    # another list object = another table
        if getattr(self, "_tp_key", None) != key:
            table = ints_to_mont_array(c.table) if c.table else np.zeros((1, 4), dtype=np.uint64)
            pi = ints_to_mont_array(list(c.pi.values())) if c.pi else np.zeros((1, 4), dtype=np.uint64)
            self._tp_key, self._tp = key, (table, pi)
        return self._tp

    def _prove_vars(self, blinders, timings):
        import ctypes
        c = self.circuit
        vals = np.ascontiguousarray(c.var_values, dtype=np.uint64)
        table, pi = self._table_pi()
        bl = ints_to_mont_array(blinders)
        out = np.zeros(self.ctx.proof_bytes(), dtype=np.uint8)
        tm = (ctypes.c_float * 8)() if timings else None
        vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        self.ctx._check(self.ctx._lib.zkb_plonk_prove_vars(self.ctx._h, self._pk, vp(vals), vals.shape[0], vp(table), len(c.table),
                                                           vp(pi), vp(bl), vp(out), tm))
        raw = out.tobytes()
        if timings:
            names = ("h2d_wires_ms", "round1_wires_ms", "host_lookup_plumbing_ms", "round2_lookup_ms", "round3_grand_products_ms",
                     "round4_quotient_ms", "round5_linearisation_openings_ms", "total_ms")
            return raw, dict(zip(names, [float(x) for x in tm]))
        return raw

    def close(self):
        if getattr(self, "_pk", None):
            self.ctx._lib.zkb_plonk_pk_destroy(self.ctx._h, self._pk)
            self._pk = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def proof_from_bytes(raw):
    """Inverse of Proof.to_bytes for the verifier: decompress the 13 G1 points (y from x, sign from the flag bit)."""
    nb = field.FQ_BYTES                                            # 32 (BN254: 802-byte proofs) or 48 (BLS12: 1010 bytes)
    assert len(raw) == 13 * nb + 2 + 12 * 32
    top = 8 * nb - 2

    def point(b):
        v = int.from_bytes(b, "little")
        if (v >> top) & 1:
            return None
        x = v & ((1 << top) - 1)
        y = field.sqrt_q(x * x * x + field.CURVE_B)
        assert y is not None, "x is not on the curve"
        if (y > (field.Q_MOD - y) % field.Q_MOD) != bool((v >> (top + 1)) & 1):
            y = (field.Q_MOD - y) % field.Q_MOD
        return (x, y)

    pts = [point(raw[nb * k: nb * k + nb]) for k in range(11)]
    aw, saw = point(raw[11 * nb: 12 * nb]), point(raw[12 * nb + 1: 13 * nb + 1])
    assert raw[12 * nb] == 0 and raw[13 * nb + 1] == 0
    e0 = 13 * nb + 2
    evals = [int.from_bytes(raw[e0 + 32 * k: e0 + 32 + 32 * k], "little") for k in range(12)]
    return Proof(dict(zip(Proof.COMMITS, pts)), aw, saw, dict(zip(Proof.EVALS, evals)))
