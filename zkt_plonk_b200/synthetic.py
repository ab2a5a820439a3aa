"""Synthetic ConstraintSystem-shaped circuits for the full-prove configs (BASELINE.json configs[0], configs[3]).

Produces what SetupComposer / ProvingComposer hold after synthesis and pad_to(n): selector evaluations, the three
sigma evaluation vectors of a copy-constraint permutation (permutation/mod.rs:40-178 encodes positions as
k_w * omega^row with k = 1, K1 = 7, K2 = 13), a lookup table, the wire values and the public inputs.  Gates are
random but satisfiable: additions, multiplications, constants, ~1 % lookup rows against a 1024-entry table and a
few public-input rows, wired together by reusing earlier outputs as later inputs.  Pure host code (Python ints).
"""
import numpy as np

from . import field
from .prover import Circuit, ints_to_mont_array



def make_circuit(log_n, seed=0, table_size=None, n_public=4, lookup_frac=0.01, fill=0.9):
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    if table_size is None:
        table_size = min(1024, n // 4)
    table = [int(v) for v in rng.integers(1, 1 << 62, size=table_size - 1)]
    table = list(dict.fromkeys(table))                       # IndexSet: distinct, insertion ordered
    used = max(8, int(n * fill))                              # rows >= used are padding (all selectors 0, Variable::Zero)
    kind = rng.random(used)
    n_lookup = max(1, int(used * lookup_frac)) if lookup_frac > 0 else 0
    a, b, c = [0] * n, [0] * n, [0] * n
    var_a, var_b, var_c = [0] * n, [0] * n, [0] * n           # variable ids; 0 = Variable::Zero
    q_m, q_l, q_r, q_o, q_c, q_lk = ([0] * n for _ in range(6))
    values = [0]                                              # variable id -> value
    pi = {}
    rnd = rng.integers(0, 1 << 62, size=(used, 4))
    pick = rng.integers(0, 1 << 30, size=(used, 2))

    def new_var(v):
        values.append(v % field.R_MOD)
        return len(values) - 1

    def old_or_new(i, k):
        if len(values) > 8 and rnd[i, k] & 3:                 # 75 %: reuse an earlier variable (copy constraint)
            return 1 + int(pick[i, k]) % (len(values) - 1)
        return new_var(int(rnd[i, 2]) * int(rnd[i, 3]) + k)

    for i in range(used):
        va, vb = old_or_new(i, 0), old_or_new(i, 1)
        x, y = values[va], values[vb]
        if i < n_public:                                      # public input row: q_l * a + PI = 0
            q_l[i] = 1
            pi[i] = (-x) % field.R_MOD
            vc = 0
        elif i < n_public + n_lookup:                         # lookup row: c in table, arithmetic part c - c = 0
            vc = new_var(table[int(pick[i, 0]) % len(table)])
            q_lk[i] = 1
        elif kind[i] < 0.45:                                  # a + b - c = 0
            q_l[i], q_r[i], q_o[i] = 1, 1, field.R_MOD - 1
            vc = new_var(x + y)
        elif kind[i] < 0.9:                                   # a * b - c = 0
            q_m[i], q_o[i] = 1, field.R_MOD - 1
            vc = new_var(x * y)
        else:                                                 # 3a + 5b + k - c = 0
            k = int(rnd[i, 2])
            q_l[i], q_r[i], q_c[i], q_o[i] = 3, 5, k, field.R_MOD - 1
            vc = new_var(3 * x + 5 * y + k)
        var_a[i], var_b[i], var_c[i] = va, vb, vc
        a[i], b[i], c[i] = x, y, values[vc]

    return _finish(log_n, (var_a, var_b, var_c), (a, b, c), (q_m, q_l, q_r, q_o, q_c, q_lk), table, table_size, pi, values)


def _finish(log_n, var_ids, wires, sels, table, table_size, pi, values=None):
    """sigma evaluations from the variable ids (positions of the same variable form one cycle) + Montgomery arrays."""
    n = 1 << log_n
    var_a, var_b, var_c = var_ids
    a, b, c = wires
    q_m, q_l, q_r, q_o, q_c, q_lk = sels
    w = field.root_of_unity(log_n)
    roots = [1] * n
    for i in range(1, n):
        roots[i] = roots[i - 1] * w % field.R_MOD
    ks = (1, field.K1, field.K2)
    var_of = np.array(list(var_a) + list(var_b) + list(var_c), dtype=np.int64)  # position p = wire * n + row
    order = np.argsort(var_of, kind="stable")
    sorted_vars = var_of[order]
    nxt = np.empty(3 * n, dtype=np.int64)
    start = 0
    bounds = np.flatnonzero(np.diff(sorted_vars)) + 1
    for end in list(bounds) + [3 * n]:
        grp = order[start:end]
        nxt[grp] = np.roll(grp, -1)
        start = end
    sig = [ks[int(p) // n] * roots[int(p) % n] % field.R_MOD for p in nxt]
    sigma = tuple(ints_to_mont_array(sig[k * n:(k + 1) * n]) for k in range(3))

    selectors = {"q_m": ints_to_mont_array(q_m), "q_l": ints_to_mont_array(q_l), "q_r": ints_to_mont_array(q_r),
                 "q_o": ints_to_mont_array(q_o), "q_c": ints_to_mont_array(q_c), "q_lookup": ints_to_mont_array(q_lk)}
    wiring = np.array([list(var_a), list(var_b), list(var_c)], dtype=np.uint32)
    var_values = ints_to_mont_array(values) if values is not None else None
    return Circuit(log_n, selectors, sigma, table, table_size, ints_to_mont_array(a), ints_to_mont_array(b),
                   ints_to_mont_array(c), pi, wiring=wiring, var_values=var_values)


def make_edge_circuit(log_n, kind, seed=0, table_size=4):
    """Degenerate shapes the reference handles and a prover must too (add_blinders_to_poly with len < k, prove.rs:472-483):
    "no_lookup"  -- empty lookup table and no lookup rows: t = f = h1 = h2 = 0, so h1 and h2 are ZERO polynomials (len 0)
                    before their 3 blinders are appended;
    "const_wire" -- wire a holds the same value on every row of the domain (no padding rows): a(X) has ONE coefficient
                    before its 2 blinders; also no lookups."""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    vals = [0]

    def new_var(v):
        vals.append(v % field.R_MOD)
        return len(vals) - 1

    q_m, q_l, q_r, q_o, q_c, q_lk = ([0] * n for _ in range(6))
    var_a, var_b, var_c = [0] * n, [0] * n, [0] * n
    a, b, c = [0] * n, [0] * n, [0] * n
    pi = {}
    if kind == "no_lookup":
        used = max(8, n - n // 4)
        for i in range(used):
            va = new_var(int(rng.integers(1, 1 << 62))) if i < 4 or i % 3 else var_c[i - 1]
            vb = new_var(int(rng.integers(1, 1 << 62)))
            x, y = vals[va], vals[vb]
            if i < 2:
                q_l[i] = 1
                pi[i] = (-x) % field.R_MOD
                vc = 0
            elif i % 2:
                q_m[i], q_o[i] = 1, field.R_MOD - 1
                vc = new_var(x * y)
            else:
                q_l[i], q_r[i], q_o[i] = 1, 1, field.R_MOD - 1
                vc = new_var(x + y)
            var_a[i], var_b[i], var_c[i] = va, vb, vc
            a[i], b[i], c[i] = x, y, vals[vc]
        table = []
    elif kind == "const_wire":
        five = new_var(5)
        for i in range(n):                                      # every row: 5 * b - c = 0, a is the same variable everywhere
            vb = new_var(int(rng.integers(1, 1 << 62)))
            vc = new_var(5 * vals[vb])
            q_m[i], q_o[i] = 1, field.R_MOD - 1
            var_a[i], var_b[i], var_c[i] = five, vb, vc
            a[i], b[i], c[i] = 5, vals[vb], vals[vc]
        table = []
    else:
        raise ValueError(kind)
    return _finish(log_n, (var_a, var_b, var_c), (a, b, c), (q_m, q_l, q_r, q_o, q_c, q_lk), table, table_size, pi, vals)


def check_gates(circ):
    """What constraint_system/helper.rs::check_gate verifies: every row satisfies its arithmetic gate and every
    lookup row's output is in the table."""
    from .prover import mont_array_to_ints
    s = {k: mont_array_to_ints(v) for k, v in circ.selectors.items()}
    a, b, c = (mont_array_to_ints(x) for x in (circ.a, circ.b, circ.c))
    tset = set(circ.table) | {0}
    for i in range(circ.n):
        lhs = (s["q_m"][i] * a[i] * b[i] + s["q_l"][i] * a[i] + s["q_r"][i] * b[i] + s["q_o"][i] * c[i] + s["q_c"][i]
               + circ.pi.get(i, 0)) % field.R_MOD
        if lhs:
            return False
        if s["q_lookup"][i] and (s["q_lookup"][i] * c[i] % field.R_MOD) not in tset:
            return False
    return True
