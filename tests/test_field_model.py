"""CPU-only: the limb-level model of the device field routines (tools/sqr/model.py) -- the CIOS product, the dedicated
triangular squaring, and the sums of 2-4 products under ONE Montgomery reduction (csrc/ff.cuh: fmul, fsqr_tri, fmadd2,
fmaddn) -- replayed chain by chain with every dropped carry asserted to be zero, against Python integers."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("ff_model", os.path.join(ROOT, "tools", "sqr", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_carry_chains_of_the_device_field_routines():
    m = _model()
    rnd = random.Random(123)
    for p in (m.FR, m.FQ):
        rinv = pow(1 << 256, -1, p)
        worst = [0, 1, p - 1, p - 2, int("ffffffff" * 8, 16) % p, int("80000000" * 8, 16) % p, (1 << 253) + 12345]
        vals = worst + [rnd.randrange(p) for _ in range(300)]
        for a in vals:
            b, c, d = rnd.randrange(p), rnd.choice(vals), rnd.randrange(p)
            assert m.fmul(a, b, p) == a * b * rinv % p
            assert m.fsqr(a, p) == a * a * rinv % p
            assert m.fmul2(a, b, c, d, p) == (a * b + c * d) * rinv % p
        for n in (3, 4):
            assert m.fmaddn([(p - 1, p - 1)] * n, p) == n * (p - 1) ** 2 * rinv % p
            prs = [(rnd.choice(vals), rnd.choice(vals)) for _ in range(n)]
            assert m.fmaddn(prs, p) == sum(u * v for u, v in prs) * rinv % p
