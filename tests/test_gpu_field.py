"""GPU: device field library (8 x 32-bit limbs, IMAD.WIDE carry chains) bit-exact against the oracle."""
import numpy as np
import pytest

from oracle import cref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field", [cref.FR, cref.FQ])
def test_field_ops_bit_exact(ctx, field):
    n = 4096
    a = cref.to_mont(field, cref.rand_fe(field, n, 11))
    b = cref.to_mont(field, cref.rand_fe(field, n, 12))
    pm1 = cref.normalize(field, np.full((1, 4), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64))   # some large value
    a[0], b[0] = 0, 0
    a[1], b[1] = pm1[0], pm1[0]
    a[2] = 0
    for op in (0, 1, 2, 3):
        assert np.array_equal(ctx.fp_binop(field, op, a, b), cref.binop(field, op, a, b)), op
    nz = a[3:260]
    assert np.array_equal(ctx.fp_binop(field, 4, nz), cref.binop(field, 4, nz))
    canon = cref.rand_fe(field, n, 13)
    assert np.array_equal(ctx.fp_binop(field, 5, canon), cref.to_mont(field, canon))
    assert np.array_equal(ctx.fp_binop(field, 6, a), cref.from_mont(field, a))


@pytest.mark.parametrize("field,p", [(cref.FR, 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001),
                                     (cref.FQ, 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47)])
def test_fp64_pipe_montgomery_product(ctx, field, p):
    """csrc/ff52.cuh (experimental): the Montgomery product built on FP64 fused multiply-adds (52-bit limbs,
    R = 2^260) returns a * b * 2^-260 mod p exactly.  Checked against Python big ints (the definition)."""
    n = 2048
    a, b = cref.rand_fe(field, n, 21), cref.rand_fe(field, n, 22)
    a[0] = cref.ints_to_limbs([p - 1])[0]
    b[0] = a[0]
    a[1] = 0
    got = cref.limbs_to_ints(ctx.fp_binop(field, 7, a, b))
    inv = pow(1 << 260, -1, p)
    for x, y, g in zip(cref.limbs_to_ints(a), cref.limbs_to_ints(b), got):
        assert g == x * y * inv % p
