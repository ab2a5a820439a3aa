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
