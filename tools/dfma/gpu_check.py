"""GPU: the FP64-pipe Montgomery product (csrc/ff52.cuh) against big-int arithmetic, and its throughput alone and
mixed with the integer product in alternating warps (zkb_bench_int modes 2 / 4 / 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import zkt_plonk_b200 as z
from oracle import cref
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
Q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
for field, p, F in ((1, Q, cref.FQ), (0, R, cref.FR)):
    n = 20000
    a = cref.rand_fe(F, n, 11); b = cref.rand_fe(F, n, 12)
    a[0] = cref.ints_to_limbs([p - 1])[0]; b[0] = a[0]; a[1] = 0
    got = cref.limbs_to_ints(ctx.fp_binop(field, 7, a, b))
    ai, bi = cref.limbs_to_ints(a), cref.limbs_to_ints(b)
    inv = pow(1 << 260, -1, p)
    bad = sum(1 for x, y, g in zip(ai, bi, got) if g != x * y * inv % p)
    print("field", field, "fmul52 checked", n, "bad", bad, flush=True)
    assert bad == 0
for mode, name in ((0, "IMAD/s"), (3, "DFMA/s"), (2, "Fq products/s, integer pipe (ff.cuh)"), (4, "Fq products/s, FP64 pipe (ff52.cuh)"),
                   (5, "Fq products/s, alternating warps on both pipes")):
    print(f"{name}: {ctx.bench_int(mode):.4e}", flush=True)
for mode, name in ((6, "pairs/s of (IMAD + DFMA) in one thread"), (7, "groups/s of (IMAD + ADD + XOR)"), (8, "groups/s of (DFMA + ADD + XOR)")):
    print(f"{name}: {ctx.bench_int(mode):.4e}", flush=True)
for field, F in ((1, cref.FQ), (0, cref.FR)):
    n = 50000
    a = cref.to_mont(F, cref.rand_fe(F, n, 31)); b = cref.to_mont(F, cref.rand_fe(F, n, 32))
    pm1 = cref.normalize(F, np.full((1, 4), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64))
    a[0] = pm1[0]; b[0] = pm1[0]; a[1] = 0; b[2] = 0; a[3] = pm1[0]; b[3, :] = 0; b[3, 0] = 1
    a[4] = np.array([0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0, 0], dtype=np.uint64); b[4] = a[4]
    a[5] = np.array([0, 0, 0xFFFFFFFFFFFFFFFF, 0x0FFFFFFFFFFFFFFF], dtype=np.uint64); b[5] = a[4]
    same = np.array_equal(ctx.fp_binop(field, 8, a, b), ctx.fp_binop(field, 0, a, b))
    print("field", field, "Karatsuba product == CIOS product:", same, flush=True)
    assert same
print(f"Fq products/s, Karatsuba (ff_kara.cuh): {ctx.bench_int(9):.4e}", flush=True)
print(f"Fq products/s, CIOS (ff.cuh): {ctx.bench_int(2):.4e}", flush=True)
