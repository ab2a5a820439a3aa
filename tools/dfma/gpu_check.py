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
