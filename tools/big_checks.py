"""Size-independent properties at sizes the oracle cannot reach: plain vs fixed-base MSM agree, shards add up,
NTT round trips at 2^26."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
def rnd(n, w=4):
    a = torch.randint(0, 2**62, (n, w), dtype=torch.int64, device="cuda"); a[:, w - 1] &= (1 << 58) - 1
    return a
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
for log_n in (22, 24):
    n = 1 << log_n
    P = torch.empty((n, 8), dtype=torch.int64, device="cuda"); ctx.g1_fixed_base_mul_dev(G, rnd(n), n, P)
    ctx.srs_load(P); s = rnd(n)
    plain, _ = ctx.msm(s)
    t0 = time.time(); ctx.srs_precompute(0); torch.cuda.synchronize(); tb = time.time() - t0
    fixed, _ = ctx.msm(s)
    h = n // 2
    parts = np.stack([ctx.msm_partial(s[:h].contiguous(), 0, h), ctx.msm_partial(s[h:].contiguous(), h, n - h)])
    both, _ = z.sum_partials(parts)
    ok = np.array_equal(plain, fixed) and np.array_equal(plain, both)
    print(f"msm 2^{log_n}: plain == fixed-base == sum of two shards: {ok} (table build {tb:.2f}s, timing {ctx.msm_last_timing()})", flush=True)
    assert ok
    ctx.srs_precompute(-1); del P, s
x = rnd(1 << 26); ref = x.clone()
ctx.ntt_dev(x, 26, False, True); ctx.ntt_dev(x, 26, True, True); torch.cuda.synchronize()
ok = bool(torch.equal(x, ref)); print("ntt 2^26 coset round trip:", ok); assert ok
