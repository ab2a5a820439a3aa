import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
import zkt_plonk_b200 as z
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
def rnd(n, w=4):
    a = torch.randint(0, 2**62, (n, w), dtype=torch.int64, device="cuda"); a[:, w - 1] &= (1 << 58) - 1
    return a
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
for log_n in (18, 20):
    n = 1 << log_n
    P = torch.empty((n, 8), dtype=torch.int64, device="cuda"); ctx.g1_fixed_base_mul_dev(G, rnd(n), n, P)
    ctx.srs_load(P); ctx.srs_precompute(0)
    polys = [rnd(n) for _ in range(3)]
    for _ in range(2):
        ctx.commit_batch_dev(polys, [n] * 3); [ctx.commit_dev(p, 0, n) for p in polys]
    torch.cuda.synchronize()
    for name, fn in (("sequential x3", lambda: [ctx.commit_dev(p, 0, n) for p in polys]), ("pipelined batch of 3", lambda: ctx.commit_batch_dev(polys, [n] * 3))):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(f"2^{log_n} {name}: {min(ts):.3f} ms")
    a = ctx.commit_batch_dev(polys, [n] * 3); b = [ctx.commit_dev(p, 0, n) for p in polys]
    assert all(np.array_equal(x[0], y[0]) for x, y in zip(a, b))
