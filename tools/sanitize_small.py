"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from zkt_plonk_b200 import prover, synthetic

ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
rng = np.random.default_rng(1)
def rnd(n, w=4):
    a = rng.integers(0, 2**62, size=(n, w), dtype=np.int64); a[:, w - 1] &= (1 << 58) - 1
    return torch.from_numpy(a).cuda()
# NTT: 1, 2 and 3 passes, both table kinds
for direct in (True, False):
    ctx.ntt_set_direct_tables(direct)
    for log_n in (3, 9, 13, 16):
        d = rnd(1 << log_n)
        for inv, cos in ((0, 0), (0, 1), (1, 1)):
            ctx.ntt_dev(d, log_n, bool(inv), bool(cos), length=(1 << log_n) - 1)
ctx.ntt_set_direct_tables(True)
d = rnd(1 << 23); ctx.ntt_dev(d, 23, False, True); ctx.ntt_dev(d, 23, True, True); del d
# MSM: plain / forced windows / fixed-base / batch / skew
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
n = 3000
P = torch.empty((n, 8), dtype=torch.int64, device="cuda"); ctx.g1_fixed_base_mul_dev(G, rnd(n), n, P)
ctx.srs_load(P)
s = rnd(n); s[:700] = 0; s[700:1500, 1:] = 0; s[700:1500, 0] = 1
for c in (0, 3, 9, 16):
    ctx.set_msm_window(c); ctx.msm(s)
ctx.set_msm_window(0)
for c in (0, 5, 14):
    ctx.srs_precompute(c); ctx.msm(s); ctx.msm(s[:100].contiguous(), offset=7)
ctx.commit_batch_dev([rnd(n), rnd(100), rnd(n)], [n, 100, n - 1])
ctx.srs_precompute(-1)
ctx.commit_batch_dev([rnd(n), rnd(5)], [n, 5])
# prover end to end (covers grand products, quotient, poly utilities)
circ = synthetic.make_circuit(7, seed=2, table_size=16)
tau = 123456789
pw = [pow(tau, i, prover.P) for i in range(circ.n + 8)]
k = np.array([[(v >> (64 * j)) & (2**64 - 1) for j in range(4)] for v in pw], dtype=np.uint64)
srs = torch.empty((circ.n + 8, 8), dtype=torch.int64, device="cuda")
ctx.g1_fixed_base_mul_dev(G, torch.from_numpy(k.view(np.int64)).cuda(), circ.n + 8, srs)
kzg = z.GpuKZG10(ctx); kzg.load_committer_key(srs); ctx.srs_precompute(0)
be = prover.GpuBackend(kzg)
pk, vk = prover.setup(be, circ)
proof = prover.prove(be, pk, vk, circ, list(range(1, 20)))
torch.cuda.synchronize()
print("sanitize_small ok", len(proof.to_bytes()))
