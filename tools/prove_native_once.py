"""setup + two proofs through the C++ driver at the given size (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from zkt_plonk_b200 import prover, synthetic
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
P = prover.P; TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P
n = 1 << log_n
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
circ = synthetic.make_circuit(log_n, seed=1)
pw = np.empty(n + 8, dtype=object); x = 1
for i in range(n + 8):
    pw[i] = x; x = x * TAU % P
k = np.empty((n + 8, 4), dtype=np.uint64)
for j in range(4):
    k[:, j] = ((pw >> (64 * j)) & ((1 << 64) - 1)).astype(np.uint64)
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
srs = torch.empty((n + 8, 8), dtype=torch.int64, device="cuda")
ctx.g1_fixed_base_mul_dev(G, torch.from_numpy(k.view(np.int64)).cuda(), n + 8, srs)
kzg = z.GpuKZG10(ctx); kzg.load_committer_key(srs); ctx.srs_precompute(0)
native = prover.NativeProver(ctx, circ)
for r in range(2):
    raw, tm = native.prove_bytes(list(range(1000 + r, 1019 + r)), timings=True)
torch.cuda.synchronize()
print("MARK launches_total", ctx.launch_count(), tm)
