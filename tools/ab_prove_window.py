"""A/B of the fixed-base window size in the context of a whole proof (the cost model behind zkb_srs_precompute(0) is fitted
to stand-alone MSMs; inside a proof the reduction tails of all but the last commitment of a batch are hidden).

  python tools/ab_prove_window.py 18 17 18 19 20        # log_n, then the window sizes to try (0 = cost model)
One JSON line per window size: median wall time of zkb_plonk_prove; every proof must have the same bytes."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from zkt_plonk_b200 import prover, synthetic

log_n = int(sys.argv[1])
windows = [int(x) for x in sys.argv[2:]] or [0]
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
bl = list(range(1000, 1019))
ref = native.prove_bytes(bl)
for c in windows:
    ctx.srs_precompute(c)
    ts = []
    for r in range(11):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        raw = native.prove_bytes(bl)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert raw == ref
    ts = ts[2:]
    print(json.dumps({"log_n": log_n, "window_bits": c, "prove_ms_median": statistics.median(ts), "prove_ms_min": min(ts)}), flush=True)
