"""A/B of round 2's witness plumbing: the sparse host path against the device path (csrc/lookup.cu, zkb_plonk_pk_set_lookup_mode)
inside whole proofs, for a circuit with few lookup gates and one with many.  One JSON line per (lookup share, mode).

  python tools/lookup_ab.py [--log-n 18] [--fracs 0.01 0.6]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from zkt_plonk_b200 import prover, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", dest="log_n", type=int, default=18)
ap.add_argument("--fracs", type=float, nargs="+", default=[0.01, 0.6])
args = ap.parse_args()
P = prover.P
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
n = 1 << args.log_n
pw = np.empty(n + 8, dtype=object)
x = 1
for i in range(n + 8):
    pw[i] = x
    x = x * TAU % P
k = np.empty((n + 8, 4), dtype=np.uint64)
for j in range(4):
    k[:, j] = ((pw >> (64 * j)) & ((1 << 64) - 1)).astype(np.uint64)
srs = torch.empty((n + 8, 8), dtype=torch.int64, device="cuda")
ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), torch.from_numpy(k.view(np.int64)).cuda(), n + 8, srs)
ctx.srs_load(srs); ctx.srs_precompute(0)
blinders = list(range(1000, 1019))
for frac in args.fracs:
    circ = synthetic.make_circuit(args.log_n, seed=1, lookup_frac=frac)
    native = prover.NativeProver(ctx, circ)
    raws = {}
    for mode in (1, 2):
        native.set_lookup_mode(mode)
        walls = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            raws[mode] = native.prove_bytes(blinders)
            walls.append((time.perf_counter() - t0) * 1e3)
        _, tm = native.prove_bytes(blinders, timings=True)
        print(json.dumps({"log_n": args.log_n, "lookup_share_of_used_rows": frac, "mode": {1: "host", 2: "device"}[mode],
                          "prove_ms": min(walls[1:]), "round2_lookup_ms": tm["round2_lookup_ms"],
                          "host_lookup_plumbing_ms": tm["host_lookup_plumbing_ms"], "round1_wires_ms": tm["round1_wires_ms"]}), flush=True)
    assert raws[1] == raws[2]
    native.close()
