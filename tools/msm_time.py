"""MSM timing sweep on one GPU: device-timed phases per size and scalar distribution (fixed-base tables or plain bases).

  python tools/msm_time.py --sizes 16 18 20 22 [--plain] [--c C] [--skewed] [--curve bls12_381]
Prints one JSON line per size: ms per MSM (CUDA events), phase split, window plan; also the integer / FP64 peaks."""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import uniform_scalars

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[16, 18, 20, 22])
ap.add_argument("--plain", action="store_true")
ap.add_argument("--c", type=int, default=0)
ap.add_argument("--skewed", action="store_true")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--peaks", action="store_true")
ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381", "bls12_377"])
args = ap.parse_args()
ctx = z.Context(0, curve=args.curve); ctx.set_stream(torch.cuda.current_stream())
if args.curve != "bn254":                                       # 252 random bits: below every curve's r
    def uniform_scalars(n, seed):
        a = np.random.default_rng(seed).integers(0, 2**64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
        return a
if args.peaks:
    print(json.dumps({"imad_per_s": ctx.bench_int(0), "imad_wide_per_s": ctx.bench_int(1), "fq_mul_per_s": ctx.bench_int(2),
                      "dfma_per_s": ctx.bench_int(3)}), flush=True)
G = ctx.g1_generator()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")
for ln in args.sizes:
    n = 1 << ln
    k = torch.from_numpy(uniform_scalars(n, 7).view(np.int64)).cuda()
    P = torch.empty((n, ctx.aff_words), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(G, k, n, P)
    ctx.srs_load(P)
    del k
    if not args.plain:
        ctx.srs_precompute(args.c)
    elif args.c:
        ctx.set_msm_window(args.c)
    sc = uniform_scalars(n, 100)
    if args.skewed:
        rng = np.random.default_rng(1)
        kind = rng.random(n)
        sc[kind < 0.2] = 0
        ones = (kind >= 0.2) & (kind < 0.4); sc[ones] = 0; sc[ones, 0] = 1
        small = (kind >= 0.4) & (kind < 0.6); sc[small] = 0
        sc[small, 0] = rng.integers(0, 1 << 16, size=int(small.sum()), dtype=np.uint64)
    d = torch.from_numpy(sc.view(np.int64)).cuda()
    for _ in range(3):
        ctx.msm(d)
    ts, ph = [], []
    for _ in range(args.reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.msm(d); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); ph.append(ctx.msm_last_timing())
    best = min(range(len(ts)), key=lambda i: ts[i])
    print(json.dumps({"log_n": ln, "ms": statistics.mean(ts), "ms_min": ts[best], "points_per_s": n / (statistics.mean(ts) * 1e-3),
                      "phases": ph[best], "tables": not args.plain, "skewed": args.skewed, "curve": args.curve}), flush=True)
    del P, d
