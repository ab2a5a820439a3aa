"""A/B of the batched-affine pair rounds (csrc/msm_pairs.cuh, zkb_msm_set_mode) against the XYZZ-only accumulation: same MSM
results bit for bit (and the oracle at small sizes) and the device time of every phase, per number of rounds.

  python tools/check_msm_pairs.py [--sizes 16 18 20] [--modes 0 1 2 3 4]        # one JSON line per (size, tables, scalars, mode)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import uniform_scalars, witness_like_scalars

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[16, 18, 20])
ap.add_argument("--modes", type=int, nargs="+", default=[0, 1, 2, 3, 4])
ap.add_argument("--no-plain", action="store_true")
ap.add_argument("--oracle-below", type=int, default=15)
args = ap.parse_args()
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")


def timed(d):
    for _ in range(2):
        ctx.msm(d)
    ts, ph = [], []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ctx.msm(d); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); ph.append(ctx.msm_last_timing())
    k = int(np.argmin(ts))
    return r, ts[k], ph[k]


for ln in args.sizes:
    n = 1 << ln
    k = uniform_scalars(n, 7)
    k[1] = k[0]                                                   # a repeated point
    P = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(G, torch.from_numpy(k.view(np.int64)).cuda(), n, P)
    cases = {"uniform": uniform_scalars(n, 100), "witness_like": witness_like_scalars(n, 5)}
    for tables in ((True,) if args.no_plain else (False, True)):
        ctx.srs_load(P)
        if tables:
            ctx.srs_precompute(0)
        for name, sc in cases.items():
            d = torch.from_numpy(sc.view(np.int64)).cuda()
            ref = None
            for mode in args.modes:
                ctx.set_msm_mode(mode)
                (got, inf), t, ph = timed(d)
                if ref is None:
                    ref = (got.copy(), inf)
                row = {"log_n": ln, "tables": tables, "scalars": name, "mode": mode, "rounds": ph["pair_rounds"], "ms": round(t, 4),
                       "same_as_first_mode": bool(inf == ref[1] and np.array_equal(got, ref[0])),
                       "phases_ms": {p: round(ph[p], 4) for p in ("sort_ms", "pair_rounds_ms", "accumulate_ms", "heavy_ms", "reduce_ms", "total_ms")},
                       "c": ph["c"], "windows": ph["windows"]}
                if ln < args.oracle_below:
                    from oracle import cref
                    exp, einf = cref.msm_g1(P.cpu().numpy().view(np.uint64), sc)
                    row["bit_exact_vs_oracle"] = bool(einf == inf and np.array_equal(exp, got))
                print(json.dumps(row), flush=True)
            ctx.set_msm_mode(0)
        ctx.srs_precompute(-1)
