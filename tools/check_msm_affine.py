"""EXPERIMENT: the batch-affine bucket accumulation (csrc/msm_affine.cu, zkb_msm_set_mode(ctx, 1)) against the
default XYZZ accumulation: same MSM results (bit for bit, including the oracle at small sizes and adversarial inputs:
repeated points, P and -P in one bucket, zero / one / r-1 scalars) and device time of both.  Not part of the test suite:
the kernel was written without a GPU at hand (round 1 ended on its budget); run this first.

  python tools/check_msm_affine.py [--sizes 12 16 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import uniform_scalars, witness_like_scalars
from oracle import cref

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[10, 14, 18, 20])
args = ap.parse_args()
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")
R_MINUS_1 = np.array([0x43e1f593f0000000, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029], dtype=np.uint64)


def timed(d):
    for _ in range(2):
        ctx.msm(d)
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ctx.msm(d); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return r, min(ts), ctx.msm_last_timing()["accumulate_ms"]


for ln in args.sizes:
    n = 1 << ln
    k = uniform_scalars(n, 7)
    k[1] = k[0]                                                   # a repeated point
    P = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(G, torch.from_numpy(k.view(np.int64)).cuda(), n, P)
    cases = {"uniform": uniform_scalars(n, 100), "witness_like": witness_like_scalars(n, 5)}
    adv = uniform_scalars(n, 9)
    adv[0] = adv[1] = 5                                           # the same point twice in one bucket -> doubling
    adv[2] = R_MINUS_1                                            # -1: lands next to +1 entries
    adv[3] = 0
    adv[3, 0] = 1
    adv[4:8] = adv[8:12]                                          # equal digits on different points
    cases["adversarial"] = adv
    for tables in (False, True):
        ctx.srs_load(P)
        if tables:
            ctx.srs_precompute(0)
        for name, sc in cases.items():
            d = torch.from_numpy(sc.view(np.int64)).cuda()
            ctx.set_msm_mode(0)
            (ref, ref_inf), t0, a0 = timed(d)
            ctx.set_msm_mode(1)
            (got, got_inf), t1, a1 = timed(d)
            ctx.set_msm_mode(0)
            ok = ref_inf == got_inf and np.array_equal(ref, got)
            row = {"log_n": ln, "tables": tables, "scalars": name, "bit_exact_vs_xyzz": bool(ok), "xyzz_ms": t0, "affine_ms": t1,
                   "xyzz_accumulate_ms": a0, "affine_accumulate_ms": a1}
            if ln <= 14:
                exp, einf = cref.msm_g1(P.cpu().numpy().view(np.uint64), sc)
                row["bit_exact_vs_oracle"] = bool(einf == got_inf and np.array_equal(exp, got))
            print(json.dumps(row), flush=True)
        ctx.srs_precompute(-1)
