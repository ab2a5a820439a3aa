"""NTT timing on one GPU as JSON lines: coset forward / inverse, zero-padded coset (the quotient round's shape) and the batch of
nine, per size.  ZKB_NTT_KERNEL=0/1/2 in the environment selects the pass kernel (zkb_ntt_set_kernel).

  python tools/ntt_ab.py [--sizes 16 18 20 22 24]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[16, 18, 20, 22, 24])
ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381", "bls12_377"])
args = ap.parse_args()
ctx = z.Context(0, curve=args.curve); ctx.set_stream(torch.cuda.current_stream())
FLUSH = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")
peak = ctx.bench_int(0)


def timeit(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        FLUSH.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


variant = {"0": "default", "1": "generic", "2": "radix4_all"}[os.environ.get("ZKB_NTT_KERNEL", "0")]
for log_n in args.sizes:
    n = 1 << log_n
    d = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda"); d[:, 3] &= (1 << 60) - 1
    prods = (n // 2) * log_n + n
    for name, inv, cos, ln in (("coset_fwd", False, True, n), ("coset_inv", True, True, n), ("plain_inv", True, False, n),
                               ("coset_fwd_zero_padded_n/4+3", False, True, n // 4 + 3)):
        best, med = timeit(lambda: ctx.ntt_dev(d, log_n, inv, cos, length=ln))
        print(json.dumps({"curve": args.curve, "kernel": variant, "log_n": log_n, "what": name, "ms_best": round(best, 4), "ms_median": round(med, 4),
                          "gelem_per_s": round(n / best / 1e6, 3), "imad_frac_algorithmic": round(2 * 136 * prods / (best * 1e-3) / peak, 4)}), flush=True)
    if log_n <= 22:
        ds = [d] + [torch.roll(d, k + 1, 0).contiguous() for k in range(8)]
        best, med = timeit(lambda: ctx.ntt_batch_dev(ds, log_n, False, True, length=n // 4 + 3), reps=5)
        print(json.dumps({"kernel": variant, "log_n": log_n, "what": "batch9_coset_fwd_zero_padded", "ms_best": round(best, 4), "ms_median": round(med, 4),
                          "gelem_per_s": round(9 * n / best / 1e6, 3)}), flush=True)
        del ds
    del d
