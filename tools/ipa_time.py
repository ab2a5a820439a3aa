"""Times inner-product-argument openings (zkt_plonk_b200.ipa.GpuIPA over csrc/ipa.cu) on one GPU: one JSON line per size with the
whole opening, the first round's two halves and the check.  `ZKB_IPA_NAF=0 python tools/ipa_time.py` walks the challenge's plain
binary expansion in the key fold instead of its non-adjacent form (A/B).  --curve bn254 | bls12_381 | bls12_377."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bn254")
    ap.add_argument("--logs", type=int, nargs="*", default=[12, 16, 18, 20])
    args = ap.parse_args()
    import torch
    import zkt_plonk_b200 as z
    from zkt_plonk_b200 import field
    from zkt_plonk_b200.ipa import GpuIPA
    field.use_curve(args.curve)
    ctx = z.Context(0, curve=args.curve)
    ctx.set_stream(torch.cuda.current_stream())
    r = field.R_MOD
    rng = np.random.default_rng(1)
    top = max(args.logs)
    k = rng.integers(0, 1 << 62, size=((1 << top) + 1, 4), dtype=np.int64)
    k[:, 3] &= (1 << 58) - 1
    pts = torch.empty(((1 << top) + 1, ctx.aff_words), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), torch.from_numpy(k).to("cuda"), (1 << top) + 1, pts)
    torch.cuda.synchronize()
    for log_d in args.logs:
        n = 1 << log_d
        pc = GpuIPA(ctx)
        pc.load_committer_key(pts[:n].contiguous(), pc._pt_ints(pts[n].cpu().numpy().view(np.uint64).reshape(-1), False))
        coeffs = torch.from_numpy(k[:n].copy()).to("cuda")
        point = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % r
        C = pc.commit_dev(coeffs, n)
        pc.open(coeffs, n, C, point)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof, value = pc.open(coeffs, n, C, point)
        torch.cuda.synchronize()
        t_open = time.perf_counter() - t0
        c, zv, key = coeffs.clone(), coeffs.clone(), pts[:n].clone()
        x = np.array(field.int_to_limbs(field.to_mont(point)), dtype=np.uint64)
        xi = np.array(field.int_to_limbs(field.to_mont(pow(point, -1, r))), dtype=np.uint64)
        full = np.array(field.int_to_limbs(field.to_mont(r - 2)), dtype=np.uint64)          # a full-length challenge
        fulli = np.array(field.int_to_limbs(field.to_mont(pow(r - 2, -1, r))), dtype=np.uint64)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.ipa_round_lr_dev(c, zv, key, n)
        torch.cuda.synchronize()
        t_lr = time.perf_counter() - t0
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        ctx.ipa_round_fold_dev(c, zv, key, n, x, xi)
        e[1].record()
        key2 = pts[:n].clone()
        e[2].record()
        ctx.ipa_round_fold_dev(c, zv, key2, n, full, fulli)
        e[3].record()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ok = pc.check(C, point, value, proof) if log_d <= 18 else None
        t_check = time.perf_counter() - t0
        print(json.dumps({"curve": args.curve, "log_d": log_d, "naf": os.environ.get("ZKB_IPA_NAF", "1") != "0", "open_ms": t_open * 1e3,
                          "first_round_lr_ms": t_lr * 1e3, "first_round_fold_ms_192bit_challenge": e[0].elapsed_time(e[1]),
                          "first_round_fold_ms_full_challenge": e[2].elapsed_time(e[3]),
                          "fold_scalar_muls_per_s_full": (n / 2) / (e[2].elapsed_time(e[3]) * 1e-3), "check_ms": t_check * 1e3, "check": ok}),
              flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
