"""EXPERIMENT: XYZZ mixed additions against batch-affine additions with one shared inversion per CTA and step
(csrc/probe_batch_affine.cu), on the same points of an L2-resident table.  Prints additions per second of both loops and the
number of accumulators whose results differ (must be 0).      python tools/probe_batch_affine.py [--log-table 12] [--steps 128]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import uniform_scalars

ap = argparse.ArgumentParser()
ap.add_argument("--log-table", type=int, default=12)
ap.add_argument("--steps", type=int, default=128)
args = ap.parse_args()
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
n = 1 << args.log_table
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
k = torch.from_numpy(uniform_scalars(n, 11).view(np.int64)).cuda()
table = torch.empty((n, 8), dtype=torch.int64, device="cuda")
ctx.g1_fixed_base_mul_dev(G, k, n, table)
torch.cuda.synchronize()
for m in (4, 8, -4, -8, -16):        # negative: accumulators in global memory (L2) instead of shared memory
    out = (ctypes.c_double * 2)()
    mis = ctypes.c_uint(0)
    ctx._check(ctx._lib.zkb_probe_batch_affine(ctx._h, ctypes.c_void_p(table.data_ptr()), args.log_table, m, args.steps, out, ctypes.byref(mis)))
    print(json.dumps({"accumulators_per_thread": m, "steps": args.steps, "xyzz_adds_per_s": out[0], "batch_affine_adds_per_s": out[1],
                      "ratio": out[1] / out[0], "mismatches": mis.value}), flush=True)
