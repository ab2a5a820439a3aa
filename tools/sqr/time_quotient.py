"""Device time of the fused quotient kernel at n = 2^log_n (4n coset elements) on random field data, and a checksum of
its output (two builds of the same sources must print the same checksum).   python tools/sqr/time_quotient.py [log_n]"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import uniform_scalars

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n4 = 4 << log_n
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
base = torch.from_numpy(uniform_scalars(n4, 3).view(np.int64)).cuda()          # values below r: valid field elements
tabs = [torch.roll(base, shifts=977 * (k + 1), dims=0).contiguous() for k in range(20)]
out = torch.empty_like(base)
ch = uniform_scalars(5, 9)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    ctx.quotient_evals_dev(log_n, ch, tabs[:9], tabs[9:], out)
ts = []
for _ in range(5):
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.quotient_evals_dev(log_n, ch, tabs[:9], tabs[9:], out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
print(json.dumps({"quotient_log_n": log_n, "ms_min": min(ts), "ms_mean": sum(ts) / len(ts), "out_sha256_16": digest,
                  "lib": os.environ.get("ZKB200_LIB", "default")}), flush=True)
