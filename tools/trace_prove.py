"""Kernel timeline of ONE proof through the C++ driver (CUPTI activity records via torch.profiler; there is no nsys here).

Writes gpurun_out/<tag>_trace_2^<log_n>.jsonl: one line per kernel / memcpy / memset (name, stream, start and duration in us
relative to the first record) and a summary line (busy time per stream, union of busy intervals, idle gaps > 5 us with the
kernels on either side).  The profiler slows the launches down, so the wall time of the traced proof is reported beside the
untraced one: read the trace for ORDER and GAPS, not for the absolute proof time.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from zkt_plonk_b200 import prover, synthetic

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
tag = sys.argv[2] if len(sys.argv) > 2 else "trace"
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
for r in range(4):
    t0 = time.perf_counter()
    raw = native.prove_bytes(bl)
    plain_ms = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter()
    raw2 = native.prove_bytes(bl)
    traced_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
assert raw2 == raw
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/{tag}_chrome_2^{log_n}.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
recs = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
recs.sort(key=lambda e: e["ts"])
t_first = recs[0]["ts"]
out = open(f"gpurun_out/{tag}_trace_2^{log_n}.jsonl", "w")
def short(nm):
    return nm.replace("(anonymous namespace)::", "").replace("zkb::", "").replace("void ", "").split("(")[0]
busy = {}
for e in recs:
    name = short(e["name"])
    st = e.get("args", {}).get("stream", -1)
    busy[st] = busy.get(st, 0.0) + e["dur"]
    out.write(json.dumps({"name": name[:60], "cat": e["cat"], "stream": st, "t_us": round(e["ts"] - t_first, 2), "dur_us": round(e["dur"], 2)}) + "\n")
# union of busy intervals and the idle gaps between them
iv = sorted((e["ts"], e["ts"] + e["dur"], short(e["name"])[-40:]) for e in recs)
gaps, union, cur_lo, cur_hi, cur_name = [], 0.0, iv[0][0], iv[0][1], iv[0][2]
for lo, hi, nm in iv[1:]:
    if lo > cur_hi:
        union += cur_hi - cur_lo
        if lo - cur_hi > 5.0: gaps.append({"at_us": round(cur_hi - t_first, 1), "gap_us": round(lo - cur_hi, 1), "after": cur_name, "before": nm})
        cur_lo, cur_hi, cur_name = lo, hi, nm
    elif hi > cur_hi:
        cur_hi, cur_name = hi, nm
union += cur_hi - cur_lo
summary = {"summary": True, "log_n": log_n, "prove_ms_untraced": plain_ms, "prove_ms_traced": traced_ms, "records": len(recs),
           "span_ms": (max(h for _, h, _ in iv) - t_first) / 1e3, "gpu_busy_union_ms": union / 1e3,
           "busy_ms_per_stream": {str(k): v / 1e3 for k, v in busy.items()}, "idle_gaps_over_5us": len(gaps),
           "idle_ms_in_those_gaps": sum(g["gap_us"] for g in gaps) / 1e3, "largest_gaps": sorted(gaps, key=lambda g: -g["gap_us"])[:40]}
out.write(json.dumps(summary) + "\n")
out.close()
os.remove(path)
print(json.dumps({k: v for k, v in summary.items() if k != "largest_gaps"}))
