"""Kernel / copy timeline of ONE host-scalar MSM (zkb_msm_g1: ranges through shared buckets) -- CUPTI activity records via
torch.profiler, as tools/trace_prove.py does for a proof.  Writes gpurun_out/<tag>_trace_msm_2^<log_n>.jsonl (one line per
kernel / memcpy / memset: name, stream, start and duration in us) and a summary (busy time per stream, idle gaps > 5 us).

  python tools/trace_msm.py [log_n] [tag]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import uniform_scalars

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
tag = sys.argv[2] if len(sys.argv) > 2 else "trace"
n = 1 << log_n
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
k = torch.from_numpy(uniform_scalars(n, 7).view(np.int64)).cuda()
P = torch.empty((n, 8), dtype=torch.int64, device="cuda")
ctx.g1_fixed_base_mul_dev(ctx.g1_generator(), k, n, P)
ctx.srs_load(P); ctx.srs_precompute(0)
s = torch.from_numpy(uniform_scalars(n, 100).view(np.int64)).pin_memory()
sh = s.numpy().view(np.uint64)
for _ in range(4):
    t0 = time.perf_counter(); ref = ctx.msm(sh); plain_ms = (time.perf_counter() - t0) * 1e3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter(); got = ctx.msm(sh); traced_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
assert np.array_equal(got[0], ref[0])
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/{tag}_chrome_msm.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
recs = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
t_first = recs[0]["ts"]
short = lambda nm: nm.replace("(anonymous namespace)::", "").replace("zkb::", "").replace("void ", "").split("(")[0]
out = open(f"gpurun_out/{tag}_trace_msm_2^{log_n}.jsonl", "w")
busy = {}
for e in recs:
    st = e.get("args", {}).get("stream", -1)
    busy[st] = busy.get(st, 0.0) + e["dur"]
    out.write(json.dumps({"name": short(e["name"])[:50], "cat": e["cat"], "stream": st, "t_us": round(e["ts"] - t_first, 1), "dur_us": round(e["dur"], 1)}) + "\n")
iv = sorted((e["ts"], e["ts"] + e["dur"], short(e["name"])[-40:]) for e in recs)
gaps, union, lo0, hi0, nm0 = [], 0.0, iv[0][0], iv[0][1], iv[0][2]
for lo, hi, nm in iv[1:]:
    if lo > hi0:
        union += hi0 - lo0
        if lo - hi0 > 5.0: gaps.append({"at_us": round(hi0 - t_first, 1), "gap_us": round(lo - hi0, 1), "after": nm0, "before": nm})
        lo0, hi0, nm0 = lo, hi, nm
    elif hi > hi0:
        hi0, nm0 = hi, nm
union += hi0 - lo0
summary = {"summary": True, "log_n": log_n, "msm_ms_untraced": plain_ms, "msm_ms_traced": traced_ms, "records": len(recs),
           "span_ms": (max(h for _, h, _ in iv) - t_first) / 1e3, "gpu_busy_union_ms": union / 1e3,
           "busy_ms_per_stream": {str(a): b / 1e3 for a, b in busy.items()}, "idle_ms_in_gaps_over_5us": sum(g["gap_us"] for g in gaps) / 1e3,
           "largest_gaps": sorted(gaps, key=lambda g: -g["gap_us"])[:20]}
out.write(json.dumps(summary) + "\n"); out.close(); os.remove(path)
print(json.dumps(summary))
