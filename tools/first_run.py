"""Quick device timing sweep (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z

ctx = z.Context(0)
ctx.set_stream(torch.cuda.current_stream())

FLUSH = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        FLUSH.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

for mode, name in ((0, "IMAD"), (1, "IMAD.WIDE"), (2, "Fq montmul")):
    print(f"int peak {name}: {ctx.bench_int(mode)/1e12:.3f} T/s", flush=True)

def rand_words(n, words):
    t = torch.randint(0, 2**62, (n, words), dtype=torch.int64, device="cuda")
    return t

for log_n in ():
    n = 1 << log_n
    d = rand_words(n, 4)
    d[:, 3] &= (1 << 60) - 1   # < p
    for inv, cos in ((False, False), (False, True), (True, True)):
        best, med = timeit(lambda: ctx.ntt_dev(d, log_n, inv, cos))
        print(f"ntt 2^{log_n} inv={int(inv)} coset={int(cos)}: {best:.3f} ms best, {med:.3f} med, {n/best/1e6:.2f} Gelem/s, {64*n/best/1e6:.1f} GB/s alg", flush=True)
    del d

G = np.zeros(8, dtype=np.uint64)
# generator (1,2) in Montgomery form via the device field library
one_two = np.zeros((2, 4), dtype=np.uint64); one_two[0, 0] = 1; one_two[1, 0] = 2
G = ctx.fp_binop(1, 5, one_two).reshape(8)
for log_n in (18, 20, 22):
    n = 1 << log_n
    k = rand_words(n, 4); k[:, 3] &= (1 << 60) - 1
    P = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    t0 = time.time(); ctx.g1_fixed_base_mul_dev(G, k, n, P); torch.cuda.synchronize()
    print(f"gen 2^{log_n} points: {time.time()-t0:.3f}s", flush=True)
    ctx.srs_load(P)
    s = rand_words(n, 4); s[:, 3] &= (1 << 60) - 1
    for c in (0,):
        ctx.set_msm_window(c)
        best, med = timeit(lambda: ctx.msm(s), reps=5, warm=2)
        print(f"msm 2^{log_n} c={c}: {best:.3f} ms best, {med:.3f} med, {n/best/1e3:.1f} Mpts/s", flush=True)
    ctx.set_msm_window(0)
    for c in (0, 17, 19, 20, 21):
        t0 = time.time(); ctx.srs_precompute(c); torch.cuda.synchronize(); tp = time.time() - t0
        best, med = timeit(lambda: ctx.msm(s), reps=5, warm=2)
        tm = ctx.msm_last_timing()
        print(f"msm 2^{log_n} FIXED-BASE c={tm['c']} (build {tp:.3f}s): {best:.3f} ms best, {med:.3f} med, {n/best/1e3:.1f} Mpts/s  phases sort={tm['sort_ms']:.3f} acc={tm['accumulate_ms']:.3f} heavy={tm['heavy_ms']:.3f} red={tm['reduce_ms']:.3f}", flush=True)
    ctx.srs_precompute(-1)
    best, med = timeit(lambda: ctx.msm(s), reps=5, warm=2)
    tm = ctx.msm_last_timing()
    print(f"msm 2^{log_n} plain c={tm['c']}: {best:.3f} ms  phases sort={tm['sort_ms']:.3f} acc={tm['accumulate_ms']:.3f} heavy={tm['heavy_ms']:.3f} red={tm['reduce_ms']:.3f}", flush=True)
