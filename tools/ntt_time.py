import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
FLUSH = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device="cuda")
def timeit(fn, reps=7, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        FLUSH.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
for log_n in (16, 18, 20, 22, 24, 26):
    n = 1 << log_n
    d = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda"); d[:, 3] &= (1 << 60) - 1
    for inv, cos, ln in ((False, False, n), (False, True, n), (True, True, n), (False, True, n // 4)):
        best, med = timeit(lambda: ctx.ntt_dev(d, log_n, inv, cos, length=ln))
        print(f"ntt 2^{log_n} inv={int(inv)} coset={int(cos)} len={'n' if ln==n else 'n/4'}: {best:.3f} ms best, {med:.3f} med, {n/best/1e6:.2f} Gelem/s", flush=True)
    del d
