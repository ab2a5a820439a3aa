import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkt_plonk_b200 as z
ctx = z.Context(0); ctx.set_stream(torch.cuda.current_stream())
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log_n
d = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda"); d[:, 3] &= (1 << 60) - 1
for _ in range(3):
    ctx.ntt_dev(d, log_n, False, True)
torch.cuda.synchronize()
print("ok")
