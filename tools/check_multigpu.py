"""torchrun script: point-range sharded MSM over NCCL checked against the oracle (run with --nproc-per-node N)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import zkt_plonk_b200 as z
from zkt_plonk_b200.parallel import ShardedMSM, shard_bounds
from oracle import cref

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ctx = z.Context(local); ctx.set_stream(torch.cuda.current_stream())
n = (1 << 16) + 5
G = cref.to_mont(cref.FQ, cref.ints_to_limbs([1, 2])).reshape(8)
k = cref.rand_fe(cref.FR, n, 1)
s = cref.rand_fe(cref.FR, n, 2)
b = shard_bounds(n, world); lo, hi = b[rank], b[rank + 1]
dk = torch.from_numpy(k[lo:hi].view(np.int64)).cuda()
dP = torch.empty((hi - lo, 8), dtype=torch.int64, device="cuda")
ctx.g1_fixed_base_mul_dev(G, dk, hi - lo, dP)
m = ShardedMSM(ctx)
m.load_srs_range(dP)
got, inf = m.msm(torch.from_numpy(s[lo:hi].view(np.int64)).cuda())
if rank == 0:
    P = cref.g1_mul(G, k)
    exp, einf = cref.msm_g1(P, s)
    ok = inf == einf and np.array_equal(got, exp)
    print(f"sharded MSM over {world} GPUs, n={n}: {'BIT-EXACT' if ok else 'MISMATCH'}", flush=True)
    assert ok
dist.barrier(); dist.destroy_process_group()
