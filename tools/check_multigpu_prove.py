"""torchrun script: SPMD full prove with point-range sharded commitments (csrc/comm.cu) on N GPUs.

Every rank runs zkb_plonk_setup / zkb_plonk_prove on the same synthetic circuit; the committer key is split into N
contiguous ranges (one per GPU, each with its own fixed-base tables) and the XYZZ partial sums of every batch of
commitments are all-gathered over NCCL.  Checks: all ranks emit identical proof bytes; rank 0 compares them with
an unsharded single-GPU proof of the same inputs (byte-identical) and runs the restated verifier.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/check_multigpu_prove.py --log-n 12
"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import zkt_plonk_b200 as z
from zkt_plonk_b200 import prover, synthetic
from zkt_plonk_b200.parallel import attach_sharded_srs

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=12)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-verify", action="store_true")
ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381", "bls12_377"])
args = ap.parse_args()
from zkt_plonk_b200 import field
field.use_curve(args.curve)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
P = prover.P
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P
n = 1 << args.log_n
circ = synthetic.make_circuit(args.log_n, seed=3, table_size=min(1024, n // 4))
ctx = z.Context(local, curve=args.curve); ctx.set_stream(torch.cuda.current_stream())
G = ctx.g1_generator()


def srs_range(c, lo, hi):
    """[tau^i] G for i in [lo, hi), built in HBM."""
    k = np.empty((hi - lo, 4), dtype=np.uint64)
    x = pow(TAU, lo, P)
    for i in range(hi - lo):
        for j in range(4):
            k[i, j] = (x >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        x = x * TAU % P
    out = torch.empty((hi - lo, c.aff_words), dtype=torch.int64, device=f"cuda:{local}")
    c.g1_fixed_base_mul_dev(G, torch.from_numpy(k.view(np.int64)).to(out.device), hi - lo, out)
    torch.cuda.synchronize()
    return out


from zkt_plonk_b200.parallel import attach_replicated_srs
blinders = list(range(500, 519))
layouts = {}
raw = None
for layout in ("point_range", "replicated_fanout", "replicated_shard", "replicated_auto"):
    c = ctx if layout == "point_range" else z.Context(local, curve=args.curve)
    c.set_stream(torch.cuda.current_stream())
    if layout == "point_range":
        lo, hi = attach_sharded_srs(c, lambda a, b: srs_range(c, a, b), n + 8)
    else:
        attach_replicated_srs(c, lambda a, b: srs_range(c, a, b), n + 8, fanout={"replicated_fanout": 1, "replicated_shard": 0, "replicated_auto": -1}[layout])
        lo, hi = 0, n + 8
    assert c.srs_size() == n + 8
    nat = prover.NativeProver(c, circ)
    r0 = nat.prove_bytes(blinders)
    times = []
    for _ in range(args.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        r1 = nat.prove_bytes(blinders)
        times.append((time.perf_counter() - t0) * 1e3)
        assert r1 == r0
    digest = np.frombuffer(hashlib.sha256(r0).digest(), dtype=np.uint8)
    alld = c.comm_allgather(digest)
    assert all(np.array_equal(alld[r], digest) for r in range(world)), f"{layout}: ranks disagree on the proof bytes"
    assert raw is None or raw == r0, f"{layout}: proof differs from the point-range layout's"
    raw = r0
    tmax = torch.tensor([min(times)], device="cuda"); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    layouts[layout] = float(tmax.item())
    if layout == "point_range":
        native, rounds_src = nat, nat
        _, rounds = nat.prove_bytes(blinders, timings=True)
    else:
        nat.close(); c.close()
if rank == 0:
    res = {"curve": args.curve, "world": world, "log_n": args.log_n, "range": [lo, hi], "prove_ms_by_layout": layouts, "ranks_agree": True, "layouts_byte_identical": True,
           "rounds_ms_rank0_with_syncs": {k: round(v, 3) for k, v in rounds.items()}}
    ref_ctx = z.Context(local, curve=args.curve); ref_ctx.set_stream(torch.cuda.current_stream())
    ref_ctx.srs_load(srs_range(ref_ctx, 0, n + 8)); ref_ctx.srs_precompute(0)
    ref = prover.NativeProver(ref_ctx, circ)
    raw_ref = ref.prove_bytes(blinders)
    t0 = time.perf_counter(); ref.prove_bytes(blinders); res["prove_ms_single_gpu"] = (time.perf_counter() - t0) * 1e3
    res["byte_identical_to_single_gpu"] = raw_ref == raw
    if not args.no_verify:
        from oracle import plonk_ref, pyref
        pyref.use_curve(args.curve)
        res["verifier_accepts"] = plonk_ref.verify(native.vk(), prover.proof_from_bytes(raw), list(circ.pi.values()), TAU) == 0
    print(json.dumps(res), flush=True)
    assert res["byte_identical_to_single_gpu"] and res.get("verifier_accepts", True)
    ref.close(); ref_ctx.close()
dist.barrier()
native.close(); ctx.close()
dist.destroy_process_group()
