"""Size sweeps of BASELINE.json's configs 3 and 5 (device-timed with CUDA events, L2 flushed between repetitions).

  python tools/sweep.py ntt  [--sizes 16 18 20 22 24]            # Fr NTT / iNTT / coset variants, single and batch of 9
  python tools/sweep.py msm  [--sizes 16 18 20 22 24 26] [--cpu]  # G1 MSM on one GPU (fixed-base tables and plain bases)
  torchrun --nproc-per-node N tools/sweep.py msm ...              # the same totals sharded by point range over N GPUs

One JSON object per line on stdout (rank 0); `--out FILE` also writes them to FILE.  Roofline columns: NTT against the
measured HBM copy peak (64 B per element) and against the integer pipe; MSM bucket accumulation against
the integer pipe (n*W mixed additions x 10 products x 136 MACs x 2 IMAD slots)."""
import argparse, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkt_plonk_b200 as z
from bench import R_LIMBS, measured_peaks

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["ntt", "msm"])
ap.add_argument("--sizes", type=int, nargs="+", default=None)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--cpu", action="store_true", help="msm: also time the restated arkworks VariableBaseMSM on the host cores (<= 2^22)")
ap.add_argument("--out", default=None)
args = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = torch.device(f"cuda:{local}")
ctx = z.Context(local); ctx.set_stream(torch.cuda.current_stream())
FLUSH = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.int64, device=dev)
peaks, peak_src = measured_peaks()
INT_PEAK = ctx.bench_int(0)
lines = []


def emit(obj):
    if rank == 0:
        print(json.dumps(obj), flush=True)
        lines.append(obj)


def device_uniform_fr(n, seed):
    """n x 4 int64 limbs uniform in [0, r) (rejection sampling on the device), canonical little-endian limbs."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    MIN = -(1 << 63)
    r = [int(x) for x in R_LIMBS]
    signed = lambda u: u - (1 << 64) if u >= (1 << 63) else u
    rs = [torch.tensor(signed(x ^ (1 << 63)), dtype=torch.int64, device=dev) for x in r]   # r limbs with the top bit flipped

    def draw(m):
        a = torch.randint(MIN, (1 << 63) - 1, (m, 4), dtype=torch.int64, device=dev, generator=g)
        a[:, 3] &= 0x3FFFFFFFFFFFFFFF
        return a

    def ge_r(a):                                                # unsigned lexicographic a >= r, most significant limb first
        f = a ^ MIN
        ge = torch.ones(a.shape[0], dtype=torch.bool, device=dev)
        decided = torch.zeros_like(ge)
        for k in (3, 2, 1, 0):
            gt, lt = f[:, k] > rs[k], f[:, k] < rs[k]
            ge = torch.where(~decided & lt, torch.zeros_like(ge), ge)
            decided |= gt | lt
        return ge

    a = draw(n)
    while True:
        bad = ge_r(a).nonzero().flatten()
        if bad.numel() == 0:
            return a
        a[bad] = draw(bad.numel())


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        FLUSH.zero_()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.mean(ts), min(ts)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


def sweep_ntt():
    sizes = args.sizes or [16, 18, 20, 22, 24]
    for ln in sizes:
        n = 1 << ln
        d = device_uniform_fr(n, 5)
        ref = d.clone()
        ctx.ntt_dev(d, ln, False, True); ctx.ntt_dev(d, ln, True, True); torch.cuda.synchronize()
        rt = bool(torch.equal(d, ref))
        # algorithmic integer work: (N/2) log N butterflies' products (+ N for the coset scale, + N for 1/N), 272 IMAD slots each
        for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
            mean, best = timed(lambda: ctx.ntt_dev(d, ln, inv, cos), args.reps)
            muls = n / 2 * ln + (n if cos else 0) + (n if inv else 0)
            emit({"config": "ntt", "op": name, "log_n": ln, "ms": mean, "ms_min": best, "elems_per_s": n / (mean * 1e-3),
                  "hbm_GBps_algorithmic": 64.0 * n / (mean * 1e-3) / 1e9, "hbm_frac": 64.0 * n / (mean * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "imad_frac_algorithmic": muls * 272 / (mean * 1e-3) / INT_PEAK, "coset_round_trip_exact": rt})
        if ln <= 24:
            bat = [device_uniform_fr(n, 50 + k) for k in range(9)]
            mean, best = timed(lambda: ctx.ntt_batch_dev(bat, ln, False, True), args.reps)
            emit({"config": "ntt", "op": "coset_fft_batch9", "log_n": ln, "ms": mean, "ms_min": best, "elems_per_s": 9 * n / (mean * 1e-3),
                  "hbm_frac": 9 * 64.0 * n / (mean * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "imad_frac_algorithmic": 9 * (n / 2 * ln + n) * 272 / (mean * 1e-3) / INT_PEAK})
            del bat
        del d, ref


def sweep_msm():
    from zkt_plonk_b200.parallel import shard_bounds
    sizes = args.sizes or [16, 18, 20, 22, 24, 26]
    G = ctx.fp_binop(1, 5, np.array([[1, 0, 0, 0], [2, 0, 0, 0]], dtype=np.uint64)).reshape(8)
    gather = torch.zeros((world, 16), dtype=torch.int64, device=dev)
    for ln in sizes:
        n = 1 << ln
        b = shard_bounds(n, world); lo, hi = b[rank], b[rank + 1]
        k = device_uniform_fr(n, 7)[lo:hi].contiguous()         # the same global key on every world size
        P = torch.empty((hi - lo, 8), dtype=torch.int64, device=dev)
        ctx.g1_fixed_base_mul_dev(G, k, hi - lo, P)
        ctx.srs_load(P)
        del k
        s_full = device_uniform_fr(n, 100)
        s = s_full[lo:hi].contiguous()

        def step():
            if world == 1:
                return ctx.msm(s)
            part = ctx.msm_partial(s, 0, hi - lo)
            mine = torch.from_numpy(part.view(np.int64)).to(dev)
            dist.all_gather_into_tensor(gather, mine.reshape(1, 16))
            return z.sum_partials(gather.cpu().numpy().view(np.uint64))

        res = {}
        for mode in ("plain", "tables"):
            if mode == "tables":
                t0 = time.perf_counter(); ctx.srs_precompute(0); torch.cuda.synchronize(); t_build = time.perf_counter() - t0
            mean, best = timed(step, args.reps)
            res[mode] = step()
            tm = ctx.msm_last_timing()
            adds = (hi - lo) * tm["windows"]
            line = {"config": "msm", "log_n": ln, "n_gpus": world, "bases": mode, "ms": mean, "ms_min": best,
                    "points_per_s": n / (mean * 1e-3), "window_bits": tm["c"], "windows": tm["windows"],
                    "phases_ms_rank0": {k_: tm[k_] for k_ in ("sort_ms", "accumulate_ms", "heavy_ms", "reduce_ms", "total_ms")},
                    "accumulate_imad_frac": adds * (6 * 136 + 2 * 108 + 200) * 2 / (tm["accumulate_ms"] * 1e-3) / INT_PEAK,
                    "group_adds_c_independent": n * 254 / ln}
            if mode == "tables":
                line["table_build_s"] = t_build
                line["same_point_as_plain_bases"] = bool(np.array_equal(res["plain"][0], res["tables"][0]))
            if args.cpu and mode == "tables" and rank == 0 and world == 1 and ln <= 22:
                from oracle import cref
                threads = max(cref.num_threads(), len(os.sched_getaffinity(0)))
                Ph, sh = P.cpu().numpy().view(np.uint64), s.cpu().numpy().view(np.uint64)
                t0 = time.perf_counter(); exp, einf = cref.msm_g1(Ph, sh, threads); dt = time.perf_counter() - t0
                line["cpu_restated_arkworks"] = {"seconds": dt, "points_per_s": n / dt, "cores": threads,
                                                 "bit_exact_vs_gpu": bool(einf == res["tables"][1] and np.array_equal(exp, res["tables"][0]))}
            emit(line)
        ctx.srs_precompute(-1)
        del P, s, s_full


if args.what == "ntt":
    sweep_ntt()
else:
    sweep_msm()
if rank == 0 and args.out:
    with open(args.out, "w") as f:
        for obj in lines:
            f.write(json.dumps(obj) + "\n")
if dist:
    dist.barrier(); dist.destroy_process_group()
