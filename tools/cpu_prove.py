"""The restated CPU prover end to end (oracle backend: C restatement of arkworks' MSM / FFT and of plonk-core's widgets
under the Python round schedule), timed on the host cores -- the "withdraw prove on CPU" row of BASELINE.json's configs,
restated (the Rust binary cannot be built here).  TEST INFRASTRUCTURE: uses oracle/, never the CUDA library's kernels.

  python tools/cpu_prove.py --log-n 18 [--verify]      # n = 2^18 is the withdraw circuit's size
Prints one JSON line: setup / prove wall times, per-round split, core count."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cref, plonk_ref
from zkt_plonk_b200 import prover, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=14)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--verify", action="store_true")
args = ap.parse_args()
P = prover.P
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P
t0 = time.perf_counter()
circ = synthetic.make_circuit(args.log_n, seed=1)
t_circ = time.perf_counter() - t0
t0 = time.perf_counter()
srs = plonk_ref.make_srs_host(circ.n + 8, TAU)
t_srs = time.perf_counter() - t0
be = plonk_ref.OracleBackend(srs)
t0 = time.perf_counter()
pk, vk = prover.setup(be, circ)
t_setup = time.perf_counter() - t0
runs = []
for r in range(args.reps):
    tm = {}
    t0 = time.perf_counter()
    proof = prover.prove(be, pk, vk, circ, list(range(1000 + r, 1019 + r)), timings=tm)
    tm["total_ms"] = (time.perf_counter() - t0) * 1e3
    runs.append(tm)
best = min(runs, key=lambda t: t["total_ms"])
out = {"workload": f"plonk_plookup_prove_n=2^{args.log_n}, restated CPU prover (C oracle under the Python round schedule)",
       "cores": os.cpu_count(), "oracle_threads": cref.num_threads(), "prove_ms": best["total_ms"],
       "rounds_ms": {k: v for k, v in best.items() if k != "total_ms"}, "setup_s": t_setup, "srs_s": t_srs, "circuit_s": t_circ,
       "proof_bytes": len(proof.to_bytes())}
if args.verify:
    from zkt_plonk_b200 import verifier
    t0 = time.perf_counter()
    out["verify_rc"] = verifier.verify(vk, proof.to_bytes(), list(circ.pi.values()), plonk_ref.make_cvk(TAU))
    out["verify_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out), flush=True)
