"""CPU-only: the parts of bench.py's contract that can be checked without a GPU -- the reference arm prints exactly ONE JSON
line on stdout (library chatter goes to stderr) with the keys the driver reads, and the synthetic scalar generators stay in
range."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--log-n", "10"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "g1_msm_points_per_s" and d["unit"] == "points/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_synthetic_scalars_are_canonical():
    sys.path.insert(0, ROOT)
    import bench
    for gen in (bench.uniform_scalars, bench.witness_like_scalars):
        a = gen(1 << 12, 3)
        vals = [sum(int(a[i, k]) << (64 * k) for k in range(4)) for i in range(a.shape[0])]
        assert a.dtype == np.uint64 and a.shape == (1 << 12, 4) and max(vals) < R
    w = [sum(int(x) << (64 * k) for k, x in enumerate(row)) for row in bench.witness_like_scalars(1 << 12, 3)]
    assert 0.1 < sum(v == 0 for v in w) / len(w) < 0.3 and 0.1 < sum(v == 1 for v in w) / len(w) < 0.3


def test_closed_form_check_agrees_with_the_oracle_msm():
    """bench.py checks every MSM against (sum s_i k_i mod r) * G for bases P_i = k_i * G.  Pin that closed form (and the
    limb <-> integer helpers it rests on) against the oracle's VariableBaseMSM on the same points."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import cref
    n = 300
    k = bench.uniform_scalars(n, 7)
    s = bench.uniform_scalars(n, 8)
    G = cref.to_mont(cref.FQ, cref.ints_to_limbs([1, 2])).reshape(8)
    P = cref.g1_mul(G, k)
    got, inf = cref.msm_g1(P, s)
    d = bench.dot_mod_r(bench.limbs_to_ints(s), bench.limbs_to_ints(k))
    assert d == sum(a * b for a, b in zip(bench.limbs_to_ints(s), bench.limbs_to_ints(k))) % R
    assert bench.dot_mod_r_fast(s, k) == d and bench.dot_mod_r_fast(s, k, chunk=64) == d          # GEMM on 13-bit limbs
    big_s, big_k = bench.uniform_scalars(50000, 9), bench.uniform_scalars(50000, 10)
    big_s[:3] = np.uint64(0xFFFFFFFFFFFFFFFF)                                                      # all-ones limbs (not canonical: fine here)
    assert bench.dot_mod_r_fast(big_s, big_k) == bench.dot_mod_r(bench.limbs_to_ints(big_s), bench.limbs_to_ints(big_k))
    exp = cref.g1_mul(G, bench.int_to_limbs(d))
    assert not inf and np.array_equal(exp.reshape(8), got)
    assert bench.bench_config(20, 2) == {"workload": "kzg_commit_g1_msm_2^20", "points_per_gpu": 1 << 20, "total_points": 2 << 20,
                                         "scalars": "uniform in [0,r), canonical", "curve": "BN254"}
