"""Generates tests/golden/vectors.npz from oracle/pyref.py (Python big-int definitions, seeded).

Run from the repo root:  python tests/golden/make_golden.py
The reference's own implementation (Rust + crates.io arkworks 0.3) cannot run in this image, so these are
definition-level golden vectors: O(n^2) DFT sums, per-point double-and-add, and the public alt_bn128 known
answers 2G, 3G (EIP-196 test vectors).  Both the C oracle and the CUDA path are tested against them.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as P  # noqa: E402


def limbs(vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            a[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return a


def main():
    rnd = random.Random(20261018)
    out = {}
    # --- NTT family, canonical integers in and out
    for log_n in (0, 1, 2, 3, 5, 7):
        n = 1 << log_n
        x = [rnd.randrange(P.R_MOD) for _ in range(n)]
        out[f"ntt_in_{log_n}"] = limbs(x)
        out[f"ntt_fwd_{log_n}"] = limbs(P.dft_naive(x, log_n))
        out[f"ntt_inv_{log_n}"] = limbs(P.dft_naive(x, log_n, inverse=True))
        out[f"ntt_cfwd_{log_n}"] = limbs(P.coset_ntt(x, log_n))
        out[f"ntt_cinv_{log_n}"] = limbs(P.coset_intt(x, log_n))
    # short input, zero padded by fft_in_place's resize
    x = [rnd.randrange(P.R_MOD) for _ in range(5)]
    out["ntt_short_in"] = limbs(x)
    out["ntt_short_fwd_4"] = limbs(P.dft_naive(x, 4))
    out["ntt_short_cfwd_4"] = limbs(P.coset_ntt(x, 4))
    # --- G1 known answers
    g2, g3 = P.g1_mul(2, P.G1_GEN), P.g1_mul(3, P.G1_GEN)
    assert g2 == (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
                  0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)
    assert g3 == (0x0769bf9ac56bea3ff40232bcb1b6bd159315d84715b8e679f2d355961915abf0,
                  0x2ab799bee0489429554fdb7c8d086475319e63b40b9c5b57cdf1ff3dd9fe2261)
    out["g1_2g"] = limbs(list(g2))
    out["g1_3g"] = limbs(list(g3))
    # --- MSM: 24 points k_i * G, mixed scalars incl. 0, 1, r-1, duplicates, an infinity
    ks = [rnd.randrange(1, P.R_MOD) for _ in range(24)]
    pts = [P.g1_mul(k, P.G1_GEN) for k in ks]
    pts[5] = pts[4]
    pts[7] = None
    sc = [rnd.randrange(P.R_MOD) for _ in range(24)]
    sc[0], sc[1], sc[2], sc[3] = 0, 1, P.R_MOD - 1, 2
    sc[5] = sc[4]
    sc[9] = rnd.randrange(1 << 16)
    res = P.msm_naive(pts, sc)
    out["msm_points"] = limbs([c for p in pts for c in (p if p else (0, 0))])
    out["msm_scalars"] = limbs(sc)
    out["msm_result"] = limbs(list(res))
    # cancellation: s*P + (r-s)*P = infinity
    out["msm_cancel_points"] = limbs([c for p in (pts[0], pts[0]) for c in p])
    out["msm_cancel_scalars"] = limbs([12345, P.R_MOD - 12345])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.npz"), **out)
    print("wrote vectors.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
