"""GPU, slow: byte identity of whole proofs at the north-star sizes.  The native round driver (zkb_plonk_setup /
zkb_plonk_prove, fixed-base SRS tables, every kernel at production size: batched NTTs of 2^18 / 2^20 and 2^20 / 2^22 points,
MSMs of n + 3 points, the fused quotient kernel over 4n) against the restated CPU prover (oracle backend: C restatement of
arkworks' VariableBaseMSM / radix-2 FFT and of plonk-core's widgets under the same round schedule) on the same circuit,
witness, SRS and blinders; then the library's pairing verifier (zkb_plonk_verify, Proof::verify of proof.rs:285-503) must
accept the bytes.  n = 2^18 is the withdraw circuit's size (BASELINE.json configs[0]), n = 2^20 configs[3].
The CPU side takes ~1 min (2^18) and a few minutes (2^20) on the box's host cores."""
import numpy as np
import pytest

from oracle import cref, plonk_ref
from tests.util import gen_xy, to_dev, to_host
from zkt_plonk_b200 import prover, synthetic, verifier

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
P = prover.P
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % P


def _srs(ctx, n_points):
    import torch
    pw = np.empty(n_points, dtype=object)
    x = 1
    for i in range(n_points):
        pw[i] = x
        x = x * TAU % P
    k = np.empty((n_points, 4), dtype=np.uint64)
    for j in range(4):
        k[:, j] = ((pw >> (64 * j)) & ((1 << 64) - 1)).astype(np.uint64)
    out = torch.empty((n_points, 8), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(gen_xy(), to_dev(k), n_points, out)
    torch.cuda.synchronize()
    return out, to_host(out)


@pytest.mark.parametrize("log_n", [18, 20])
def test_native_proof_equals_the_cpu_prover_bytes(ctx, log_n):
    circ = synthetic.make_circuit(log_n, seed=1)
    d_srs, h_srs = _srs(ctx, circ.n + 8)
    ctx.srs_load(d_srs)
    ctx.srs_precompute(0)
    blinders = list(range(1000, 1019))
    native = prover.NativeProver(ctx, circ)
    raw = native.prove_bytes(blinders)
    vk = native.vk()
    native.close()
    ctx.srs_precompute(-1)
    del d_srs
    pub = list(circ.pi.values())
    assert verifier.verify(vk, raw, pub, verifier.make_cvk(TAU)) == 0            # pairing check, host C++
    obe = plonk_ref.OracleBackend(h_srs)
    opk, ovk = prover.setup(obe, circ)
    assert vk.commits == ovk.commits and vk.pi_roots == ovk.pi_roots
    oraw = prover.prove(obe, opk, ovk, circ, blinders).to_bytes()
    assert raw == oraw, "GPU proof differs from the CPU (oracle-backend) proof"
    tampered = bytearray(raw)
    tampered[11 * 32 + 2 * 33 + 5] ^= 1                                           # one bit of the first evaluation
    assert verifier.verify(vk, bytes(tampered), pub, verifier.make_cvk(TAU)) != 0
