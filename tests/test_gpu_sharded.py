"""GPU: point-range sharded commitments behind the C ABI (csrc/comm.cu).

One GPU is enough for the range logic: two contexts each hold one half of the committer key (zkb_srs_set_range) and
commit to the overlap of every polynomial with their range; the two partial commitments must add up to the
commitment of the unsharded key.  With >= 2 GPUs the SPMD prover runs under torchrun (NCCL all-gather of the
partials) and must emit the single-GPU proof byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cref
from tests.util import gpu_points, rand_fr_mont, to_dev
from zkt_plonk_b200.context import sum_partials
from zkt_plonk_b200.parallel import affine_to_xyzz, shard_bounds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world,tables", [(1000, 2, False), (4099, 3, True)])
def test_range_commitments_add_up(ctx, n, world, tables):
    import torch
    import zkt_plonk_b200 as z
    d_pts, h_pts = gpu_points(ctx, n, 11)
    polys = [rand_fr_mont(n, 21), rand_fr_mont(n - 37, 22), rand_fr_mont(5, 23)]
    polys[1][:100] = 0                                           # leading zeros: the slice starts inside rank 0's range
    lens = [p.shape[0] for p in polys]
    devs = [to_dev(p) for p in polys]
    ctx.srs_load(d_pts)
    want = ctx.commit_batch_dev(devs, lens)
    one = cref.to_mont(cref.FQ, cref.ints_to_limbs([1]))[0]
    b = shard_bounds(n, world)
    parts = []
    for r in range(world):
        c = z.Context(0)
        c.set_stream(torch.cuda.current_stream())
        c.srs_load(d_pts[b[r]:b[r + 1]].contiguous())
        c.srs_set_range(b[r], n)
        assert c.srs_size() == n
        if tables:
            c.srs_precompute(0)
        parts.append(c.commit_batch_dev(devs, lens))
        # a single commitment takes the same route
        xy, inf = c.commit_dev(devs[0], 0, lens[0])
        assert inf == parts[-1][0][1] and np.array_equal(xy, parts[-1][0][0])
        c.close()
    for k in range(len(polys)):
        xyzz = np.stack([affine_to_xyzz(parts[r][k][0], parts[r][k][1], one) for r in range(world)])
        got, inf = sum_partials(xyzz)
        assert inf == want[k][1] and np.array_equal(got, want[k][0])
        exp, einf = cref.msm_g1(h_pts[:lens[k]], cref.from_mont(cref.FR, polys[k]))
        assert inf == einf and np.array_equal(got, exp)


@pytest.mark.parametrize("world,fanout", [(2, 1), (3, 1), (8, 1), (8, 0), (5, -1)])
def test_replicated_key_shares_add_up(ctx, world, fanout):
    """zkb_srs_set_replicated: every "rank" holds the whole key and computes its share of each commitment of a batch --
    one group of ranks per commitment (fan-out) or every commitment cut over all ranks.  Run on one GPU with contexts that
    are told their rank (test hook), closing each batch without the exchange: the ranks' partial sums must add up to the
    commitments of the plain single-context batch, for batches of 3, 2 and 1 commitments with unequal lengths."""
    import torch
    import zkt_plonk_b200 as z
    n = 3001
    d_pts, h_pts = gpu_points(ctx, n, 14)
    polys = [rand_fr_mont(n, 31), rand_fr_mont(n - 37, 32), rand_fr_mont(n - 2, 33)]
    polys[1][:50] = 0
    devs = [to_dev(p) for p in polys]
    lens = [p.shape[0] for p in polys]
    ctx.srs_load(d_pts)
    want = ctx.commit_batch_dev(devs, lens)
    parts = {}
    for r in range(world):
        c = z.Context(0)
        c.set_stream(torch.cuda.current_stream())
        c.srs_load(d_pts)
        c._check(c._lib.zkb_test_set_rank_world(c._h, r, world))
        c.srs_set_replicated(fanout)
        c.srs_precompute(0)
        for batch in ((0, 1, 2), (0, 1), (2,)):
            c.commit_expect(len(batch))
            for k in batch:
                c.commit_push(devs[k], lens[k], 0)
            parts[(r, batch)] = c.commit_finish_partials(len(batch))
        c.close()
    for batch in ((0, 1, 2), (0, 1), (2,)):
        busy = [sum(bool(parts[(r, batch)][j].any()) for j in range(len(batch))) for r in range(world)]
        if fanout == 1 and world >= len(batch) > 1:
            assert max(busy) == 1                                   # fan-out: every rank works on ONE commitment of the batch
        for j, k in enumerate(batch):
            got, inf = sum_partials(np.stack([parts[(r, batch)][j] for r in range(world)]))
            assert inf == want[k][1] and np.array_equal(got, want[k][0]), (batch, k)


def test_range_must_fit_the_key(ctx):
    from zkt_plonk_b200._lib import ZkbError
    d_pts, _ = gpu_points(ctx, 64, 12)
    ctx.srs_load(d_pts)
    with pytest.raises(ZkbError):
        ctx.srs_set_range(10, 64)
    ctx.srs_set_range(0, 64)


def test_spmd_prove_two_gpus_byte_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run by tools/check_multigpu_prove.py under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_multigpu_prove.py"), "--log-n", "10"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"byte_identical_to_single_gpu": true' in out.stdout
