"""CPU, world_size 2 over gloo: the host logic of the point-range sharded MSM (shard bounds, all-gather of the
XYZZ partials, final combine through the C ABI's host-side zkb_g1_sum_partials).  The per-shard partial sums come
from the oracle here; on GPUs they come from the CUDA bucket method (tests/test_gpu_msm.py covers that leg)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from oracle import cref
from zkt_plonk_b200.parallel import affine_to_xyzz, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zkt_plonk_b200.parallel import ShardedMSM
    G = cref.to_mont(cref.FQ, cref.ints_to_limbs([1, 2])).reshape(8)
    P = cref.g1_mul(G, cref.rand_fe(cref.FR, n, 1))
    s = cref.rand_fe(cref.FR, n, 2)
    s[0] = 0
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    one = cref.to_mont(cref.FQ, cref.ints_to_limbs([1]))[0]

    def partial(sc):
        xy, inf = cref.msm_g1(P[lo:hi], sc)
        return affine_to_xyzz(xy, inf, one)

    m = ShardedMSM(ctx=None, partial_fn=partial)
    got, inf = m.msm(np.ascontiguousarray(s[lo:hi]))
    exp, einf = cref.msm_g1(P, s)
    ok = bool(inf == einf and np.array_equal(got, exp))
    # cancelling shards: rank 0 holds +X, rank 1 holds -X  -> identity on every rank
    xy, _ = cref.msm_g1(P[:4], s[:4])
    neg = xy.copy()
    neg[4:] = cref.binop(cref.FQ, 2, np.zeros((1, 4), dtype=np.uint64), xy[4:].reshape(1, 4))[0]
    m2 = ShardedMSM(ctx=None, partial_fn=lambda _: affine_to_xyzz(xy if rank == 0 else neg, False, one))
    got2, inf2 = m2.msm(None)
    ok = ok and inf2 and not got2.any()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(8, 8) == list(range(9))
    assert shard_bounds(3, 4) == [0, 1, 2, 3, 3]
    b = shard_bounds((1 << 20) + 3, 8)
    assert b[0] == 0 and b[-1] == (1 << 20) + 3 and all(0 <= y - x - (1 << 17) <= 1 for x, y in zip(b, b[1:]))


def test_replicated_key_shares_partition_every_commitment():
    """zkb_srs_set_replicated (csrc/msm.cu replicated_share, host arithmetic): for every world size, batch size and layout the
    shares of the ranks tile [offset, offset + len) exactly once per commitment; with fan-out every rank serves exactly one
    commitment of the batch and the groups differ in size by at most one rank."""
    import ctypes
    from zkt_plonk_b200 import _lib
    fn = _lib.lib().zkb_test_replicated_share
    out = (ctypes.c_size_t * 2)()
    for world in (1, 2, 3, 4, 8):
        for E in (1, 2, 3, 10):
            for fanout in (0, 1, -1):
                for offset, length in ((0, (1 << 20) + 3), (5, 1000), (0, 3), (7, 0)):
                    served = [0] * world
                    for k in range(E):
                        pieces = []
                        for r in range(world):
                            assert fn(world, r, fanout, E, k, offset, length, out) == 0
                            if out[1] > out[0]:
                                pieces.append((out[0], out[1]))
                                served[r] += 1
                        pieces.sort()
                        pos = offset
                        for lo, hi in pieces:
                            assert lo == pos
                            pos = hi
                        assert pos == offset + length
                        if fanout == 1 and 2 <= E <= world and length >= world:
                            assert world // E <= len(pieces) <= world // E + 1
                    if fanout == 1 and 2 <= E <= world and length >= world:
                        assert max(served) == 1
    assert fn(2, 2, 0, 1, 0, 0, 10, out) != 0                        # rank out of range


def test_sharded_msm_world2_gloo():
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, 2, port, 301, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
