"""Multi-GPU MSM: point-range sharding with an NCCL all-gather of the 128-byte XYZZ partial sums.

One process per GPU (`torch.distributed`, backend "nccl"; "gloo" for the CPU tests of the host logic).  The SRS is
split into `world` contiguous ranges that stay resident on their GPU; a commitment hands every rank the matching
slice of the scalar vector, each GPU runs the full bucket method on its range, and the only exchange is
world x 128 B.  NCCL has no elliptic-curve reduction, so ranks gather the partials and add them (<= 7 additions,
host side, identical on every rank).  The reference has no analogue (single process, rayon only: SURVEY.md 2.3).
"""
import numpy as np

from .context import sum_partials


def shard_bounds(n, world):
    """Contiguous, balanced ranges: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(n, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def affine_to_xyzz(xy, is_inf, one_mont):
    """(8,) affine Montgomery point -> (16,) XYZZ with ZZ = ZZZ = 1 (or all-zero for the identity)."""
    out = np.zeros(16, dtype=np.uint64)
    if not is_inf:
        out[:8] = xy
        out[8:12] = one_mont
        out[12:16] = one_mont
    return out


def attach_sharded_srs(ctx, load_range, global_n, group=None, precompute=True):
    """SPMD set-up of point-range sharded commitments behind the C ABI (csrc/comm.cu): attaches an NCCL communicator
    to `ctx`, loads this rank's contiguous range of the committer key and declares it (zkb_srs_set_range).
    `load_range(lo, hi)` returns powers_of_g[lo:hi] (host array or CUDA tensor).  Afterwards zkb_commit_batch_dev,
    zkb_plonk_setup and zkb_plonk_prove run unchanged on every rank and return identical results."""
    rank, world = ctx.comm_init(group)
    b = shard_bounds(global_n, world)
    ctx.srs_load(load_range(b[rank], b[rank + 1]))
    ctx.srs_set_range(b[rank], global_n)
    if precompute:
        ctx.srs_precompute(0)
    return b[rank], b[rank + 1]


def attach_replicated_srs(ctx, load_range, global_n, group=None, fanout=-1, precompute=True):
    """The other SPMD layout (csrc/comm.cu zkb_srs_set_replicated): every rank loads the WHOLE committer key
    (`load_range(0, global_n)`); per batch of commitments the library either cuts every commitment over all ranks or gives
    each commitment its own group of ranks (fan-out of the 3 / 3 / 2 / 3 independent commitments of a round).  fanout: -1
    cost model, 0 always shard, 1 always fan out.  zkb_commit_batch_dev / zkb_plonk_setup / zkb_plonk_prove run unchanged
    on every rank and return identical results."""
    rank, world = ctx.comm_init(group)
    ctx.srs_load(load_range(0, global_n))
    ctx.srs_set_replicated(fanout)
    if precompute:
        ctx.srs_precompute(0)
    return rank, world


class ShardedMSM:
    """`partial_fn(scalars_shard) -> (16,) uint64 XYZZ` computes this rank's partial sum; by default it is the
    CUDA bucket method on the context's resident SRS range."""

    def __init__(self, ctx=None, group=None, partial_fn=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.ctx = ctx
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.partial_fn = partial_fn if partial_fn is not None else self._cuda_partial

    def _cuda_partial(self, scalars_dev):
        return self.ctx.msm_partial(scalars_dev, 0, scalars_dev.numel() // 4)

    def load_srs_range(self, points):
        """points: this rank's contiguous range of powers_of_g (device tensor or host array)."""
        self.ctx.srs_load(points)

    def msm(self, scalars_shard):
        """All ranks call this with their own slice; every rank returns the same ((8,) affine, is_inf)."""
        import torch
        part = np.ascontiguousarray(self.partial_fn(scalars_shard), dtype=np.uint64).reshape(16)
        if self.world == 1:
            return sum_partials(part.reshape(1, 16))
        backend = self.dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(part.view(np.int64)).to(dev)
        gathered = torch.empty((self.world, 16), dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(gathered, mine.reshape(1, 16), group=self.group)
        return sum_partials(gathered.cpu().numpy().view(np.uint64))
