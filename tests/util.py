"""Shared helpers for the parity tests (host <-> device movement, seeded inputs)."""
import numpy as np

from oracle import cref


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.cpu().numpy().view(np.uint64)


def rand_fr_mont(n, seed):
    return cref.to_mont(cref.FR, cref.rand_fe(cref.FR, n, seed))


def gen_xy():
    """BN254 G1 generator (1, 2) in Montgomery form, shape (8,)."""
    return cref.to_mont(cref.FQ, cref.ints_to_limbs([1, 2])).reshape(8)


def gpu_points(ctx, n, seed):
    """n pseudo-random G1 points k_i * G built on the device; returns (device tensor, host (n, 8) array)."""
    import torch
    k = cref.rand_fe(cref.FR, n, seed)
    out = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.g1_fixed_base_mul_dev(gen_xy(), to_dev(k), n, out)
    torch.cuda.synchronize()
    return out, to_host(out)


def skewed_scalars(n, seed):
    """Witness-like scalars: ~60% in {0, 1, small < 2^16}, rest uniform."""
    rng = np.random.default_rng(seed)
    s = cref.rand_fe(cref.FR, n, seed + 1)
    kind = rng.random(n)
    s[kind < 0.2] = 0
    ones = (kind >= 0.2) & (kind < 0.4)
    s[ones] = 0
    s[ones, 0] = 1
    small = (kind >= 0.4) & (kind < 0.6)
    s[small] = 0
    s[small, 0] = rng.integers(0, 1 << 16, size=int(small.sum()), dtype=np.uint64)
    return s
