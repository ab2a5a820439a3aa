#!/usr/bin/env python
"""Control-flow model of csrc/msm_affine.cu (the experimental batch-affine bucket accumulation) in Python integers: lock-step
slots, the SET / ADD / DBL / CANCEL decisions, per-thread prefixes, the product tree over the thread totals, ONE inversion per
step, the walk back and the back-substitution -- checked against plain affine sums on task lists that contain repeated
points, a point and its negative, and tasks of different lengths.      python tools/sqr/model_batch_affine.py"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref

Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
NONE, ADD, DBL, SET, CANCEL = range(5)


def run_cta(tasks, threads, m):
    """tasks: list of lists of points (threads * m of them, possibly empty); returns the per-task sums (None = identity)."""
    assert len(tasks) == threads * m
    acc = [None] * len(tasks)                       # affine accumulators; None = empty
    slot = lambda i, t: i * threads + t             # task handled by thread t, slot i (same layout as the kernel)
    steps = max(len(x) for x in tasks)
    inversions = 0
    for k in range(steps):
        actions = [[NONE] * m for _ in range(threads)]
        pre = [[1] * m for _ in range(threads)]
        tree = [1] * (2 * threads)
        for t in range(threads):
            run = 1
            for i in range(m):
                T = slot(i, t)
                if k < len(tasks[T]):
                    p = tasks[T][k]
                    if acc[T] is None:
                        actions[t][i] = SET
                    else:
                        d = (p[0] - acc[T][0]) % Q
                        act = ADD
                        if d == 0:
                            if p[1] == acc[T][1]:
                                act, d = DBL, 2 * acc[T][1] % Q
                            else:
                                act = CANCEL
                        if act != CANCEL:
                            run = run * d % Q
                        actions[t][i] = act
                pre[t][i] = run
            tree[threads + t] = run
        w = threads // 2
        while w >= 1:
            for t in range(w):
                tree[w + t] = tree[2 * (w + t)] * tree[2 * (w + t) + 1] % Q
            w //= 2
        assert tree[1] != 0
        tree[1] = pow(tree[1], -1, Q)
        inversions += 1
        w = 1
        while w < threads:
            for t in range(w):
                inv_parent, left, right = tree[w + t], tree[2 * (w + t)], tree[2 * (w + t) + 1]
                tree[2 * (w + t)], tree[2 * (w + t) + 1] = inv_parent * right % Q, inv_parent * left % Q
            w *= 2
        for t in range(threads):
            inv_run = tree[threads + t]
            for i in range(m - 1, -1, -1):
                act, T = actions[t][i], slot(i, t)
                if act == NONE:
                    continue
                if act == CANCEL:
                    acc[T] = None
                    continue
                p = tasks[T][k]
                if act == SET:
                    acc[T] = p
                    continue
                x1, y1 = acc[T]
                d = (p[0] - x1) % Q if act == ADD else 2 * y1 % Q
                inv_d = inv_run * pre[t][i - 1] % Q if i else inv_run
                if i:
                    inv_run = inv_run * d % Q
                assert inv_d * d % Q == 1
                num = (p[1] - y1) % Q if act == ADD else 3 * x1 * x1 % Q
                lam = num * inv_d % Q
                x3 = (lam * lam - x1 - p[0]) % Q
                acc[T] = (x3, (lam * (x1 - x3) - y1) % Q)
    return acc, inversions


def main():
    rnd = random.Random(4)
    base = [pyref.g1_mul(rnd.randrange(1, 1 << 64), pyref.G1_GEN) for _ in range(40)]
    for threads, m in ((4, 2), (8, 3), (2, 8)):
        for trial in range(6):
            tasks = []
            for _ in range(threads * m):
                n = rnd.choice([0, 1, 2, 5, 9])
                pts = [rnd.choice(base) for _ in range(n)]
                if n >= 2 and rnd.random() < 0.5:
                    pts[1] = pts[0]                                  # doubling in the second step
                if n >= 5 and rnd.random() < 0.5:
                    pts[3] = pyref.g1_neg(pts[2])                    # ... + P - P ...
                    pts[1], pts[0] = pyref.g1_neg(pts[0]), pts[0]    # and a cancellation right after the first point
                if rnd.random() < 0.3:
                    pts = [pyref.g1_neg(p) for p in pts]             # negative digits
                tasks.append(pts)
            got, inversions = run_cta(tasks, threads, m)
            for T, pts in enumerate(tasks):
                exp = None
                for p in pts:
                    exp = pyref.g1_add(exp, p)
                assert got[T] == exp, (threads, m, trial, T)
            assert inversions == max(len(x) for x in tasks)
    print("batch-affine control-flow model ok")


if __name__ == "__main__":
    main()
