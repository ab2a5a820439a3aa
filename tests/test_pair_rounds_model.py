"""CPU-only model of the bookkeeping of csrc/msm_pairs.cuh (batched-affine pair rounds): the host plan's upper bounds and
static pool offsets, pack -> exclusive scan -> finish, the per-thread walk over "m consecutive pairs of the round" (binary
search of the first bucket, runs inside a bucket, empty buckets skipped) and the carried-over odd entries.  Points are
integers and the group law is integer addition, so the test isolates the indexing: after R rounds the entries left in every
bucket must add up to the bucket's original sum, for uniform, skewed and degenerate bucket loads."""
import random

import pytest

POOL, SIGN, ID_MASK = 0x40000000, 0x80000000, 0x3FFFFFFF
THREADS, M_MIN, M_MAX = 128, 32, 128


def make_plan(rounds, entries, nb, resident):
    e_ub, p_ub, base, ms, grids = [entries], [], [0], [], []
    for r in range(rounds):
        p_ub.append(e_ub[r] // 2)
        nonempty = min(e_ub[r], nb)
        nxt = min((e_ub[r] + nonempty + 1) // 2, e_ub[r])
        e_ub.append(nxt)
        base.append(base[r] + p_ub[r])
        m = min(max((p_ub[r] + resident - 1) // resident, M_MIN), M_MAX)
        ms.append(m)
        grids.append(max(1, (p_ub[r] + m * THREADS - 1) // (m * THREADS)))
    return e_ub, p_ub, base, ms, grids


def value(points, pool, ref):
    v = (pool if ref & POOL else points)[ref & ID_MASK]
    return -v if ref & SIGN else v


def run_rounds(points, refs, counts, starts, rounds, resident):
    nb = len(counts)
    entries = sum(counts)
    e_ub, p_ub, base, ms, grids = make_plan(rounds, entries, nb, resident)
    pool = [None] * (base[rounds] + 1)
    for r in range(rounds):
        assert sum(counts) <= e_ub[r]
        # pack + exclusive scan (nb + 1 elements)
        sc, px, py = [], 0, 0
        for b in range(nb + 1):
            c = counts[b] if b < nb else 0
            sc.append((px, py))
            px += c >> 1
            py += (c + 1) >> 1
        total = sc[nb][0]
        assert total <= p_ub[r] and sc[nb][1] <= e_ub[r + 1]
        nrefs = [None] * (sc[nb][1] + 1)
        # finish
        ncounts = [(c + 1) >> 1 for c in counts]
        nstarts = [sc[b][1] for b in range(nb)]
        for b in range(nb):
            if counts[b] & 1:
                nrefs[nstarts[b] + (counts[b] >> 1)] = refs[starts[b] + counts[b] - 1]
        # pair_add: every thread of the grid
        m, T = ms[r], grids[r] * THREADS
        assert m * T >= total
        for t in range(T):
            j0 = t * m
            if j0 >= total:
                continue
            cnt = min(m, total - j0)
            lo, hi = 0, nb
            while hi - lo > 1:
                mid = (lo + hi) >> 1
                if sc[mid][0] <= j0:
                    lo = mid
                else:
                    hi = mid
            b = lo
            cur, nxt = sc[b], sc[b + 1]
            k, pb, sb = j0 - cur[0], nxt[0] - cur[0], starts[b]
            assert pb > 0
            i, mine = 0, []
            while i < cnt:
                if k == pb:
                    b += 1
                    cur, nxt = nxt, sc[b + 1]
                    pb, k = nxt[0] - cur[0], 0
                    if pb:
                        sb = starts[b]
                    continue
                ln = min(pb - k, cnt - i)
                for u in range(ln):
                    mine.append((refs[sb + 2 * (k + u)], refs[sb + 2 * (k + u) + 1]))
                    assert nrefs[cur[1] + k + u] is None
                    nrefs[cur[1] + k + u] = POOL | (base[r] + j0 + i + u)
                i += ln
                k += ln
            for i, (ra, rb) in enumerate(mine):
                assert pool[base[r] + j0 + i] is None
                pool[base[r] + j0 + i] = value(points, pool, ra) + value(points, pool, rb)
        refs, counts, starts = nrefs, ncounts, nstarts
        assert all(refs[starts[b] + k] is not None for b in range(nb) for k in range(counts[b]))
    return refs, counts, starts, pool


def loads(kind, nb, rnd):
    if kind == "uniform":
        return [rnd.randrange(10, 40) for _ in range(nb)]
    if kind == "sparse":
        return [rnd.choice([0, 0, 0, 1, 2, 3]) for _ in range(nb)]
    if kind == "skewed":
        c = [rnd.choice([0, 0, 1, 5]) for _ in range(nb)]
        c[0], c[nb // 2], c[nb - 1] = 4000, 777, 1001
        return c
    if kind == "ones":
        return [1] * nb
    if kind == "single":
        return [0] * (nb - 1) + [2049]
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["uniform", "sparse", "skewed", "ones", "single"])
@pytest.mark.parametrize("rounds", [1, 3, 6])
def test_pair_round_bookkeeping(kind, rounds):
    rnd = random.Random(hash((kind, rounds)) & 0xFFFF)
    nb = 300
    counts = loads(kind, nb, rnd)
    entries = sum(counts)
    points = [rnd.randrange(1, 1 << 40) for _ in range(entries + 5)]
    starts, refs, want, pos = [], [], [], 0
    for b in range(nb):
        starts.append(pos)
        s = 0
        for _ in range(counts[b]):
            idx = rnd.randrange(len(points))
            neg = rnd.random() < 0.5
            refs.append(idx | (SIGN if neg else 0))
            s += -points[idx] if neg else points[idx]
        want.append(s)
        pos += counts[b]
    for resident in (256, 4096):
        r2, c2, s2, pool = run_rounds(points, list(refs), list(counts), list(starts), rounds, resident)
        for b in range(nb):
            exp_c = counts[b]
            for _ in range(rounds):
                exp_c = (exp_c + 1) >> 1
            assert c2[b] == exp_c
            assert sum(value(points, pool, r2[s2[b] + k]) for k in range(c2[b])) == want[b]
