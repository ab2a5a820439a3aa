"""GpuIPA: the reference's second `PC` -- `IPA<G, D>` = ark-poly-commit 0.3 `ipa_pc::InnerProductArgPC` over G1 with Blake2s
(plonk-core/src/commitment.rs:49-86; every test of plonk-core/src/test.rs runs on it, :73,84) -- over the sm_100a kernels.

`commit` is the same MSM as KZG10's over `ck.comm_key` (held in HBM by the caller); `open` keeps the three vectors of the
folding argument (coefficients, powers of the point, committer key) in HBM and runs each round as zkb_ipa_round_lr_dev (two
MSMs + two inner products) and zkb_ipa_round_fold_dev (csrc/ipa.cu); `check` is succinct_check + the final-key MSM.  One
polynomial, no hiding, no degree bounds: how the prover calls its PC.  The transcript is a host-side hash between the rounds
(`oracle=`: bytes -> challenge); the default follows the dependency's compute_random_oracle_challenge with encodings recalled
from ark-ec / ark-ff 0.3 (unpinned: the dependency is not vendored in the reference; the Rust seam rust/zkb200/src/ipa.rs
keeps arkworks' own hash and calls only the two round entry points).

Nothing here touches oracle/: host-side group arithmetic (a point times a scalar, sums of a few points) goes through the
library's MSM over host bases.
"""
import hashlib

import numpy as np

from . import field

PROTOCOL_NAME = b"PC-DL-2020"


class IpaProof:
    def __init__(self, l_vec, r_vec, final_comm_key, c):
        self.l_vec, self.r_vec, self.final_comm_key, self.c = l_vec, r_vec, final_comm_key, c


def fr_bytes(v):
    return int(v % field.R_MOD).to_bytes(32, "little")


def g1_bytes(pt):
    """ark-ec 0.3 `impl ToBytes for GroupAffine`: x, y canonical little endian, then the infinity flag as one byte."""
    nb = field.FQ_BYTES
    if pt is None:
        return (0).to_bytes(nb, "little") + (1).to_bytes(nb, "little") + b"\x01"
    return int(pt[0]).to_bytes(nb, "little") + int(pt[1]).to_bytes(nb, "little") + b"\x00"


def random_oracle_challenge(data):
    """Blake2s(data || i as u64), i = 0, 1, ..: the first digest that Fr::from_random_bytes accepts."""
    bits = field.R_MOD.bit_length()
    i = 0
    while True:
        v = int.from_bytes(hashlib.blake2s(data + i.to_bytes(8, "little")).digest(), "little") & ((1 << bits) - 1)
        if v < field.R_MOD:
            return v
        i += 1


def _fr_limbs(v, mont=True):
    v %= field.R_MOD
    return np.array(field.int_to_limbs(field.to_mont(v) if mont else v), dtype=np.uint64)


class GpuIPA:
    def __init__(self, ctx):
        self.ctx = ctx
        self.key = None          # (n, aff_words) CUDA tensor: ck.comm_key, n a power of two
        self.h = None            # ck.h as (x, y) canonical ints
        self.n = 0

    # -- host-side group helpers over the library's MSM (no oracle, no Python curve arithmetic)
    def _pt_array(self, pts):
        w = field.FQ_WORDS
        out = np.zeros((len(pts), 2 * w), dtype=np.uint64)
        for i, pt in enumerate(pts):
            if pt is not None:
                for j, v in enumerate(pt):
                    out[i, w * j: w * j + w] = field.int_to_limbs(field.to_mont(v, field.Q_MOD), w)
        return out

    def _pt_ints(self, xy, inf):
        if inf:
            return None
        w = field.FQ_WORDS
        return (field.from_mont(field.limbs_to_int(xy[:w]), field.Q_MOD), field.from_mont(field.limbs_to_int(xy[w:2 * w]), field.Q_MOD))

    def lincomb(self, pts, scalars):
        """sum s_i P_i of a few host points (canonical ints in, canonical ints out)."""
        sc = np.array([field.int_to_limbs(s % field.R_MOD) for s in scalars], dtype=np.uint64)
        return self._pt_ints(*self.ctx.msm_bases(self._pt_array(pts), sc))

    # -- the PC interface
    def load_committer_key(self, comm_key_dev, h):
        """comm_key_dev: CUDA tensor of n affine points (n a power of two = supported degree + 1); h: ck.h as canonical ints."""
        n = comm_key_dev.numel() // self.ctx.aff_words
        if n < 2 or n & (n - 1):
            raise ValueError("ipa_pc committer keys hold a power of two of generators")
        self.key, self.h, self.n = comm_key_dev, h, n

    def commit_dev(self, coeffs_dev, length):
        """cm_commit(ck.comm_key, coeffs, None, None) for Montgomery coefficients in HBM; canonical-int point or None."""
        if length > self.n:
            raise ValueError(f"TooManyCoefficients: {length} > {self.n}")
        return self._pt_ints(*self.ctx.msm_points_dev(self.key, coeffs_dev, length))

    def open(self, coeffs_dev, length, commitment, point, oracle=random_oracle_challenge):
        """InnerProductArgPC::open.  coeffs_dev: Montgomery coefficients in HBM (left untouched).  Returns (IpaProof, value)."""
        import torch
        ctx, n, r = self.ctx, self.n, field.R_MOD
        dev = self.key.device
        c = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        c[:length] = coeffs_dev.view(-1, 4)[:length]
        # z = 1, point, point^2, ..: built in HBM by doubling (z[m .. 2m) = point^m * z[0 .. m)), log n small launches
        z = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        z[0] = torch.from_numpy(_fr_limbs(1).view(np.int64)).to(dev)
        filled = 1
        while filled < n:
            ctx.poly_lincomb_dev([z[:filled]], [filled], _fr_limbs(pow(point, filled, r)).reshape(1, 4), z[filled: 2 * filled], filled)
            filled *= 2
        value = field.from_mont(field.limbs_to_int(ctx.poly_eval_dev(c, n, _fr_limbs(point))))
        key = self.key.clone()
        x = oracle(g1_bytes(commitment) + fr_bytes(point) + fr_bytes(value))
        h_prime = self._pt_array([self.lincomb([self.h], [x])])[0]
        l_vec, r_vec = [], []
        while n > 1:
            (l_xy, l_inf), (r_xy, r_inf), _, _ = ctx.ipa_round_lr_dev(c, z, key, n, h_prime)
            L, R = self._pt_ints(l_xy, l_inf), self._pt_ints(r_xy, r_inf)
            l_vec.append(L)
            r_vec.append(R)
            x = oracle(fr_bytes(x) + g1_bytes(L) + g1_bytes(R))
            ctx.ipa_round_fold_dev(c, z, key, n, _fr_limbs(x), _fr_limbs(pow(x, -1, r)))
            n //= 2
        torch.cuda.synchronize()
        final_key = self._pt_ints(key[:1].cpu().numpy().view(np.uint64).reshape(-1), False)
        c0 = field.from_mont(field.limbs_to_int(c[0].cpu().numpy().view(np.uint64)))
        return IpaProof(l_vec, r_vec, final_key, c0), value

    def open_many(self, polys_dev, lengths, commitments, point, opening_challenge, oracle=random_oracle_challenge):
        """PC::open(ck, polys, commitments, point, opening_challenge, ..) as the prover calls it (prove.rs:381-451): one opening
        of sum_j xi^(2 j) p_j (ipa_pc draws two opening challenges per polynomial; the odd ones belong to degree-bounded
        polynomials' shifts, of which plonk-core has none).  Returns (IpaProof, [p_j(point)])."""
        import torch
        r = field.R_MOD
        k = len(polys_dev)
        n = max(lengths)
        weights = [pow(opening_challenge, 2 * j, r) for j in range(k)]
        combined = torch.empty((n, 4), dtype=torch.int64, device=self.key.device)
        w_arr = np.array([field.int_to_limbs(field.to_mont(w)) for w in weights], dtype=np.uint64).reshape(k, 4)
        self.ctx.poly_lincomb_dev(list(polys_dev), list(lengths), w_arr, combined, n)
        C = self.lincomb(list(commitments), weights)
        pts = np.tile(_fr_limbs(point), (k, 1))
        values = [field.from_mont(field.limbs_to_int(v)) for v in self.ctx.poly_eval_many_dev(list(polys_dev), list(lengths), pts)]
        proof, value = self.open(combined, n, C, point, oracle)
        assert value == sum(w * v for w, v in zip(weights, values)) % r
        return proof, values

    def check_many(self, commitments, point, values, proof, opening_challenge, oracle=random_oracle_challenge):
        """PC::check for the same batch: the commitments and values combined with xi^(2 j), then `check`."""
        r = field.R_MOD
        weights = [pow(opening_challenge, 2 * j, r) for j in range(len(commitments))]
        C = self.lincomb(list(commitments), weights)
        return self.check(C, point, sum(w * v for w, v in zip(weights, values)) % r, proof, oracle)

    def check(self, commitment, point, value, proof, oracle=random_oracle_challenge):
        """succinct_check (log n group operations) and the final-key check (zkb_ipa_final_key_dev: the check polynomial's
        coefficients expanded in HBM and one MSM of n over ck.comm_key)."""
        r = field.R_MOD
        k = len(proof.l_vec)
        if 1 << k != self.n or len(proof.r_vec) != k:
            return False
        x = oracle(g1_bytes(commitment) + fr_bytes(point) + fr_bytes(value))
        h_prime = self.lincomb([self.h], [x])
        pts, sc, challenges = [commitment, h_prime], [1, value], []
        for L, R in zip(proof.l_vec, proof.r_vec):
            x = oracle(fr_bytes(x) + g1_bytes(L) + g1_bytes(R))
            challenges.append(x)
            pts += [L, R]
            sc += [pow(x, -1, r), x]
        h_at_point = 1
        for i, ch in enumerate(challenges):
            h_at_point = h_at_point * (1 + ch * pow(point, 1 << (k - 1 - i), r)) % r
        if self.lincomb(pts, sc) != self.lincomb([proof.final_comm_key, h_prime], [proof.c, proof.c * h_at_point % r]):
            return False
        ch_arr = np.array([field.int_to_limbs(field.to_mont(v)) for v in challenges], dtype=np.uint64).reshape(-1, 4)
        return self._pt_ints(*self.ctx.ipa_final_key_dev(self.key, self.n, ch_arr)) == proof.final_comm_key
