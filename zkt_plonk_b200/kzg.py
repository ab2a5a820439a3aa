"""GpuKZG10: the `PC: HomomorphicCommitment<F>` seam (plonk-core/src/commitment.rs:10-46) for
`KZG10<Bn254> = SonicKZG10<Bn254, DensePolynomial<Fr>>`, backed by the sm_100a Pippenger MSM.

`commit` follows ark-poly-commit 0.3 kzg10::commit with hiding_bound = None (what prove.rs passes):
skip the low-degree zero coefficients, convert the rest with into_repr, and take the inner product with
powers_of_g[skipped..]; trailing zero coefficients are dropped the way DensePolynomial::from_coefficients_vec
does.  The committer key's powers stay resident in HBM.
"""
import numpy as np

from .context import Context


class PCError(Exception):
    """Mirror of ark_poly_commit::Error for the cases this path can raise."""


class GpuKZG10:
    def __init__(self, ctx=None):
        self.ctx = ctx if ctx is not None else Context()

    def load_committer_key(self, powers_of_g):
        """powers_of_g: (n, 8) uint64 host array or CUDA tensor of affine G1 points (Montgomery x||y)."""
        self.ctx.srs_load(powers_of_g)

    def load_committer_key_file(self, path, max_points=0):
        """deserialize_from_file::<CommitterKey>(ck_path) (bin/src/parser.rs:5-14): ark-serialize unchecked bytes."""
        self.ctx.srs_load_ck_file(path, max_points)

    def supported_degree(self):
        return self.ctx.srs_size() - 1

    def _bounds(self, nz_mask):
        nz = np.flatnonzero(nz_mask)
        if nz.size == 0:
            return 0, 0
        return int(nz[0]), int(nz[-1]) + 1

    def commit_one(self, coeffs_mont):
        """coeffs_mont: host (len, 4) uint64 Montgomery coefficients.  Returns ((8,) uint64 affine, is_inf)."""
        lo, hi = self._bounds(coeffs_mont.any(axis=1))
        if hi > self.ctx.srs_size():
            raise PCError(f"TooManyCoefficients: {hi} > {self.ctx.srs_size()}")
        if hi == lo:
            return np.zeros(self.ctx.aff_words, dtype=np.uint64), True
        import torch
        d = torch.from_numpy(np.ascontiguousarray(coeffs_mont[lo:hi]).view(np.int64)).to(f"cuda:{self.ctx.device}")
        return self.ctx.commit_dev(d, lo, hi - lo)

    def commit(self, polys):
        """PolynomialCommitment::commit(ck, polys, None): one commitment per polynomial, in order."""
        return [self.commit_one(p) for p in polys]

    def commit_dev(self, coeffs_dev, n, skip=0):
        """HBM-resident polynomial of n coefficients (the caller has already located leading zeros)."""
        if skip + n > self.ctx.srs_size():
            raise PCError(f"TooManyCoefficients: {skip + n} > {self.ctx.srs_size()}")
        return self.ctx.commit_dev(coeffs_dev, skip, n)

    def commit_many_dev(self, coeffs_list, lens):
        """PolynomialCommitment::commit for a batch of HBM-resident polynomials (pipelined MSMs)."""
        for n in lens:
            if n > self.ctx.srs_size():
                raise PCError(f"TooManyCoefficients: {n} > {self.ctx.srs_size()}")
        return self.ctx.commit_batch_dev(coeffs_list, lens)

    def multi_scalar_mul(self, commitments, scalars_canonical):
        """HomomorphicCommitment::multi_scalar_mul (commitment.rs:31-46): sum_i scalars[i] * commitments[i]."""
        pts = np.ascontiguousarray(commitments, dtype=np.uint64).reshape(-1, self.ctx.aff_words)
        sc = np.ascontiguousarray(scalars_canonical, dtype=np.uint64).reshape(-1, 4)
        return self.ctx.msm_bases(pts, sc)
