"""GpuEvaluationDomain: the `D: EvaluationDomain<F> + EvaluationDomainExt<F>` seam of the reference
(plonk-core/src/plonk.rs:39-46, plonk-core/src/util.rs:27-59), backed by the sm_100a NTT kernels.

Same names, argument meaning and error behaviour as ark-poly 0.3 Radix2EvaluationDomain for BN254 Fr:
`new(k)` rounds up to a power of two and returns None above 2^28; `fft`/`coset_fft` zero-pad short input;
natural order in and out; data is Montgomery-form limbs ((n, 4) uint64 on the host, or a CUDA tensor with
4 x 8-byte words per element for HBM-resident use).
"""
import numpy as np

from . import field
from .context import Context


class GpuEvaluationDomain:
    def __init__(self, ctx, log_size):
        self.ctx = ctx
        self.log_size_of_group = log_size
        self._size = 1 << log_size
        self._group_gen = field.root_of_unity(log_size)

    # -- constructors / metadata (EvaluationDomain::new, size, EvaluationDomainExt)
    @classmethod
    def new(cls, num_coeffs, ctx=None):
        size = 1 if num_coeffs <= 1 else 1 << (num_coeffs - 1).bit_length()
        log_size = size.bit_length() - 1
        if log_size > field.TWO_ADICITY:
            return None                       # -> Error::InvalidEvalDomainSize at the call site (prove.rs:77-81)
        return cls(ctx if ctx is not None else Context(), log_size)

    def size(self):
        return self._size

    def log_size(self):
        return self.log_size_of_group

    def group_gen(self):
        return self._group_gen

    def element(self, i):
        return pow(self._group_gen, i, field.R_MOD)

    def elements(self):
        x = 1
        for _ in range(self._size):
            yield x
            x = x * self._group_gen % field.R_MOD

    def evaluate_vanishing_polynomial(self, tau):
        return (pow(tau, self._size, field.R_MOD) - 1) % field.R_MOD

    def vanishing_polynomial(self):
        """Sparse (degree, coeff) pairs of x^n - 1."""
        return [(0, field.R_MOD - 1), (self._size, 1)]

    # -- transforms
    def _run(self, x, inverse, coset, in_place):
        if isinstance(x, np.ndarray):
            if x.shape[0] > self._size:
                x = x[: self._size]           # resize() truncates longer input
            length = x.shape[0]
            if in_place and length == self._size and x.flags["C_CONTIGUOUS"]:
                buf = x
            else:
                buf = np.zeros((self._size, 4), dtype=np.uint64)
                buf[:length] = x
            self.ctx.ntt_host(buf, self.log_size_of_group, inverse, coset, length)
            return buf
        # CUDA tensor: must already be full size (the caller owns HBM-resident buffers)
        if x.numel() != 4 * self._size:
            raise ValueError("device tensors must hold exactly size() elements")
        t = x if in_place else x.clone()
        self.ctx.ntt_dev(t, self.log_size_of_group, inverse, coset)
        return t

    def fft(self, coeffs):
        return self._run(coeffs, False, False, False)

    def ifft(self, evals):
        return self._run(evals, True, False, False)

    def coset_fft(self, coeffs):
        return self._run(coeffs, False, True, False)

    def coset_ifft(self, evals):
        return self._run(evals, True, True, False)

    def fft_in_place(self, coeffs):
        return self._run(coeffs, False, False, True)

    def ifft_in_place(self, evals):
        return self._run(evals, True, False, True)

    def coset_fft_in_place(self, coeffs):
        return self._run(coeffs, False, True, True)

    def coset_ifft_in_place(self, evals):
        return self._run(evals, True, True, True)

    def __eq__(self, other):
        return isinstance(other, GpuEvaluationDomain) and other.log_size_of_group == self.log_size_of_group

    def __hash__(self):
        return hash(("GpuEvaluationDomain", self.log_size_of_group))
