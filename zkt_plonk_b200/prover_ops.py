"""Host-side mirrors of the reference's free functions on the prover hot path, same names and argument order,
running on the sm_100a kernels with every polynomial resident in HBM (CUDA tensors, 4 x int64 words per Fr,
Montgomery form).  The reference versions take `&[F]` / `&DensePolynomial<F>` and return `DensePolynomial<F>`:

  compute_z1_poly      plonk-core/src/permutation/mod.rs:181-257
  compute_z2_poly      plonk-core/src/lookup/mod.rs:25-85
  quotient_compute     plonk-core/src/proof_system/quotient_poly.rs:20-227  (quotient_poly::compute)
  extend_prover_key    plonk-core/src/proof_system/keys/mod.rs:78-146
  add_blinders_to_poly plonk-core/src/proof_system/prove.rs:472-483
  kzg_open             ark-poly-commit 0.3 SonicKZG10::open -> kzg10::open, prove.rs:381-451
"""
import numpy as np

from . import field


def _fr(x):
    """Python int (canonical) or (4,) uint64 Montgomery limbs -> (4,) uint64 Montgomery limbs."""
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint64).reshape(4)
    return np.array(field.int_to_limbs(field.to_mont(x)), dtype=np.uint64)


def _empty_like_fr(n, like):
    import torch
    return torch.empty((n, 4), dtype=torch.int64, device=like.device)


def compute_z1_poly(domain, beta, gamma, a, b, c, sigma1, sigma2, sigma3):
    """Returns the n coefficients of z1 (device tensor).  Raises like the reference's unwrap() on a zero denominator."""
    n = domain.size()
    for t in (a, b, c, sigma1, sigma2, sigma3):
        assert t.numel() == 4 * n                                   # assert_eq!(a.len(), n) ... (mod.rs:197-202)
    z = _empty_like_fr(n, a)
    domain.ctx.z1_evals_dev(domain.log_size(), _fr(beta), _fr(gamma), a, b, c, sigma1, sigma2, sigma3, z)
    if domain.ctx.grand_product_failed():
        raise ZeroDivisionError("compute_z1_poly: zero denominator (reference: dominator.inverse().unwrap())")
    return domain.ifft_in_place(z)                                  # poly_from_evals


def compute_z2_poly(domain, delta, epsilon, f, t, h1, h2):
    n = domain.size()
    for x in (f, t, h1, h2):
        assert x.numel() == 4 * n                                   # lookup/mod.rs:40-43
    z = _empty_like_fr(n, f)
    domain.ctx.z2_evals_dev(domain.log_size(), _fr(delta), _fr(epsilon), f, t, h1, h2, z)
    if domain.ctx.grand_product_failed():
        raise ZeroDivisionError("compute_z2_poly: zero denominator (reference: dominator.inverse().unwrap())")
    return domain.ifft_in_place(z)


EPK_ORDER = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_lookup", "q_table", "sigma1", "sigma2", "sigma3", "l1")
WIT_ORDER = ("z1", "z2", "a", "b", "c", "pi", "t", "h1", "h2")


def _coset_4n(domain_4n, poly, length):
    """coset_evals_from_poly_ref: zero-extend `length` coefficients to 4n and run the coset FFT in HBM."""
    import torch
    n4 = domain_4n.size()
    buf = torch.zeros((n4, 4), dtype=torch.int64, device=poly.device)
    buf[:length] = poly.reshape(-1, 4)[:length]
    domain_4n.ctx.ntt_dev(buf, domain_4n.log_size(), False, True, length=length)
    return buf


def extend_prover_key(domain, polys):
    """polys: dict name -> (device tensor of coefficients) for the 10 selector/sigma polynomials.
    Returns the 11 static 4n coset tables the quotient kernel streams (x_coset and zh_coset are not materialised)."""
    from .domain import GpuEvaluationDomain
    d4 = GpuEvaluationDomain.new(4 * domain.size(), domain.ctx)
    if d4 is None:
        raise ValueError("InvalidEvalDomainSize")
    epk = {}
    for name in EPK_ORDER[:-1]:
        p = polys[name]
        epk[name] = _coset_4n(d4, p, p.numel() // 4)
    epk["l1"] = domain.ctx.l1_coset_dev(domain.log_size(), _empty_like_fr(d4.size(), polys["q_m"]))
    return epk


def quotient_compute(domain, epk, alpha, beta, gamma, delta, epsilon, z1_poly, z2_poly, a_poly, b_poly, c_poly, pi_poly,
                     h1_poly, h2_poly, t_poly):
    """quotient_poly::compute: 9 coset FFTs on 4n, the fused quotient kernel, one coset iFFT.  Polynomials are
    device tensors of coefficients (any length <= 4n).  Returns the 4n coefficients of the quotient."""
    from .domain import GpuEvaluationDomain
    n = domain.size()
    assert n >= 5                                                   # quotient_poly.rs:44
    d4 = GpuEvaluationDomain.new(4 * n, domain.ctx)
    if d4 is None:
        raise ValueError("InvalidEvalDomainSize")
    polys = dict(z1=z1_poly, z2=z2_poly, a=a_poly, b=b_poly, c=c_poly, pi=pi_poly, t=t_poly, h1=h1_poly, h2=h2_poly)
    wit = [_coset_4n(d4, polys[k], polys[k].numel() // 4) for k in WIT_ORDER]
    ch = np.stack([_fr(x) for x in (alpha, beta, gamma, delta, epsilon)])
    out = _empty_like_fr(4 * n, wit[0])
    domain.ctx.quotient_evals_dev(domain.log_size(), ch, wit, [epk[k] for k in EPK_ORDER], out)
    return d4.coset_ifft_in_place(out)                              # poly_from_coset_evals


def add_blinders_to_poly(ctx, coeffs, length, blinders):
    """coeffs: device buffer with room for length + k coefficients.  Returns the new length."""
    b = np.ascontiguousarray(blinders, dtype=np.uint64).reshape(-1, 4)
    ctx.poly_add_blinders_dev(coeffs, length, b)
    return length + b.shape[0]


def kzg_open(kzg, polys, lens, point, eta):
    """SonicKZG10::open for polynomials resident in HBM: combined = sum_i eta^i p_i; witness = (combined - combined(z))
    / (X - z); returns (commitment to the witness, is_inf, combined(z)).  point/eta: Montgomery limbs or ints."""
    import torch
    ctx = kzg.ctx
    z, e = _fr(point), _fr(eta)
    e_int = field.from_mont(field.limbs_to_int(e))
    scal = np.array([field.int_to_limbs(field.to_mont(pow(e_int, i, field.R_MOD))) for i in range(len(polys))],
                    dtype=np.uint64)
    m = max(lens)
    comb = torch.empty((m, 4), dtype=torch.int64, device=polys[0].device)
    ctx.poly_lincomb_dev(polys, lens, scal, comb, m)
    quot = torch.empty((max(m - 1, 1), 4), dtype=torch.int64, device=comb.device)
    ev = ctx.poly_divide_linear_dev(comb, m, z, quot)
    if m <= 1:
        return np.zeros(8, dtype=np.uint64), True, ev
    w, inf = kzg.commit_dev(quot, m - 1)
    return w, inf, ev
