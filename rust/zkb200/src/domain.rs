//! `GpuDomain`: the `D: EvaluationDomain<F> + EvaluationDomainExt<F>` parameter of `ZKTPlonk` / `prove`
//! (plonk-core/src/plonk.rs:39-46, proof_system/prove.rs:59-74).  Metadata stays arkworks'
//! `Radix2EvaluationDomain`; the four transforms go to `zkb_ntt` (plonk-core/src/util.rs:63-140 call sites).
use crate::ctx::CTX;
use ark_bn254::Fr;
use ark_poly::{domain::DomainCoeff, EvaluationDomain, Radix2EvaluationDomain};
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize, Read, SerializationError, Write};   // the derive expands to these
use core::ffi::c_int;
use plonk_core::util::EvaluationDomainExt;
use zkb200_sys as sys;

#[derive(Copy, Clone, Hash, Eq, PartialEq, Debug, CanonicalSerialize, CanonicalDeserialize)]
pub struct GpuDomain {
    inner: Radix2EvaluationDomain<Fr>,
}

impl GpuDomain {
    /// `T` is `Fr` at every call site of the prover (the generic bound only exists for group-valued FFTs).
    fn run<T: DomainCoeff<Fr>>(&self, v: &mut Vec<T>, inverse: c_int, coset: c_int) {
        assert_eq!(core::mem::size_of::<T>(), 32, "GpuDomain transforms Vec<Fr> only");
        let len = v.len().min(self.size());
        v.resize(self.size(), T::zero()); // what ark-poly's *_in_place do before transforming
        CTX.with(|c| {
            let rc = unsafe {
                sys::zkb_ntt(c.raw(), v.as_mut_ptr() as *mut u64, len, self.inner.log_size_of_group, inverse, coset)
            };
            c.check(rc).expect("zkb_ntt") // the trait methods are infallible upstream
        })
    }
}

impl EvaluationDomain<Fr> for GpuDomain {
    type Elements = <Radix2EvaluationDomain<Fr> as EvaluationDomain<Fr>>::Elements;

    /// `None` above 2^28 -> `Error::InvalidEvalDomainSize` in the caller (prove.rs:77-81).
    fn new(num_coeffs: usize) -> Option<Self> {
        Radix2EvaluationDomain::new(num_coeffs).map(|inner| Self { inner })
    }
    fn compute_size_of_domain(num_coeffs: usize) -> Option<usize> {
        Radix2EvaluationDomain::<Fr>::compute_size_of_domain(num_coeffs)
    }
    fn size(&self) -> usize {
        self.inner.size()
    }
    fn fft_in_place<T: DomainCoeff<Fr>>(&self, coeffs: &mut Vec<T>) {
        self.run(coeffs, 0, 0)
    }
    fn ifft_in_place<T: DomainCoeff<Fr>>(&self, evals: &mut Vec<T>) {
        self.run(evals, 1, 0)
    }
    fn coset_fft_in_place<T: DomainCoeff<Fr>>(&self, coeffs: &mut Vec<T>) {
        self.run(coeffs, 0, 1)
    }
    fn coset_ifft_in_place<T: DomainCoeff<Fr>>(&self, evals: &mut Vec<T>) {
        self.run(evals, 1, 1)
    }
    fn evaluate_all_lagrange_coefficients(&self, tau: Fr) -> Vec<Fr> {
        self.inner.evaluate_all_lagrange_coefficients(tau)
    }
    fn vanishing_polynomial(&self) -> ark_poly::univariate::SparsePolynomial<Fr> {
        self.inner.vanishing_polynomial()
    }
    fn evaluate_vanishing_polynomial(&self, tau: Fr) -> Fr {
        self.inner.evaluate_vanishing_polynomial(tau)
    }
    fn element(&self, i: usize) -> Fr {
        self.inner.element(i)
    }
    fn elements(&self) -> Self::Elements {
        self.inner.elements()
    }
}

impl EvaluationDomainExt<Fr> for GpuDomain {
    // plonk-core/src/util.rs:27-36
    fn log_size_of_group(&self) -> u32 {
        self.inner.log_size_of_group
    }
    fn group_gen(&self) -> Fr {
        self.inner.group_gen
    }
}
