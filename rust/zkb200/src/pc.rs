//! `GpuKZG10Pc`: the `PC` type parameter of `ZKTPlonk<F, D, PC, T, C, TABLE_SIZE>` (plonk-core/src/plonk.rs:39-46) as a
//! TYPE -- `impl PolynomialCommitment<Fr, DensePolynomial<Fr>>` + `impl HomomorphicCommitment<Fr>`
//! (plonk-core/src/commitment.rs:9-21) -- over `KZG10<Bn254> = SonicKZG10<Bn254, DensePolynomial<Fr>>`.
//!
//! Every associated type IS SonicKZG10's, so `ck` / `vk` / `Proof` files written by `compile` keep their bytes
//! (bin/src/parser.rs:5-29) and `setup` / `trim` / `check` are pure delegation.  What moves to the GPU:
//!   * `commit`  -> one `zkb_msm_g1` per polynomial against the resident `powers_of_g` (prove.rs:134,179,250,307,374;
//!                  setup.rs:105).  plonk-core never asks for hiding or degree bounds (`rng = None`, labels with
//!                  `hiding_bound = None`, `degree_bound = None`: the `label_polynomial!` macro), so `Randomness` is empty;
//!                  anything else falls back to SonicKZG10 so the impl stays total.
//!   * `open`    -> eta-combination, division by (X - z) and the witness MSM (prove.rs:381-451 -> sonic_pc open ->
//!                  kzg10::open), on the host for the O(n) part and `zkb_msm_g1` for the commitment to the witness.
//!   * `multi_scalar_mul` -> `zkb_msm_g1_bases` (proof.rs:237-281: 13 points, verifier side).
//! `trim` uploads the committer key once (thread-local context, `ctx::CTX`); the key is identified by the address and
//! length of `powers_of_g`, so a second `trim` of the same parameters does not upload again.
//!
//! NOT COMPILED HERE (no rustc / cargo in the build image): trait and method signatures are ark-poly-commit 0.3.0's as
//! recalled (`open` / `check` are provided methods there; the required ones are the `*_individual_opening_challenges`
//! forms, which is what is implemented); expect small fixes on the first `cargo check`.
use crate::ctx::CTX;
use crate::kzg::GpuKZG10;
use ark_bn254::{Bn254, Fr};
use ark_ff::{Field, One, Zero};
use ark_poly::{univariate::DensePolynomial, Polynomial, UVPolynomial};
use ark_poly_commit::{
    kzg10, sonic_pc::SonicKZG10, LabeledCommitment, LabeledPolynomial, PCRandomness, PolynomialCommitment,
};
use ark_std::rand::RngCore;
use plonk_core::commitment::HomomorphicCommitment;
use std::cell::Cell;

type Poly = DensePolynomial<Fr>;
type Sonic = SonicKZG10<Bn254, Poly>;

/// Drop-in for `KZG10<Bn254>` in bin/src/instance.rs:67-68.
pub struct GpuKZG10Pc;

thread_local! {
    /// (address, length) of the `powers_of_g` that are resident on the GPU of this thread's context.
    static RESIDENT: Cell<(usize, usize)> = Cell::new((0, 0));
}

fn ensure_resident(ck: &<Sonic as PolynomialCommitment<Fr, Poly>>::CommitterKey) {
    let id = (ck.powers_of_g.as_ptr() as usize, ck.powers_of_g.len());
    RESIDENT.with(|r| {
        if r.get() != id {
            CTX.with(|ctx| GpuKZG10::load_committer_key(ctx, ck, true)).expect("zkb200: committer key upload failed");
            r.set(id);
        }
    });
}

fn plain(p: &LabeledPolynomial<Fr, Poly>) -> bool {
    p.degree_bound().is_none() && p.hiding_bound().is_none()
}

impl PolynomialCommitment<Fr, Poly> for GpuKZG10Pc {
    type UniversalParams = <Sonic as PolynomialCommitment<Fr, Poly>>::UniversalParams;
    type CommitterKey = <Sonic as PolynomialCommitment<Fr, Poly>>::CommitterKey;
    type VerifierKey = <Sonic as PolynomialCommitment<Fr, Poly>>::VerifierKey;
    type PreparedVerifierKey = <Sonic as PolynomialCommitment<Fr, Poly>>::PreparedVerifierKey;
    type Commitment = <Sonic as PolynomialCommitment<Fr, Poly>>::Commitment;
    type PreparedCommitment = <Sonic as PolynomialCommitment<Fr, Poly>>::PreparedCommitment;
    type Randomness = <Sonic as PolynomialCommitment<Fr, Poly>>::Randomness;
    type Proof = <Sonic as PolynomialCommitment<Fr, Poly>>::Proof;
    type BatchProof = <Sonic as PolynomialCommitment<Fr, Poly>>::BatchProof;
    type Error = <Sonic as PolynomialCommitment<Fr, Poly>>::Error;

    fn setup<R: RngCore>(max_degree: usize, num_vars: Option<usize>, rng: &mut R) -> Result<Self::UniversalParams, Self::Error> {
        Sonic::setup(max_degree, num_vars, rng)
    }

    fn trim(
        pp: &Self::UniversalParams,
        supported_degree: usize,
        supported_hiding_bound: usize,
        enforced_degree_bounds: Option<&[usize]>,
    ) -> Result<(Self::CommitterKey, Self::VerifierKey), Self::Error> {
        // the upload happens lazily in commit / open (the key returned here is moved by the caller, which changes the
        // address of nothing: `powers_of_g` is a Vec, its heap buffer stays put)
        Sonic::trim(pp, supported_degree, supported_hiding_bound, enforced_degree_bounds)
    }

    fn commit<'a>(
        ck: &Self::CommitterKey,
        polynomials: impl IntoIterator<Item = &'a LabeledPolynomial<Fr, Poly>>,
        rng: Option<&mut dyn RngCore>,
    ) -> Result<(Vec<LabeledCommitment<Self::Commitment>>, Vec<Self::Randomness>), Self::Error>
    where
        Poly: 'a,
    {
        let polys: Vec<_> = polynomials.into_iter().collect();
        if rng.is_some() || !polys.iter().all(|p| plain(p)) {
            return Sonic::commit(ck, polys, rng);                    // hiding / degree bounds: never asked for by plonk-core
        }
        ensure_resident(ck);
        let mut comms = Vec::with_capacity(polys.len());
        let mut rands = Vec::with_capacity(polys.len());
        for p in polys {
            // kzg10::commit: skip leading zeros, into_repr, inner product with powers_of_g -- GpuKZG10::commit does exactly that
            let c = CTX.with(|ctx| GpuKZG10::commit(ctx, p.polynomial())).expect("zkb200: zkb_msm_g1 failed");
            comms.push(LabeledCommitment::new(p.label().to_string(), c, None));
            rands.push(Self::Randomness::empty());
        }
        Ok((comms, rands))
    }

    fn open_individual_opening_challenges<'a>(
        ck: &Self::CommitterKey,
        labeled_polynomials: impl IntoIterator<Item = &'a LabeledPolynomial<Fr, Poly>>,
        commitments: impl IntoIterator<Item = &'a LabeledCommitment<Self::Commitment>>,
        point: &'a Fr,
        opening_challenges: &dyn Fn(u64) -> Fr,
        rands: impl IntoIterator<Item = &'a Self::Randomness>,
        rng: Option<&mut dyn RngCore>,
    ) -> Result<Self::Proof, Self::Error>
    where
        Poly: 'a,
        Self::Randomness: 'a,
        Self::Commitment: 'a,
    {
        let polys: Vec<_> = labeled_polynomials.into_iter().collect();
        if !polys.iter().all(|p| plain(p)) {
            return Sonic::open_individual_opening_challenges(ck, polys, commitments, point, opening_challenges, rands, rng);
        }
        ensure_resident(ck);
        // sonic_pc::open: p = sum_j challenge(j) * p_j  (challenge(j) = eta^j for `open`), no randomness
        let mut combined = Poly::zero();
        for (j, p) in polys.iter().enumerate() {
            combined += (opening_challenges(j as u64), p.polynomial());
        }
        // kzg10::open: witness = (p(X) - p(z)) / (X - z); the remainder is dropped, `random_v = None`
        let witness = {
            let coeffs = &combined.coeffs;
            let mut w = vec![Fr::zero(); coeffs.len().saturating_sub(1)];
            let mut carry = Fr::zero();
            for k in (1..coeffs.len()).rev() {
                carry = coeffs[k] + *point * carry;
                w[k - 1] = carry;
            }
            Poly::from_coefficients_vec(w)
        };
        let w = CTX.with(|ctx| GpuKZG10::commit(ctx, &witness)).expect("zkb200: zkb_msm_g1 failed");
        Ok(kzg10::Proof { w: w.0, random_v: None })
    }

    fn check_individual_opening_challenges<'a>(
        vk: &Self::VerifierKey,
        commitments: impl IntoIterator<Item = &'a LabeledCommitment<Self::Commitment>>,
        point: &'a Fr,
        values: impl IntoIterator<Item = Fr>,
        proof: &Self::Proof,
        opening_challenges: &dyn Fn(u64) -> Fr,
        rng: Option<&mut dyn RngCore>,
    ) -> Result<bool, Self::Error>
    where
        Self::Commitment: 'a,
    {
        Sonic::check_individual_opening_challenges(vk, commitments, point, values, proof, opening_challenges, rng)
    }
}

impl HomomorphicCommitment<Fr> for GpuKZG10Pc {
    /// commitment.rs:31-46: `VariableBaseMSM::multi_scalar_mul(points, scalars.into_repr())`, arbitrary bases.
    fn multi_scalar_mul(commitments: &[Self::Commitment], scalars: &[Fr]) -> Self::Commitment {
        CTX.with(|ctx| GpuKZG10::multi_scalar_mul(ctx, commitments, scalars)).expect("zkb200: zkb_msm_g1_bases failed")
    }
}

#[cfg(test)]
mod tests {
    //! What `cargo test` must show on a machine with a B200 and libzkb200.so: the GPU scheme and SonicKZG10 agree bit for bit
    //! (this is plonk.rs:191-254 `test_full` reduced to the seam).
    use super::*;
    use ark_poly_commit::LabeledPolynomial;
    use ark_std::{test_rng, UniformRand};

    #[test]
    fn commit_and_open_match_sonic_kzg10() {
        let rng = &mut test_rng();
        let pp = GpuKZG10Pc::setup(1 << 10, None, rng).unwrap();
        let (ck, vk) = GpuKZG10Pc::trim(&pp, 1 << 10, 0, None).unwrap();
        let polys: Vec<_> = (0..3)
            .map(|i| LabeledPolynomial::new(format!("p{}", i), Poly::rand(1000 - i, rng), None, None))
            .collect();
        let (gc, gr) = GpuKZG10Pc::commit(&ck, &polys, None).unwrap();
        let (sc, sr) = Sonic::commit(&ck, &polys, None).unwrap();
        assert_eq!(gc.iter().map(|c| c.commitment().0).collect::<Vec<_>>(), sc.iter().map(|c| c.commitment().0).collect::<Vec<_>>());
        let (z, eta) = (Fr::rand(rng), Fr::rand(rng));
        let gp = GpuKZG10Pc::open(&ck, &polys, &gc, &z, eta, &gr, None).unwrap();
        let sp = Sonic::open(&ck, &polys, &sc, &z, eta, &sr, None).unwrap();
        assert_eq!(gp.w, sp.w);
        let values: Vec<_> = polys.iter().map(|p| p.evaluate(&z)).collect();
        assert!(GpuKZG10Pc::check(&vk, &gc, &z, values, &gp, eta, None).unwrap());
        let s: Vec<Fr> = (0..3).map(|_| Fr::rand(rng)).collect();
        let cs: Vec<_> = gc.iter().map(|c| c.commitment().clone()).collect();
        assert_eq!(
            <GpuKZG10Pc as HomomorphicCommitment<Fr>>::multi_scalar_mul(&cs, &s).0,
            <Sonic as HomomorphicCommitment<Fr>>::multi_scalar_mul(&cs, &s).0
        );
        let _ = Fr::one();
    }
}
