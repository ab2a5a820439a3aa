//! `GpuIpaPc`: the reference's other commitment scheme, `IPA<G, D> = InnerProductArgPC<G, D, DensePolynomial<_>>`
//! (plonk-core/src/commitment.rs:49-86; exercised by plonk-core/src/test.rs:73-84), over BN254 G1 with Blake2s, as a `PC` type.
//!
//! What moves to the GPU is what IPA shares with KZG10 on this path -- the MSMs over a fixed key:
//!   * `commit`: `comm = sum_i coeff_i * ck.comm_key[i]` (ipa_pc `cm_commit` without hiding: `rng = None`, no hiding bounds, as
//!     plonk-core always calls it) -> the same resident-key MSM as KZG10 (`zkb_srs_load_g1` on `comm_key`, `zkb_msm_g1`);
//!   * `HomomorphicCommitment::multi_scalar_mul` (commitment.rs:60-86) -> `zkb_msm_g1_bases`.
//!   * `open`: the log(n) folding rounds run in HBM -- `zkb_ipa_round_lr_dev` (L and R of a round: the two MSMs over halves of the key,
//!     the two inner products and the `h'` terms) and `zkb_ipa_round_fold_dev` (coefficients, point powers and key folded in place,
//!     csrc/ipa.cu); the hash between the rounds and the proof struct stay arkworks' (`ro_challenge` below
//!     repeats ipa_pc's private `compute_random_oracle_challenge`).  Hiding or degree-bounded openings (never asked for by
//!     plonk-core) delegate to ark-poly-commit, as do `setup` / `trim` / `check`.
//! Associated types are InnerProductArgPC's own, so proofs keep their bytes.
//!
//! NOT COMPILED HERE (no rustc / cargo in the build image); signatures are ark-poly-commit 0.3.0's as recalled.
use crate::ctx::CTX;
use crate::kzg::{pack_points, unpack_point};
use ark_bn254::{Fr, G1Affine};
use ark_ec::{AffineCurve, ProjectiveCurve};
use ark_ff::{to_bytes, Field, One, PrimeField, Zero};
use ark_poly::univariate::DensePolynomial;
use ark_poly_commit::{ipa_pc, LabeledCommitment, LabeledPolynomial, PCRandomness, PolynomialCommitment};
use ark_std::rand::RngCore;
use blake2::{Blake2s, Digest};
use core::ffi::c_int;
use plonk_core::commitment::{HomomorphicCommitment, IPA};
use std::cell::Cell;
use zkb200_sys as sys;

type Poly = DensePolynomial<Fr>;
type Inner = IPA<G1Affine, Blake2s>;

pub struct GpuIpaPc;

thread_local! {
    static RESIDENT: Cell<(usize, usize)> = Cell::new((0, 0));      // (address, length) of the comm_key resident on the GPU
}

fn ensure_resident(key: &[G1Affine]) {
    let id = (key.as_ptr() as usize, key.len());
    RESIDENT.with(|r| {
        if r.get() != id {
            CTX.with(|ctx| {
                let xy = pack_points(key);
                ctx.check(unsafe { sys::zkb_srs_load_g1(ctx.raw(), xy.as_ptr(), key.len()) })?;
                ctx.check(unsafe { sys::zkb_srs_precompute(ctx.raw(), 0) })
            })
            .expect("zkb200: comm_key upload failed");
            r.set(id);
        }
    });
}

/// ipa_pc's private `compute_random_oracle_challenge`: Blake2s(bytes || i) for i = 0, 1, .. until the digest is a field element.
fn ro_challenge(bytes: &[u8]) -> Fr {
    let mut i = 0u64;
    loop {
        let hash = Blake2s::digest(&to_bytes![bytes, i].unwrap());
        if let Some(c) = Fr::from_random_bytes(&hash) {
            return c;
        }
        i += 1;
    }
}

impl PolynomialCommitment<Fr, Poly> for GpuIpaPc {
    type UniversalParams = <Inner as PolynomialCommitment<Fr, Poly>>::UniversalParams;
    type CommitterKey = <Inner as PolynomialCommitment<Fr, Poly>>::CommitterKey;
    type VerifierKey = <Inner as PolynomialCommitment<Fr, Poly>>::VerifierKey;
    type PreparedVerifierKey = <Inner as PolynomialCommitment<Fr, Poly>>::PreparedVerifierKey;
    type Commitment = <Inner as PolynomialCommitment<Fr, Poly>>::Commitment;
    type PreparedCommitment = <Inner as PolynomialCommitment<Fr, Poly>>::PreparedCommitment;
    type Randomness = <Inner as PolynomialCommitment<Fr, Poly>>::Randomness;
    type Proof = <Inner as PolynomialCommitment<Fr, Poly>>::Proof;
    type BatchProof = <Inner as PolynomialCommitment<Fr, Poly>>::BatchProof;
    type Error = <Inner as PolynomialCommitment<Fr, Poly>>::Error;

    fn setup<R: RngCore>(max_degree: usize, num_vars: Option<usize>, rng: &mut R) -> Result<Self::UniversalParams, Self::Error> {
        Inner::setup(max_degree, num_vars, rng)
    }

    fn trim(
        pp: &Self::UniversalParams,
        supported_degree: usize,
        supported_hiding_bound: usize,
        enforced_degree_bounds: Option<&[usize]>,
    ) -> Result<(Self::CommitterKey, Self::VerifierKey), Self::Error> {
        Inner::trim(pp, supported_degree, supported_hiding_bound, enforced_degree_bounds)
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
        if rng.is_some() || polys.iter().any(|p| p.degree_bound().is_some() || p.hiding_bound().is_some()) {
            return Inner::commit(ck, polys, rng);                  // hiding / shifted commitments: never asked for by plonk-core
        }
        ensure_resident(&ck.comm_key);
        let mut comms = Vec::with_capacity(polys.len());
        let mut rands = Vec::with_capacity(polys.len());
        for p in polys {
            let coeffs = &p.polynomial().coeffs;
            let bigints: Vec<_> = coeffs.iter().map(|c| c.into_repr()).collect();
            let (mut xy, mut inf) = ([0u64; 8], 0 as c_int);
            CTX.with(|ctx| {
                ctx.check(unsafe { sys::zkb_msm_g1(ctx.raw(), bigints.as_ptr() as *const u64, 0, bigints.len(), xy.as_mut_ptr(), &mut inf) })
            })
            .expect("zkb200: zkb_msm_g1 failed");
            let comm = ipa_pc::Commitment { comm: unpack_point(&xy, inf), shifted_comm: None };
            comms.push(LabeledCommitment::new(p.label().to_string(), comm, None));
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
        let comms: Vec<_> = commitments.into_iter().collect();
        let n = ck.comm_key.len();
        if rng.is_some() || !n.is_power_of_two() || polys.iter().any(|p| p.degree_bound().is_some() || p.hiding_bound().is_some()) {
            return Inner::open_individual_opening_challenges(ck, polys, comms, point, opening_challenges, rands, rng);
        }
        // combined polynomial and commitment: sum_j challenge(j) * p_j, as ipa_pc::open does before the rounds
        let mut coeffs = vec![Fr::zero(); n];
        let mut combined = <G1Affine as AffineCurve>::Projective::zero();
        for (j, (p, c)) in polys.iter().zip(&comms).enumerate() {
            // ipa_pc draws TWO opening challenges per polynomial (the second one is for its shifted polynomial, unused without a
            // degree bound): polynomial j is weighted by challenge 2 j
            let ch = opening_challenges(2 * j as u64);
            for (acc, v) in coeffs.iter_mut().zip(&p.polynomial().coeffs) {
                *acc += ch * v;
            }
            combined += c.commitment().comm.mul(ch.into_repr());
        }
        let combined = combined.into_affine();
        let mut z = Vec::with_capacity(n);
        let mut cur = Fr::one();
        for _ in 0..n {
            z.push(cur);
            cur *= point;
        }
        let value: Fr = coeffs.iter().zip(&z).map(|(a, b)| *a * b).sum();
        let mut x = ro_challenge(&to_bytes![combined, point, value].unwrap());
        let h_prime = ck.h.mul(x.into_repr()).into_affine();
        let (mut l_vec, mut r_vec) = (Vec::new(), Vec::new());
        let (final_comm_key, c) = CTX.with(|ctx| -> Result<(G1Affine, Fr), crate::ctx::Error> {
            // Fr / G1Affine are Montgomery limbs in memory: uploaded as they are
            let d_c = ctx.upload(coeffs.as_ptr() as *const u64, 32 * n)?;
            let d_z = ctx.upload(z.as_ptr() as *const u64, 32 * n)?;
            let key_xy = pack_points(&ck.comm_key);
            let d_k = ctx.upload(key_xy.as_ptr(), 8 * key_xy.len())?;
            let h_prime_xy = pack_points(&[h_prime]);
            let mut m = n;
            while m > 1 {
                let (mut lxy, mut rxy, mut li, mut ri) = ([0u64; 8], [0u64; 8], 0 as c_int, 0 as c_int);
                let (mut ipl, mut ipr) = (Fr::zero(), Fr::zero());
                ctx.check(unsafe {
                    sys::zkb_ipa_round_lr_dev(ctx.raw(), d_c.ptr(), d_z.ptr(), d_k.ptr(), m, h_prime_xy.as_ptr(), lxy.as_mut_ptr(), &mut li,
                                              rxy.as_mut_ptr(), &mut ri, &mut ipl as *mut Fr as *mut u64, &mut ipr as *mut Fr as *mut u64)
                })?;
                let (l, r) = (unpack_point(&lxy, li), unpack_point(&rxy, ri));       // the h' terms are already inside
                x = ro_challenge(&to_bytes![x, l, r].unwrap());
                let x_inv = x.inverse().unwrap();
                ctx.check(unsafe {
                    sys::zkb_ipa_round_fold_dev(ctx.raw(), d_c.ptr_mut(), d_z.ptr_mut(), d_k.ptr_mut(), m, &x as *const Fr as *const u64,
                                                &x_inv as *const Fr as *const u64)
                })?;
                l_vec.push(l);
                r_vec.push(r);
                m /= 2;
            }
            let (mut c, mut kxy) = (Fr::zero(), [0u64; 8]);
            ctx.download(&mut c as *mut Fr as *mut u64, d_c.ptr(), 32)?;
            ctx.download(kxy.as_mut_ptr(), d_k.ptr(), 64)?;
            Ok((unpack_point(&kxy, 0), c))
        })
        .expect("zkb200: inner-product-argument rounds failed");
        let _ = rands;
        Ok(ipa_pc::Proof { l_vec, r_vec, final_comm_key, c, hiding_comm: None, rand: None })
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
        Inner::check_individual_opening_challenges(vk, commitments, point, values, proof, opening_challenges, rng)
    }
}

impl HomomorphicCommitment<Fr> for GpuIpaPc {
    fn multi_scalar_mul(commitments: &[Self::Commitment], scalars: &[Fr]) -> Self::Commitment {
        let n = commitments.len().min(scalars.len());
        let pts = pack_points(&commitments[..n].iter().map(|c| c.comm).collect::<Vec<_>>());
        let reprs: Vec<_> = scalars[..n].iter().map(|s| s.into_repr()).collect();
        let (mut xy, mut inf) = ([0u64; 8], 0 as c_int);
        CTX.with(|ctx| {
            ctx.check(unsafe { sys::zkb_msm_g1_bases(ctx.raw(), pts.as_ptr(), reprs.as_ptr() as *const u64, n, xy.as_mut_ptr(), &mut inf) })
        })
        .expect("zkb200: zkb_msm_g1_bases failed");
        let _ = Fr::zero();
        ipa_pc::Commitment { comm: unpack_point(&xy, inf), shifted_comm: None }
    }
}
