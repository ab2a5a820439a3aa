//! `GpuKZG10`: the MSM-bearing half of the `PC: HomomorphicCommitment<F>` parameter (plonk-core/src/commitment.rs:10-46)
//! for `KZG10<Bn254> = SonicKZG10<Bn254, DensePolynomial<Fr>>`.
//!
//! A full `PolynomialCommitment` impl is a newtype over `SonicKZG10` that keeps every associated type (so keys and
//! proofs serialise byte-identically) and delegates `setup` / `trim` / `check`; `trim` additionally calls
//! [`GpuKZG10::load_committer_key`], `commit` maps each polynomial to [`GpuKZG10::commit`], `open` combines with powers
//! of the opening challenge, divides by (X - z) and commits to the witness the same way (ark-poly-commit 0.3
//! kzg10::{commit, open}, reached from prove.rs:134,179,250,307,374,381,427).
use crate::ctx::{Ctx, Error};
use ark_bn254::{Bn254, Fq, Fr, G1Affine};
use ark_ff::{PrimeField, Zero};
use ark_poly::univariate::DensePolynomial;
use ark_poly_commit::{kzg10::Commitment, sonic_pc::CommitterKey};
use core::ffi::c_int;
use zkb200_sys as sys;

pub struct GpuKZG10;

/// `G1Affine` is `repr(Rust)` (x, y, infinity): repack into the ABI's x || y Montgomery limbs, identity = (0, 0).
pub(crate) fn pack_points(points: &[G1Affine]) -> Vec<u64> {
    let mut out = vec![0u64; 8 * points.len()];
    for (i, p) in points.iter().enumerate() {
        if p.infinity {
            continue;
        }
        out[8 * i..8 * i + 4].copy_from_slice(&(p.x.0).0); // Fp256(BigInteger256([u64; 4])): Montgomery limbs
        out[8 * i + 4..8 * i + 8].copy_from_slice(&(p.y.0).0);
    }
    out
}

pub(crate) fn unpack_point(xy: &[u64; 8], is_inf: c_int) -> G1Affine {
    if is_inf != 0 {
        return G1Affine::zero();
    }
    let limb = |o: usize| ark_ff::BigInteger256([xy[o], xy[o + 1], xy[o + 2], xy[o + 3]]);
    G1Affine::new(Fq::new(limb(0)), Fq::new(limb(4)), false) // Fp256::new takes the Montgomery representation
}

impl GpuKZG10 {
    /// Once per key (what `PC::trim` returns): `ck.powers_of_g` become resident in HBM; `fixed_base` also builds the
    /// window tables (`zkb_srs_precompute`).
    pub fn load_committer_key(ctx: &Ctx, ck: &CommitterKey<Bn254>, fixed_base: bool) -> Result<(), Error> {
        let xy = pack_points(&ck.powers_of_g);
        ctx.check(unsafe { sys::zkb_srs_load_g1(ctx.raw(), xy.as_ptr(), ck.powers_of_g.len()) })?;
        if fixed_base {
            ctx.check(unsafe { sys::zkb_srs_precompute(ctx.raw(), 0) })?;
        }
        Ok(())
    }

    /// kzg10::commit with `hiding_bound = None` (what prove.rs always passes): skip the leading zero coefficients,
    /// `into_repr` the rest, inner product with `powers_of_g[skipped..]`.
    pub fn commit(ctx: &Ctx, poly: &DensePolynomial<Fr>) -> Result<Commitment<Bn254>, Error> {
        let skipped = poly.coeffs.iter().take_while(|c| c.is_zero()).count();
        let bigints: Vec<_> = poly.coeffs[skipped..].iter().map(|c| c.into_repr()).collect(); // canonical 4 x u64
        let (mut xy, mut inf) = ([0u64; 8], 0 as c_int);
        ctx.check(unsafe {
            sys::zkb_msm_g1(ctx.raw(), bigints.as_ptr() as *const u64, skipped, bigints.len(), xy.as_mut_ptr(), &mut inf)
        })?;
        Ok(Commitment(unpack_point(&xy, inf)))
    }

    /// `HomomorphicCommitment::multi_scalar_mul` (commitment.rs:31-46).
    pub fn multi_scalar_mul(ctx: &Ctx, commitments: &[Commitment<Bn254>], scalars: &[Fr]) -> Result<Commitment<Bn254>, Error> {
        let n = commitments.len().min(scalars.len());
        let pts = pack_points(&commitments[..n].iter().map(|c| c.0).collect::<Vec<_>>());
        let reprs: Vec<_> = scalars[..n].iter().map(|s| s.into_repr()).collect();
        let (mut xy, mut inf) = ([0u64; 8], 0 as c_int);
        ctx.check(unsafe {
            sys::zkb_msm_g1_bases(ctx.raw(), pts.as_ptr(), reprs.as_ptr() as *const u64, n, xy.as_mut_ptr(), &mut inf)
        })?;
        Ok(Commitment(unpack_point(&xy, inf)))
    }
}
