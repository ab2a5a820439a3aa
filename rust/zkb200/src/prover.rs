//! Whole-prover entry points: `zkb_plonk_setup` / `zkb_plonk_load_keys` / `zkb_plonk_prove` behind safe wrappers.
//!
//! The smallest patch to the reference is inside `proof_system::prove` (plonk-core/src/proof_system/prove.rs:59-470),
//! after `composer.pad_to(n)`: hand the padded wire vectors, the lookup table and the public inputs to
//! [`prove_native`] and deserialize the returned bytes:
//!
//! ```ignore
//! let (a, b, c) = composer.wire_evals();                                   // prove.rs:49-55
//! let blinders: Vec<Fr> = (0..19).map(|_| Fr::rand(rng)).collect();        // same draw order as prove.rs:125-296
//! let bytes = zkb200::prove_native(&ctx, &key, &a, &b, &c, &table, &pi_values, &blinders)?;
//! let proof = Proof::<Fr, D, PC>::deserialize(&bytes[..])?;                // proof.rs:106-155
//! ```
use crate::ctx::{Ctx, Error};
use ark_bn254::Fr;
use std::{ffi::CString, path::Path};
use zkb200_sys as sys;

/// The `T: TranscriptProtocol` parameter: `MerlinTranscript` (plonk-core/src/transcript.rs:49-109) or
/// `EthereumTranscript` (gadgets/src/transcript.rs:8-90).
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Transcript {
    Merlin = 0,
    Ethereum = 1,
}

/// Proving key resident in HBM: selector / sigma polynomials, the extended key's coset tables, verifier-key commitments.
pub struct NativeKey<'a> {
    ctx: &'a Ctx,
    raw: *mut sys::zkb_plonk_pk,
}

impl<'a> NativeKey<'a> {
    /// `proof_system::setup` (setup.rs:42-166, extend = true) from the padded composer columns.
    /// selectors = [q_m, q_l, q_r, q_o, q_c, q_lookup], sigma = compute_all_sigma_evals' three vectors (n each).
    pub fn setup(
        ctx: &'a Ctx,
        log_n: u32,
        selectors: [&[Fr]; 6],
        sigma: [&[Fr]; 3],
        table_size: usize,
        pi_positions: &[usize],
    ) -> Result<Self, Error> {
        let n = 1usize << log_n;
        assert!(selectors.iter().chain(sigma.iter()).all(|v| v.len() == n));
        let s: Vec<*const u64> = selectors.iter().map(|v| v.as_ptr() as *const u64).collect();
        let g: Vec<*const u64> = sigma.iter().map(|v| v.as_ptr() as *const u64).collect();
        let mut raw = core::ptr::null_mut();
        ctx.check(unsafe {
            sys::zkb_plonk_setup(ctx.raw(), log_n, s.as_ptr(), g.as_ptr(), table_size, pi_positions.as_ptr(), pi_positions.len(), &mut raw)
        })?;
        Ok(Self { ctx, raw })
    }

    /// The `pk` / `vk` files `zkt compile` wrote (bin/src/main.rs:106-112), as `prove-withdraw` reads them (main.rs:274-281).
    pub fn load(ctx: &'a Ctx, pk_path: &Path, vk_path: &Path, table_size: usize) -> Result<Self, Error> {
        let (p, v) = (CString::new(pk_path.to_str().unwrap()).unwrap(), CString::new(vk_path.to_str().unwrap()).unwrap());
        let mut raw = core::ptr::null_mut();
        ctx.check(unsafe { sys::zkb_plonk_load_keys(ctx.raw(), p.as_ptr(), v.as_ptr(), table_size, &mut raw) })?;
        Ok(Self { ctx, raw })
    }

    pub fn set_transcript(&mut self, t: Transcript) -> Result<(), Error> {
        self.ctx.check(unsafe { sys::zkb_plonk_pk_set_transcript(self.raw, t as i32) })
    }
}

impl Drop for NativeKey<'_> {
    fn drop(&mut self) {
        unsafe { sys::zkb_plonk_pk_destroy(self.ctx.raw(), self.raw) }
    }
}

/// `proof_system::prove`: wires a, b, c (n each, padded), the lookup table's entries, one value per public-input
/// row, the 19 blinders in the reference's draw order (a 2, b 2, c 2, h1 3, h2 2, z1 3, z2 3, b0, b1).
/// Returns `Proof`'s serialised bytes: 802 on BN254, 1010 on the BLS12 builds (`zkb_plonk_proof_bytes`).
pub fn prove_native(
    ctx: &Ctx,
    key: &NativeKey,
    a: &[Fr],
    b: &[Fr],
    c: &[Fr],
    table: &[Fr],
    pi_values: &[Fr],
    blinders: &[Fr],
) -> Result<Vec<u8>, Error> {
    assert_eq!(blinders.len(), 19);
    let mut out = vec![0u8; unsafe { sys::zkb_plonk_proof_bytes() }];
    let p = |v: &[Fr]| v.as_ptr() as *const u64; // Vec<Fr> is a dense array of 4 x u64 Montgomery limbs
    ctx.check(unsafe {
        sys::zkb_plonk_prove(
            ctx.raw(),
            key.raw,
            p(a),
            p(b),
            p(c),
            p(table),
            table.len(),
            p(pi_values),
            p(blinders),
            out.as_mut_ptr(),
            core::ptr::null_mut(),
        )
    })?;
    Ok(out)
}

/// `Proof::verify` (plonk-core/src/proof_system/proof.rs:285-503) on the host: transcript replay, r0, the 13-point
/// linearisation commitment, and `PC::check` twice as products of pairings.  `vk_xy`: the ten VerifierKey commitments in
/// seed_transcript order (x || y Montgomery limbs, identity = zeros); `h`, `beta_h`: the G2 half of the commitment
/// scheme's verifier key as x.c0 x.c1 y.c0 y.c1 Montgomery limbs (`G2Affine`'s in-memory order).
/// Ok(()) = accepted, Err(step) mirrors `Error::ProofVerificationError { step }`; malformed input is step 0.
pub fn verify_native(
    n: usize,
    pi_roots: &[Fr],
    vk_xy: &[u64; 80],
    pub_inputs: &[Fr],
    proof: &[u8; 802],
    h: &[u64; 16],
    beta_h: &[u64; 16],
    transcript: Transcript,
) -> Result<(), u32> {
    assert_eq!(pi_roots.len(), pub_inputs.len());
    let rc = unsafe {
        sys::zkb_plonk_verify(
            n,
            pi_roots.as_ptr() as *const u64,
            pi_roots.len(),
            vk_xy.as_ptr(),
            core::ptr::null(),
            pub_inputs.as_ptr() as *const u64,
            proof.as_ptr(),
            h.as_ptr(),
            beta_h.as_ptr(),
            transcript as i32,
        )
    };
    match rc {
        0 => Ok(()),
        1 | 2 => Err(rc as u32),
        _ => Err(0),
    }
}
