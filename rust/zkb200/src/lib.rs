//! zkb200: the Rust host side of the B200 prover hot path.
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no rustc / cargo there): this is the source a maintainer drops
//! next to zkt-plonk.  The same C ABI is exercised end to end by the C++ mirror (include/zkb200.hpp,
//! tests/cpp/test_mirror.cpp) and the Python mirror (zkt_plonk_b200/), which are compiled and tested.
//!
//! Three layers, smallest patch first:
//!   * [`prove_native`]   one call per proof: `zkb_plonk_prove` runs all five rounds on the GPU and returns the 802
//!                        bytes of `Proof`'s `CanonicalSerialize` (drop-in for the body of `ZKTPlonk::prove`,
//!                        plonk-core/src/plonk.rs:94-111);
//!   * [`GpuDomain`]      the `D` parameter (plonk-core/src/plonk.rs:39-46, util.rs:27-140);
//!   * [`GpuKZG10Pc`]     the `PC` parameter as a type: `impl PolynomialCommitment<Fr, DensePolynomial<Fr>>` +
//!                        `impl HomomorphicCommitment<Fr>` over SonicKZG10's associated types (commitment.rs:9-46); setup /
//!                        trim / check delegate, commit / open / multi_scalar_mul run on the GPU ([`GpuKZG10`] helpers).
//!                        rust/patches/instance.rs.patch swaps it (and `GpuDomain`) into bin/src/instance.rs:67-84.
//!   * [`GpuIpaPc`]       the same seam for the reference's other scheme, `IPA<G1Affine, Blake2s>` (commitment.rs:49-86):
//!                        `commit` and `multi_scalar_mul` on the GPU, `open` / `check` delegated to ark-poly-commit.
mod ctx;
mod domain;
mod ipa;
mod kzg;
mod pc;
mod prover;

pub use ctx::{Ctx, Error};
pub use domain::GpuDomain;
pub use ipa::GpuIpaPc;
pub use kzg::GpuKZG10;
pub use pc::GpuKZG10Pc;
pub use prover::{prove_native, verify_native, NativeKey, Transcript};
