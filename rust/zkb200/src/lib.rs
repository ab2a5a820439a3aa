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
//!   * [`GpuKZG10`]       helpers for the `PC` parameter: commit / multi_scalar_mul on the resident committer key
//!                        (commitment.rs:10-46).
mod ctx;
mod domain;
mod kzg;
mod prover;

pub use ctx::{Ctx, Error};
pub use domain::GpuDomain;
pub use kzg::GpuKZG10;
pub use prover::{prove_native, verify_native, NativeKey, Transcript};
