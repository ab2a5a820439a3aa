// Link against the library built by `make -C zkt_plonk_b200/csrc` (nvcc, sm_100a) for the curve selected by a cargo feature:
// libzkb200.so (default, BN254), libzkb200_bls12_381.so (feature "bls12-381"), libzkb200_bls12_377.so (feature "bls12-377").
// Same entry points in all three; zkb_curve_info tells which one was loaded.  ZKB200_LIB_DIR overrides the in-tree location.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("ZKB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../zkt_plonk_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    let lib = if env::var("CARGO_FEATURE_BLS12_381").is_ok() {
        "zkb200_bls12_381"
    } else if env::var("CARGO_FEATURE_BLS12_377").is_ok() {
        "zkb200_bls12_377"
    } else {
        "zkb200"
    };
    println!("cargo:rustc-link-lib=dylib={}", lib);
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
