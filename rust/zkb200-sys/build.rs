// Link against libzkb200.so built by `make -C zkt_plonk_b200/csrc` (nvcc, sm_100a).  ZKB200_LIB_DIR overrides the
// in-tree location.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("ZKB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../zkt_plonk_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
