// UNCOMPILED SOURCE (see ../README.md).
// Builds libpcdgpu.so with the repository's own Makefile (nvcc -gencode arch=compute_100a,code=sm_100a) and links it.
// PCDGPU_ROOT overrides the repository root (default: two levels above this crate).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = env::var("PCDGPU_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../.."));
    let csrc = root.join("pcd_b200").join("csrc");
    let status = Command::new("make")
        .arg("-C")
        .arg(&csrc)
        .arg("-j8")
        .status()
        .expect("running make for libpcdgpu.so (needs nvcc 12.9+)");
    assert!(status.success(), "make -C pcd_b200/csrc failed");
    let libdir = root.join("pcd_b200");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=pcdgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    println!("cargo:rerun-if-changed={}", root.join("include/pcdgpu.h").display());
    println!("cargo:rerun-if-changed={}", csrc.display());
}
