// UNVERIFIED (never compiled here).  Builds liblapper_b200.so with nvcc for sm_100a and links it.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("rust_lapper_b200").join("csrc");
    // nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (rust_lapper_b200/csrc/Makefile)
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("make (nvcc) failed to start");
    assert!(status.success(), "building liblapper_b200.so failed");
    let libdir = root.join("rust_lapper_b200");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=lapper_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    for f in ["api.cu", "build.cu", "query.cu", "count_sorted.cu", "setops.cu", "engine64.cu", "index.cuh", "common.cuh",
              "packed_sector.cuh", "query_views.cuh", "abi_common.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
}
