//! Builds `libpna_cuda` from the CUDA sources of this repository (`portable-network-archive_b200/csrc/abi.cu` includes
//! every kernel header) for sm_100a with `nvcc` through the `cc` crate, or links a prebuilt library.
//! There is no CPU fallback to build: without nvcc (or a prebuilt library) the build fails.
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("portable-network-archive_b200/csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/pna_cuda.h").display());
    println!("cargo:rerun-if-env-changed=PNA_CUDA_LIB_DIR");
    if cfg!(feature = "prebuilt") || env::var_os("PNA_CUDA_LIB_DIR").is_some() {
        let dir = env::var("PNA_CUDA_LIB_DIR").unwrap_or_else(|_| root.join("portable-network-archive_b200").display().to_string());
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=pna_cuda");
        return;
    }
    cc::Build::new()
        .cuda(true)
        .cudart("shared")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-lineinfo")
        .flag("-std=c++17")
        .opt_level(3)
        .include(root.join("include"))
        .file(csrc.join("abi.cu"))
        .compile("pna_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
