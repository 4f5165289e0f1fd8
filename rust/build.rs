// build.rs: compiles the CUDA sources for sm_100a with nvcc and links them (no CPU fallback).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("jpeg_encoder_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libjpegenc_b200.so");
    let mut cmd = Command::new(&nvcc);
    cmd.args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib);
    for f in ["api.cu", "stage_a.cu", "entropy.cu", "tables.cu", "gather.cu", "scan.cu", "host.cpp"] {
        cmd.arg(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    let status = cmd.status().expect("nvcc not found: the B200 encode path cannot be built");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=jpegenc_b200");
}
