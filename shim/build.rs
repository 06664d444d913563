// build.rs — compiles the CUDA sources of libstark_b200 with nvcc for sm_100a and links them into the crate.
// (Authored for the reference crate; this image has no Rust toolchain, see shim/README.md.)
use std::{env, path::PathBuf, process::Command};

const SOURCES: [&str; 9] = ["api.cu", "merkle.cu", "ntt.cu", "fri.cu", "stark101.cu", "peaks.cu", "fourstep.cu", "verify.cu", "multi.cu"];

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").expect("OUT_DIR"));
    // where the CUDA tree lives: next to the crate by default, or wherever STARK_B200_DIR points
    let root = PathBuf::from(env::var("STARK_B200_DIR").unwrap_or_else(|_| "stark-prover_b200".to_string()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let mut objects = Vec::new();
    for src in SOURCES {
        let input = root.join("csrc").join(src);
        let object = out.join(src.replace(".cu", ".o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                   "--expt-relaxed-constexpr", "-c"])
            .arg(&input)
            .arg("-o")
            .arg(&object)
            .status()
            .expect("failed to run nvcc");
        assert!(status.success(), "nvcc failed on {}", input.display());
        println!("cargo:rerun-if-changed={}", input.display());
        objects.push(object);
    }
    for header in ["common.hpp", "field.cuh", "handles.hpp", "host_channel.hpp", "kernels.hpp", "sha256.cuh"] {
        println!("cargo:rerun-if-changed={}", root.join("csrc").join(header).display());
    }
    println!("cargo:rerun-if-changed=include/stark_b200.h");
    let archive = out.join("libstark_b200.a");
    let status = Command::new("ar").arg("crs").arg(&archive).args(&objects).status().expect("failed to run ar");
    assert!(status.success(), "ar failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=stark_b200");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");          // NCCL is bound at run time (multi.cu), not linked
}
