// build.rs of obvhs-cuda-sys (UNCOMPILED here: no Rust toolchain in the build image).
//
// Links the static archive nvcc produced (obvhs_b200/lib/libobvhs_cuda.a, built by `python -m obvhs_b200.build`), or -- with the
// `build-from-source` feature -- compiles obvhs_b200/csrc/*.cu itself with the flags that keep the arithmetic bit-exact with the
// reference's CPU path (no FMA contraction, IEEE division / sqrt, no flush-to-zero):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
//          -Xcompiler -fPIC -cudart static -c csrc/<file>.cu
// NCCL is NOT linked: comm.cu binds it at run time (dlopen libnccl.so.2).
use std::{env, path::PathBuf, process::Command};

const SOURCES: &[&str] = &[
    "api.cu", "ploc.cu", "sort.cu", "bvh2.cu", "collapse.cu", "splits.cu", "reinsertion.cu", "cwbvh_build.cu", "traverse.cu", "query.cu", "comm.cu",
];

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    if env::var("CARGO_FEATURE_BUILD_FROM_SOURCE").is_ok() {
        let out = PathBuf::from(env::var("OUT_DIR").unwrap());
        let mut objs = Vec::new();
        for src in SOURCES {
            let obj = out.join(src.replace(".cu", ".o"));
            let ok = Command::new(format!("{cuda}/bin/nvcc"))
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
                .args(["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"])
                .args(["-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-cudart", "static", "-c"])
                .arg(root.join("obvhs_b200/csrc").join(src))
                .arg("-o")
                .arg(&obj)
                .status()
                .expect("nvcc not found")
                .success();
            assert!(ok, "nvcc failed on {src}");
            println!("cargo:rerun-if-changed={}", root.join("obvhs_b200/csrc").join(src).display());
            objs.push(obj);
        }
        let lib = out.join("libobvhs_cuda.a");
        assert!(Command::new("ar").arg("rcs").arg(&lib).args(&objs).status().unwrap().success());
        println!("cargo:rustc-link-search=native={}", out.display());
    } else {
        let dir = env::var("OBVHS_CUDA_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| root.join("obvhs_b200/lib"));
        println!("cargo:rustc-link-search=native={}", dir.display());
    }
    println!("cargo:rustc-link-lib=static=obvhs_cuda");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    for l in ["stdc++", "dl", "rt", "pthread"] {
        println!("cargo:rustc-link-lib=dylib={l}");
    }
    println!("cargo:rerun-if-changed={}", root.join("include/obvhs_cuda.h").display());
}
