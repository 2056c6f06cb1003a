// build.rs of the forked zkp crate: links libzkp_b200.so when the `cuda_backend` feature is on.
// ZKP_B200_LIB_DIR points at the directory that holds the library (zkp_b200/lib of this repository after
// `python -m zkp_b200.build`); the CUDA runtime is found through CUDA_HOME (default /usr/local/cuda).
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=ZKP_B200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
    if env::var_os("CARGO_FEATURE_CUDA_BACKEND").is_none() {
        return;
    }
    let lib_dir = env::var("ZKP_B200_LIB_DIR").expect("cuda_backend: set ZKP_B200_LIB_DIR to the directory of libzkp_b200.so");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    println!("cargo:rustc-link-search=native={}", lib_dir);
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=dylib=zkp_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    // let the test and bench binaries find the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir);
}
