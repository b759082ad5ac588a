// Links libyasph_gpu.so (built by `python __graft_entry__.py` / yasph2d_b200/build.py in the yasph2d_b200 repository) and the
// CUDA runtime (cudaHostRegister, for page-locking the particle Vecs).
//   YASPH_GPU_LIB_DIR   directory that holds libyasph_gpu.so            (default: ../../yasph2d_b200)
//   CUDA_HOME           CUDA toolkit root                               (default: /usr/local/cuda)
use std::env;
use std::path::PathBuf;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let lib_dir = env::var("YASPH_GPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../../yasph2d_b200"));
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=dylib=yasph_gpu");
    println!("cargo:rustc-link-lib=dylib=cudart");
    // the library is found at run time next to where it was linked from
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-env-changed=YASPH_GPU_LIB_DIR");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
}
