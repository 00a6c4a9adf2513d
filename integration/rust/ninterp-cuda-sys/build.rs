// Links libninterp_cuda.so.  NINTERP_CUDA_LIB_DIR names the directory that holds it (the repository's
// ninterp_b200/ after `python -m ninterp_b200.build`); the library links the CUDA runtime statically, so nothing
// else is needed at link time.
fn main() {
    println!("cargo:rerun-if-env-changed=NINTERP_CUDA_LIB_DIR");
    if let Ok(dir) = std::env::var("NINTERP_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=ninterp_cuda");
}
