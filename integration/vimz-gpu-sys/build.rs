// Link against libvimz_gpu.so built by `make` in the vimz-b200 repository.
fn main() {
    let dir = std::env::var("VIMZ_GPU_LIB_DIR").unwrap_or_else(|_| "../../vimz_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=vimz_gpu");
    println!("cargo:rerun-if-env-changed=VIMZ_GPU_LIB_DIR");
}
