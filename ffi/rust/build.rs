// Link against the prebuilt libnerf_b200.so (built by `python -m nerf_rs_b200.build`, nvcc sm_100a).
fn main() {
    let dir = std::env::var("NERF_B200_LIB_DIR").unwrap_or_else(|_| "../../nerf_rs_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=nerf_b200");
    println!("cargo:rerun-if-env-changed=NERF_B200_LIB_DIR");
}
