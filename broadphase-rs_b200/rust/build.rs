fn main() {
    // libbroadphase_b200.so is built by broadphase-rs_b200/build.py (nvcc, sm_100a)
    let dir = std::env::var("BROADPHASE_B200_LIB_DIR").unwrap_or_else(|_| "..".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=broadphase_b200");
}
