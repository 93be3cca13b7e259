// SOURCE ONLY (never compiled in the build image). Links the prebuilt C-ABI library.
fn main() {
    let dir = std::env::var("H2AGG_LIB_DIR").expect("set H2AGG_LIB_DIR to the directory holding libh2agg.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=h2agg");
    println!("cargo:rerun-if-env-changed=H2AGG_LIB_DIR");
}
