// Links the C ABI library built from this repository (garage_b200/libgarage_ec.so).
fn main() {
    let dir = std::env::var("GARAGE_EC_LIB_DIR").expect("set GARAGE_EC_LIB_DIR to the directory holding libgarage_ec.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=garage_ec");
    println!("cargo:rerun-if-env-changed=GARAGE_EC_LIB_DIR");
}
