// SOURCE ONLY (never compiled here).  Links the in-tree C-ABI library.
fn main() {
    let dir = std::env::var("POSEIDON252_B200_LIB_DIR").unwrap_or_else(|_| "../../poseidon252_b200/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=poseidon252_b200");
}
