// Link against fluid-rs_b200/csrc/libfluid_b200.so (built by `python -c 'import __graft_entry__ as g; g.build()'`).
fn main() {
    let dir = std::env::var("FLUID_B200_LIB_DIR").expect("set FLUID_B200_LIB_DIR to the directory of libfluid_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fluid_b200");
    println!("cargo:rerun-if-env-changed=FLUID_B200_LIB_DIR");
}
