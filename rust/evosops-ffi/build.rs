// Links libevosops_b200.so; point EVOSOPS_B200_LIB at the directory that holds it (evosops_b200/ in this repo).
fn main() {
    let dir = std::env::var("EVOSOPS_B200_LIB").expect("set EVOSOPS_B200_LIB to the directory of libevosops_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=evosops_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=EVOSOPS_B200_LIB");
}
