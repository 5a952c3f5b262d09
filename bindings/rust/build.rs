// build.rs -- builds the CUDA library with nvcc for sm_100a when the `b200` feature is on.
fn main() {
    if std::env::var_os("CARGO_FEATURE_B200").is_none() { return; }
    let csrc = std::path::PathBuf::from(std::env::var("SAF_B200_CSRC")      // checkout of this repository
        .expect("set SAF_B200_CSRC to <symbanafis-b200>/symbanafis_b200/csrc"));
    let out = std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let status = std::process::Command::new("make")
        .arg("-C").arg(&csrc)
        .arg(format!("OUT={}", out.join("libsaf_b200.so").display()))
        .status().expect("make (nvcc -gencode arch=compute_100a,code=sm_100a) failed to start");
    assert!(status.success(), "building libsaf_b200.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=saf_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
