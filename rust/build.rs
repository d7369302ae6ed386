// Addition to rs-tfhe's build.rs (the reference already uses `cc` for SPQLIOS, build.rs:7-24):
// compile the CUDA sources for sm_100a and link the C ABI.  No multi-backend dispatch.
fn build_cuda() {
    let srcs = ["engine.cu", "blind_rotate.cu", "blind_rotate_s.cu", "fft_seam.cu", "keyswitch.cu", "keyswitch_umma.cu", "keygen.cu", "aux.cu"];
    let dir = std::path::Path::new("rs_tfhe_b200/csrc");
    let out = std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let lib = out.join("libtfhe_b200.so");
    let status = std::process::Command::new("nvcc")
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .args(srcs.iter().map(|s| dir.join(s)))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=tfhe_b200");
    for s in srcs { println!("cargo:rerun-if-changed={}", dir.join(s).display()); }
}
