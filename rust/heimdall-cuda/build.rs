// build.rs -- compiles the hand-written .cu kernels for sm_100a and links them statically.
// NOT COMPILED IN THIS ENVIRONMENT (no Rust toolchain); mirrors heimdall-vision_b200/Makefile.
use std::path::PathBuf;

fn main() {
    let root = PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../..");
    let csrc = root.join("heimdall-vision_b200/csrc");
    let sources = ["hv_api.cu", "k_preprocess.cu", "k_ccl.cu", "k_ccl_frame.cu", "k_score.cu", "k_stage.cu"];
    let mut build = cc::Build::new();
    build
        .cuda(true)
        .cudart("static")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("-std=c++17")
        .flag("--expt-relaxed-constexpr")
        .include(root.join("include"));
    for s in sources {
        build.file(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/heimdall_cuda.h").display());
    build.compile("heimdall_cuda");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
