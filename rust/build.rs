// Links libbasic_dsp_b200.so (CUDA kernels + C ABI, built by `python -m basic_dsp_b200.build`, nvcc -gencode
// arch=compute_100a,code=sm_100a).  BASIC_DSP_B200_DIR points at the directory holding the library; with
// BASIC_DSP_B200_BUILD=1 the library is built here through `cc` driving nvcc (cc::Build::cuda), which is what
// north_star's "thin C-ABI FFI shim built with cc/bindgen" asks for.  src/ffi.rs is generated from the header by
// rust/gen_ffi.py (bindgen's job; bindgen needs libclang, which this image does not have).
use std::env;
use std::path::PathBuf;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    println!("cargo:rerun-if-changed={}", root.join("include/basic_dsp_b200.h").display());
    if env::var("BASIC_DSP_B200_BUILD").map(|v| v == "1").unwrap_or(false) {
        let csrc = root.join("basic_dsp_b200/csrc");
        let mut b = cc::Build::new();
        b.cuda(true)
            .cudart("static")
            .flag("-gencode")
            .flag("arch=compute_100a,code=sm_100a")
            .flag("-O3")
            .flag("-lineinfo")
            .flag("-std=c++17")
            .include(root.join("include"));
        for f in &[
            "common.cu", "fft.cu", "conv.cu", "ols4096i.cu", "ols8192i.cu", "ols64.cu", "fftp.cu", "fftp16k.cu", "fftc.cu", "interp.cu", "elementwise.cu",
            "mathops.cu", "reduce.cu", "capi.cu",
        ] {
            b.file(csrc.join(f));
            println!("cargo:rerun-if-changed={}", csrc.join(f).display());
        }
        b.compile("basic_dsp_b200");   // static archive linked into the crate
    } else {
        let dir = env::var("BASIC_DSP_B200_DIR").unwrap_or_else(|_| root.join("basic_dsp_b200").display().to_string());
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=basic_dsp_b200");
    }
}
