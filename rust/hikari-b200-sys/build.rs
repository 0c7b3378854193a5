//! Builds libhikari_b200 from the CUDA / C++ sources of this repository with `cc` driving nvcc, as BASELINE.json's north_star
//! sketches ("cudarc + cc"), or links a prebuilt library when HIKARI_B200_LIB_DIR is set.
//! NOT COMPILED in the repository's build container (it has no Rust toolchain): the flags below are the ones
//! bevy_hikari_b200/build.py uses there.
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    if let Ok(dir) = env::var("HIKARI_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=hikari_b200");
        return;
    }
    let csrc = root.join("bevy_hikari_b200/csrc");
    let mut cuda = cc::Build::new();
    cuda.cuda(true)
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")      // B200 only, no other targets
        .flag("-lineinfo").flag("-fmad=false")                          // explicit fmaf only: the parity rule of include/hk_math.h
        .flag("-std=c++17").flag("-O3")
        .include(root.join("include")).include(&csrc);
    for f in ["context.cu", "kernels_light.cu", "kernels_pool.cu", "kernels_spatial.cu", "kernels_post.cu", "kernels_upscale.cu"] {
        cuda.file(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    cuda.compile("hikari_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", root.join("include/hikari_b200.h").display());
}
