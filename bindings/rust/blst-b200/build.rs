// Link against the prebuilt CUDA library instead of compiling sppark with nvcc (compare blst-sppark/build.rs:57-92).
// B200KZG_LIB_DIR = directory holding libb200kzg.so (rust-kzg_b200/ in this repository after `make -C rust-kzg_b200/csrc`).
use std::env;

fn main() {
    let dir = env::var("B200KZG_LIB_DIR").expect("set B200KZG_LIB_DIR to the directory containing libb200kzg.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200kzg");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=B200KZG_LIB_DIR");
}
