// build.rs -- compiles the CUDA sources of yagi_b200 with nvcc for sm_100a and links them.
// NOT BUILT IN THE DEVELOPMENT IMAGE (no cargo/rustc there); written to the C ABI in
// include/yagi_b200.h, which is what the Python mirror and all tests exercise.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("yagi_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());

    let mut objs = Vec::new();
    for name in ["common", "firpfbch2", "firpfbch2_fast", "firpfbch2_small", "firpfbch2_synth_fast", "firpfbch2_large", "firpfbch", "firpfbch_fast", "firfilt", "firfilt_fast"] {
        let src = csrc.join(format!("{name}.cu"));
        let obj = out.join(format!("{name}.o"));
        println!("cargo:rerun-if-changed={}", src.display());
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "-c"])
            .arg(&src)
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: yagi-b200-sys has no CPU fallback")
            .success();
        assert!(ok, "nvcc failed on {}", src.display());
        objs.push(obj);
    }
    let lib = out.join("libyagi_b200.a");
    let ok = Command::new(&nvcc).arg("-lib").arg("-o").arg(&lib).args(&objs).status().unwrap().success();
    assert!(ok, "nvcc -lib failed");

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=yagi_b200");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", root.join("include/yagi_b200.h").display());
}
