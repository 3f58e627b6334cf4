// build.rs -- compiles EVERY CUDA source of yagi_b200 (yagi_b200/csrc/*.cu) with nvcc for sm_100a and
// links the objects.  The list is globbed, not written out, so a new kernel file cannot be forgotten
// (tests/test_host_logic.py::test_rust_sys_crate_tracks_the_c_abi checks this file and src/lib.rs).
// NOT BUILT IN THE DEVELOPMENT IMAGE (no cargo/rustc there); written to the C ABI in
// include/yagi_b200.h, which is what the Python mirror and all tests exercise.
use std::env;
use std::fs;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("yagi_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());

    let mut sources: Vec<PathBuf> = fs::read_dir(&csrc)
        .expect("yagi_b200/csrc not found")
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    sources.sort();
    assert!(!sources.is_empty(), "no .cu sources under {}", csrc.display());
    // headers: any change rebuilds everything
    for e in fs::read_dir(&csrc).unwrap().filter_map(|e| e.ok()) {
        if e.path().extension().map_or(false, |x| x == "cuh") {
            println!("cargo:rerun-if-changed={}", e.path().display());
        }
    }

    let mut objs = Vec::new();
    for src in &sources {
        let obj = out.join(src.file_stem().unwrap()).with_extension("o");
        println!("cargo:rerun-if-changed={}", src.display());
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "-c"])
            .arg(src)
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
