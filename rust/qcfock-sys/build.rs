// build.rs -- links libqcfock.so.
//
// QCFOCK_LIB_DIR  : directory holding a prebuilt libqcfock.so (default: <repo>/qchem-rs_b200)
// feature build-from-source : run `make -C <repo>/qchem-rs_b200/csrc` first (needs nvcc 12.9+, sm_100a)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.join("..").join("..");
    let lib_dir = env::var("QCFOCK_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| repo.join("qchem-rs_b200"));
    if cfg!(feature = "build-from-source") {
        let status = Command::new("make")
            .arg("-C")
            .arg(repo.join("qchem-rs_b200").join("csrc"))
            .arg("-j8")
            .status()
            .expect("failed to run make (nvcc needed)");
        assert!(status.success(), "building libqcfock.so failed");
    }
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=qcfock");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-env-changed=QCFOCK_LIB_DIR");
    println!("cargo:rerun-if-changed={}", repo.join("include").join("qcfock.h").display());
}
