// build.rs -- link libwoxel_b200.so (default) or build it from the CUDA sources (feature "from-source").
// UNCOMPILED in this repository's build environment (no rustc); see rust/README.md.
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=WOXEL_B200_DIR");
    let repo = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");

    #[cfg(feature = "from-source")]
    {
        let csrc = repo.join("woxel_b200/csrc");
        let mut b = cc::Build::new();
        b.cuda(true)
            .cudart("static")
            .flag("-gencode")
            .flag("arch=compute_100a,code=sm_100a")
            .flag("-std=c++17")
            .flag("-O3")
            .flag("-lineinfo")
            .flag("-fmad=false") // one rounding per source operation: bit parity with the oracle
            .include(repo.join("include"));
        for f in ["wx_api.cu", "wx_raycast.cu", "wx_sdf.cu", "wx_capture.cu"] {
            let p = csrc.join(f);
            println!("cargo:rerun-if-changed={}", p.display());
            b.file(p);
        }
        b.compile("woxel_b200");
        return;
    }

    #[cfg(not(feature = "from-source"))]
    {
        let dir = env::var("WOXEL_B200_DIR").map(PathBuf::from).unwrap_or_else(|_| repo.join("woxel_b200"));
        println!("cargo:rustc-link-search=native={}", dir.display());
        println!("cargo:rustc-link-lib=dylib=woxel_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    }
}
