// Points rustc at libislands_b200.so.  ISLANDS_B200_LIB_DIR = the directory that holds it
// (islands_b200/lib of this repository after `make -C islands_b200/csrc`).
fn main() {
    println!("cargo:rerun-if-env-changed=ISLANDS_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("ISLANDS_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=islands_b200");
}
