// Links libadder_b200.so.  ADDER_B200_LIB_DIR = the directory holding the library (this repository's
// adder_codec_rs_b200/ after `python -c 'import __graft_entry__ as g; g.build()'`).
fn main() {
    let dir = std::env::var("ADDER_B200_LIB_DIR").expect("set ADDER_B200_LIB_DIR to the directory of libadder_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=adder_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=ADDER_B200_LIB_DIR");
}
