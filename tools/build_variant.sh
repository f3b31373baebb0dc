#!/bin/bash
# Builds build_variants/lib_<name>.so from the working tree with extra nvcc defines: tools/build_variant.sh <name> [-DX=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../adder_codec_rs_b200/csrc"
mkdir -p ../../build_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -Xptxas -v "$@" -shared -o ../../build_variants/lib_${name}.so adder_b200.cu 2> /tmp/ptxas_${name}.log
grep -A3 "integrate_frame_kernelILi8ELb0ELb1" /tmp/ptxas_${name}.log | grep -E "spill|Used" 
