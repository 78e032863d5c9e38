#!/bin/bash
# Builds A/B variants of libsdimb.so into variants/ (git-ignored, travels with gpurun).
# usage: tools/build_variants.sh name "-DFLAG=1 -DOTHER=2" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -I include \
       $flags -Xptxas -v -o variants/libsdimb_$name.so sdim_b200/csrc/sdimb.cu 2> variants/$name.ptxas.txt &
done
wait
sleep 1
touch variants/*.so    # newer than the sources: sdim_b200.build must not rebuild over them
grep -A2 "interp_planes_kernelILi[23]ELb1" variants/*.ptxas.txt | grep -E "registers|spill"
