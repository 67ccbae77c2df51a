#!/bin/bash
# usage: scripts/build_variant.sh <name> "<extra nvcc -D flags>"  ->  build/variants/libsrt_<name>.so  (developer experiments)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../simple_raytracer_b200/csrc"
mkdir -p ../../build/variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off"
$NV $flags -c -o /tmp/srt_api_$name.o srt_api.cu
[ -f mesh_io.o ] || g++ -std=c++17 -O2 -fPIC -ffp-contract=off -c -o mesh_io.o mesh_io.cpp
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build/variants/libsrt_$name.so /tmp/srt_api_$name.o mesh_io.o -lz
echo built build/variants/libsrt_$name.so
