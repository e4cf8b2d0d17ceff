#!/bin/bash
# Developer build of a library variant: scripts/build_variant.sh <dir under build/> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/../ffsim_b200/csrc"
out=../../build/$1; shift; mkdir -p $out
for f in tables.cpp plan.cpp; do nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 "$@" -Xcompiler -fPIC -x cu -c $f -o $out/$f.o; done
for f in givens_kernels.cu diag_kernels.cu exchange_kernels.cu capi.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo "$@" -Xcompiler -fPIC -c $f -o $out/$f.o & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/libffsim_b200.so $out/*.o
