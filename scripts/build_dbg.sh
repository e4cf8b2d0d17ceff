#!/bin/bash
# Developer build with what-if timing knobs (-DFFB_DEBUG_KNOBS $FFB_EXTRA) -> build/dbg/libffsim_b200.so
set -e
cd "$(dirname "$0")/../ffsim_b200/csrc"
out=../../build/${FFB_DBG_DIR:-dbg}; mkdir -p $out
for f in tables.cpp plan.cpp; do nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -Xcompiler -fPIC -x cu -c $f -o $out/$f.o; done
for f in givens_kernels.cu diag_kernels.cu exchange_kernels.cu capi.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -DFFB_DEBUG_KNOBS $FFB_EXTRA -Xcompiler -fPIC -c $f -o $out/$f.o; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/libffsim_b200.so $out/*.o
