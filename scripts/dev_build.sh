#!/bin/bash
# 40-second development build: only the two fp32 PhaseSin kernels of BASELINE config 2 (default and continuous
# detection).  Output: build/libvsr_b200_dev.so -- copy it over the release library ON THE BOX only, never commit it.
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=false -prec-sqrt=false -ftz=true \
  -DVSR_DEV_SUBSET -DVSR_TUNING "$@" -shared -Xcompiler -fPIC -o build/libvsr_b200_dev.so 2dhmsr_b200/csrc/vsr_abi.cu
