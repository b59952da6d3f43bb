#!/bin/bash
# Tuning aid: builds libtraycuda variants with different CTA shapes into build/variants/.
set -e
cd "$(dirname "$0")/../tray_b200/csrc"
mkdir -p ../../build/variants
rm -f ../../build/variants/*.so
for v in "128 4 8" "128 5 8" "128 6 8" "128 4 16" "128 5 16" "256 2 8" "256 2 16" "64 8 8"; do
  set -- $v
  out=../../build/variants/libtraycuda_t$1_b$2_ch$3.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -DTRAY_TPB=$1 -DTRAY_MINB=$2 -DTRAY_CH=$3 \
    -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v -shared -o $out tray_api.cu -lcudart_static -ldl -lrt -lpthread 2>&1 \
    | grep -A2 "trace_kernelIdLb0ELi$1ELi$2ELi3E" | grep -E "Used|spill" | tr '\n' ' '
  echo " <- t$1 b$2 ch$3"
done
