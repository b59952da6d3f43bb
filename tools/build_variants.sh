#!/bin/bash
# Tuning aid: builds libtraycuda variants with different CTA shapes of the pre-filter trace kernel into build/variants/.
# Each entry: threads per CTA, resident CTAs per SM the register allocation targets.
set -e
cd "$(dirname "$0")/../tray_b200/csrc"
mkdir -p ../../build/variants
rm -f ../../build/variants/*.so
for v in "128 3" "128 4" "128 5" "128 6" "256 2" "256 3" "64 8" "64 10"; do
  set -- $v
  out=../../build/variants/libtraycuda_t$1_fb$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -DTRAY_TPB=$1 -DTRAY_MINB=$2 -DTRAY_FILTER_MINB=$2 \
    -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v -shared -o $out tray_api.cu -lcudart_static -ldl -lrt -lpthread 2>&1 \
    | grep -A2 "trace_kernelIdLb0ELi$1ELi$2ELi3E" | grep -E "Used|spill" | tr '\n' ' '
  echo " <- t$1 fb$2"
done
