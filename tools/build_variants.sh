#!/bin/bash
# Tuning aid: builds libtraycuda variants into build/variants/ (they travel to the GPU box with the snapshot).
#   tools/build_variants.sh name1="-DFLAG=.. -DFLAG2=.." name2="..."        (builds run in parallel)
# Time them with tools/gpu_wf_variants.py (whole frame, every layout, image hash).
set -e
cd "$(dirname "$0")/../tray_b200/csrc"
mkdir -p ../../build/variants
rm -f ../../build/variants/*.so ../../build/variants/*.log
for v in "$@"; do
  name=${v%%=*}; flags=${v#*=}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false $flags \
      -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v -shared -o ../../build/variants/libtraycuda_$name.so tray_api.cu \
      -lcudart_static -ldl -lrt -lpthread > ../../build/variants/$name.log 2>&1
    echo "$name: $(grep -A2 'trace_kernelIdLb0ELi[0-9]*ELi[0-9]*ELi5ELb0' ../../build/variants/$name.log | grep -E 'Used|spill' | tr '\n' ' ')" ) &
done
wait
