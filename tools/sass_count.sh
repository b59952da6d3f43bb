#!/bin/bash
# Tuning aid: SASS instruction count of the default trace kernel (strict fp64, cluster scan, plain layout) in a library build.
# The kernel's per-segment code sits at the size of the 32 KB instruction cache, so its size is watched at every change.
for so in "$@"; do
  cuobjdump -sass "$so" 2>/dev/null | awk -v so="$so" '
    /Function : /{inside = ($3 ~ /trace_kernelIdLb0ELi128ELi[0-9]+ELi5ELb0/)}
    inside && /^[ \t]+\/\*[0-9a-f]+\*\/ +[A-Z@]/{n++}
    END{printf "%s: %d SASS instructions (%.1f KB)\n", so, n, n*16/1024.0}'
done
