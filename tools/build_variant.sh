#!/usr/bin/env bash
# Builds a variant of libdsep.so with extra compile-time switches for kernel experiments:
#   tools/build_variant.sh NAME -DDSEP_CONV_BK=32 ...   ->  diffsep_b200/build/variants/libdsep_NAME.so
# run with DSEP_LIB=diffsep_b200/build/variants/libdsep_NAME.so (the build/ tree ships to the GPU box).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/diffsep_b200/build/variants; mkdir -p "$out/$name"
for f in "$root"/diffsep_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$f" -o "$out/$name/$(basename "$f" .cu).o" &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out/libdsep_$name.so" "$out/$name"/*.o
echo "$out/libdsep_$name.so"
