#!/bin/bash
# Tuning builds: tools/build_variant.sh <tag> <extra nvcc flags...> -> gpurun_variants/libnode_b200_<tag>.so (travels to the GPU box;
# select it with NODE_B200_LIB). Objects go to /tmp so the product build is left alone.
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/gpurun_variants; mkdir -p $out /tmp/nodevar_$tag
cd $root/neural-ode-features_b200/csrc
for f in *.cu; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $f -o /tmp/nodevar_$tag/${f%.cu}.o ) &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libnode_b200_$tag.so /tmp/nodevar_$tag/*.o
echo $out/libnode_b200_$tag.so
