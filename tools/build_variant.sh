#!/bin/bash
# Builds an alternative libgigs_b200.so with extra -D flags into gi-gs_b200/lib/variants/<name>/ (kernel-tuning
# experiments; select it with GIGS_LIB=<path>).  usage: tools/build_variant.sh <name> -DGIGS_BL_BATCH=128 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/gi-gs_b200/lib/variants/$name
mkdir -p $out
for f in $root/gi-gs_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c $f -o $out/$b.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libgigs_b200.so $out/*.o -lcudart
rm -f $out/*.o
echo $out/libgigs_b200.so
