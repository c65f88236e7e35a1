#!/bin/bash
# Build tuning variants of the library: tools/build_variants.sh "name:-DCMX_TS=1024 -DCMX_CTAS=4" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  objs=""
  for f in capi knn knn3 idw variogram kriging; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -Xptxas -v -c climatrix_b200/csrc/$f.cu -o build/variants/${name}_$f.o 2> build/variants/${name}_$f.log
    objs="$objs build/variants/${name}_$f.o"
  done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/lib_$name.so $objs
  echo "$name: $(grep -A1 'knn_kernelILi1ELi8ELb1' build/variants/${name}_knn.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -c 'spill stores' build/variants/${name}_knn.log) $(grep -E '[1-9][0-9]* bytes spill' build/variants/${name}_knn.log | head -1)"
done
