#!/bin/bash
# Time every build/variants/lib_*.so on the C5 and C3 single-slice problems.
cd "$(dirname "$0")/.."
for lib in build/variants/lib_*.so; do
  name=$(basename $lib .so)
  a=$(CMX_LIB=$PWD/$lib python tools/prof_knn.py 1000000 1801 3600 32 2 2>&1 | tail -1)
  b=$(CMX_LIB=$PWD/$lib python tools/prof_knn.py 100000 721 1440 32 3 2>&1 | tail -1)
  echo "$name | $a | $b"
done
