#!/bin/bash
# Build A/B variants of librtx_b200.so: tools/ab_build.sh name1:"-DFOO=1 -DBAR=2" name2:"..."   -> build/ab/librtx_<name>.so
cd "$(dirname "$0")/../go-raytracing_b200/csrc"
mkdir -p ../../build/ab
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -ffp-contract=off -Xptxas -v $flags \
      -shared -o ../../build/ab/librtx_$name.so rtx_api.cu > ../../build/ab/$name.ptxas 2>&1;
    echo "$name [$flags]: $(grep -A2 'k_extendILb0' ../../build/ab/$name.ptxas | grep -E 'Used|spill' | tr '\n' ' ')" ) &
done
wait
