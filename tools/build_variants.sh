#!/bin/bash
# developer script: build kernel variants of libmphx.so (launch-bound / tuning macros) into build/variants/
# usage: tools/build_variants.sh name1:"-DA=1 -DB=2" name2:"..."
set -e
cd "$(dirname "$0")/../particlemethod_fsi_b200/csrc"
OUT=../../build/variants; mkdir -p $OUT
make -s host_constants.o io.o
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda -Xcompiler -fPIC \
     -ccbin g++ -I../../include -I. $flags -Xptxas -v -c mphx.cu -o $OUT/mphx_$name.o 2> $OUT/ptxas_$name.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin g++ -o $OUT/libmphx_$name.so $OUT/mphx_$name.o host_constants.o io.o
  echo "$name: $(grep -A3 'k_filterILi3\|k_pass1_v3ILi3ELb0ELb1\|k_pass2_v3ILi3ELb0ELb1' $OUT/ptxas_$name.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
done
