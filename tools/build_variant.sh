#!/bin/bash
# usage: tools/build_variant.sh <suffix> <extra nvcc flags...>   builds kmer-extension_b200/libkmer_cuda<suffix>.so (tuning only)
set -e
sfx=$1; shift
cd "$(dirname "$0")/../kmer-extension_b200/csrc"
bd=/tmp/kmer_build$sfx; mkdir -p $bd
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in api rows extract count_dense count_hash count_part match decode; do
  nvcc -O3 "$@" -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v -c $f.cu -o $bd/$f.o 2> $bd/$f.log &
done
wait
nvcc $ARCH -shared -o ../libkmer_cuda$sfx.so $bd/*.o -lcudart
grep -A1 "partition_kernelILi8ELi1" $bd/count_part.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | head -3
