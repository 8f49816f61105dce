#!/bin/bash
# build a tuning variant of libm324 with extra nvcc defines:  scripts/build_variant.sh <out.so> [-DFOO=1 ...]
set -e
out=$1; shift
cd "$(dirname "$0")/../motion324_b200/csrc"
tmp=$(mktemp -d)
for f in host_util gemm attention attention_bwd pointwise backward chamfer dataprep capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static "$@" -c $f.cu -o $tmp/$f.o &
done
wait
nvcc -shared -o "$out" $tmp/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
rm -rf $tmp
