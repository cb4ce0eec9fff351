#!/bin/bash
# tools/build_variant.sh NAME "-DSMJ_SEL_CTAS=3 ..." : an alternative build of libsmj.so with extra compile-time knobs,
# written to tools/bin/libsmj_NAME.so (use with SMJ_LIB=... ; tools/gpu_ab.sh compares them in one GPU round trip).
set -e
name=$1; flags=$2
here=$(cd "$(dirname "$0")/.." && pwd)
out=$here/tools/bin; tmp=$(mktemp -d)
mkdir -p $out
for f in $here/pim-sort-merge-join_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -c $f -o $tmp/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libsmj_$name.so $tmp/*.o
rm -rf $tmp
echo "$out/libsmj_$name.so"
