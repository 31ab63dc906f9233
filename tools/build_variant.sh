#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]  ->  cuda-flash-attention_b200/build/variants/NAME.so  (A/B experiments)
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/../cuda-flash-attention_b200" && pwd)
out=$here/build/variants; mkdir -p $out/$name
for f in fa2_fwd_sm100 fa2_bwd_sm100 fa2_bwd2_sm100 fa2_prepass fa2_api; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $here/csrc/$f.cu -o $out/$name/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/$name.so $out/$name/*.o -lcudart
echo built $out/$name.so
