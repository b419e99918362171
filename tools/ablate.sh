#!/bin/bash
# Builds timing-ablation variants of the library (tools only; results are WRONG by construction) next to the product:
#   tools/ablate.sh recursion_mma SMCB_ABLATE 1 2 3     -> smcpp_b200/libsmcpp_b200_abl<N>.so
#   tools/ablate.sh stats32 SMCB_STATS_ABLATE 1 2 3     -> smcpp_b200/libsmcpp_b200_sabl<N>.so
# selected at run time with SMCPP_B200_LIB=...
set -e
cd "$(dirname "$0")/../smcpp_b200/csrc"
src=$1; macro=$2; shift 2
tag=abl; [ "$src" = stats32 ] && tag=sabl
make -s -j8
for n in "$@"; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 -D$macro=$n -c -o build/${src}_$tag$n.o $src.cu
  objs=$(ls build/*.o | grep -v "/${src}\(_s\?abl[0-9]*\)\?\.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsmcpp_b200_$tag$n.so $objs build/${src}_$tag$n.o -lcudart -lquadmath -lpthread -ldl
done
