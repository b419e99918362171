#!/bin/bash
# Builds timing-ablation variants of the library (tools only; results are WRONG by construction) next to the product:
#   smcpp_b200/libsmcpp_b200_abl<N>.so with -DSMCB_ABLATE=<N> in recursion_mma.cu, selected with SMCPP_B200_LIB=...
set -e
cd "$(dirname "$0")/../smcpp_b200/csrc"
make -s -j8
for n in "$@"; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 -DSMCB_ABLATE=$n -c -o build/recursion_mma_abl$n.o recursion_mma.cu
  objs=$(ls build/*.o | grep -v "recursion_mma\(_abl[0-9]*\)\?\.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsmcpp_b200_abl$n.so $objs build/recursion_mma_abl$n.o -lcudart -lquadmath -lpthread -ldl
done
