#!/bin/bash
# Developer build: the GEMM with clock64() role stamps (-DAVI_GEMM_TIMELINE) linked with the product objects into
# build/tl/libavi_b200_tl.so; load it with AVI_B200_LIB=build/tl/libavi_b200_tl.so (profiles/gemm_timeline.py). Never the product library.
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py >/dev/null
mkdir -p build/tl
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DAVI_GEMM_TIMELINE \
     -c avi_talking_b200/csrc/gemm_tc2.cu -o build/tl/gemm_tc2.o
objs=$(ls build/obj/*.o | grep -v '/gemm_tc2.o')
nvcc -shared -o build/tl/libavi_b200_tl.so $objs build/tl/gemm_tc2.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built build/tl/libavi_b200_tl.so
