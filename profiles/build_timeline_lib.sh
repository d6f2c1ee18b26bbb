#!/bin/bash
# Developer build: the GEMM and the implicit positional conv with clock64() role stamps (-DAVI_GEMM_TIMELINE) linked with the product
# objects into build/tl/libavi_b200_tl.so; load it with AVI_B200_LIB=build/tl/libavi_b200_tl.so (profiles/gemm_timeline.py,
# profiles/posconv_timeline.py). Never the product library.
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py >/dev/null
mkdir -p build/tl
for f in gemm_tc2 posconv_tc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DAVI_GEMM_TIMELINE \
       -c avi_talking_b200/csrc/$f.cu -o build/tl/$f.o
done
objs=$(ls build/obj/*.o | grep -v '/gemm_tc2.o' | grep -v '/posconv_tc.o')
nvcc -shared -o build/tl/libavi_b200_tl.so $objs build/tl/gemm_tc2.o build/tl/posconv_tc.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built build/tl/libavi_b200_tl.so
