#!/bin/bash
# Alternative builds of liboqg for A/B runs on the GPU box (tools/ab.sh):
#   tools/build_ab.sh NAME -DMACRO=VALUE ...   ->  build_ab/NAME.so
NAME=$1; shift
mkdir -p build_ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -w "$@" -o build_ab/$NAME.so oq_engine_b200/csrc/oqg.cu && echo built build_ab/$NAME.so
