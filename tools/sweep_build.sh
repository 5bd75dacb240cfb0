#!/bin/bash
# usage: tools/sweep_build.sh <tag> <extra nvcc flags...>  -> gpurun_out/libcf_<tag>.so
tag=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo "$@" -Xcompiler -fPIC -shared \
  -o gpurun_out/libcf_$tag.so clusterfusion_b200/csrc/llama_decoder.cu -lcudart
