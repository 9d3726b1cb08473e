#!/bin/bash
# Build an engine variant for A/B timing (profiles/kernel_ab.py):
#     profiles/build_variant.sh NAME [-DFLAG=VALUE ...]   ->  profiles/_variants/NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/_variants
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
     -ccbin /usr/bin/g++ -o profiles/_variants/$name.so optnetaligncpp_b200/csrc/ona_engine.cu "$@"
echo built profiles/_variants/$name.so
