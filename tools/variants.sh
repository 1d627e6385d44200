#!/bin/bash
# Builds kernel variants for timing experiments: tools/variants.sh name "-DFLAG ..." [name2 "-D..."]...
# -> tools/variants/<name>.so (git-ignored, travels with gpurun); run with tools/variant_scan.py
set -e
cd "$(dirname "$0")/../cropsr_b200/csrc"
mkdir -p ../../tools/variants
while [ $# -ge 2 ]; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -pthread $2 \
    -shared -o ../../tools/variants/$1.so cropsr_b200.cu emit_csv.cpp &
  shift 2
done
wait
ls -la ../../tools/variants
