#!/bin/bash
# builds tools/sort_bench against the in-tree libvkgsb.so (run python -m vkgs_b200.build first)
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/sort_bench tools/sort_bench.cu \
  -L vkgs_b200/lib -lvkgsb -Xlinker -rpath -Xlinker '$ORIGIN/../vkgs_b200/lib'
