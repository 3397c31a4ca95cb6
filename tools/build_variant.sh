#!/bin/bash
# Build an experimental variant of libvkgsb.so next to the product library (git-ignored, travels with gpurun):
#   tools/build_variant.sh <name> "<extra nvcc flags for project.cu>"
# and run anything against it with VKGSB_LIB=vkgs_b200/lib/libvkgsb_<name>.so.  The product library is rebuilt last.
set -e
name=$1; flags=$2
cd "$(dirname "$0")/.."
cp vkgs_b200/lib/libvkgsb.so /tmp/libvkgsb_keep.so 2>/dev/null || true
touch vkgs_b200/csrc/project.cu
VKGSB_PROJECT_FLAGS="$flags" python -m vkgs_b200.build > /dev/null
cp vkgs_b200/lib/libvkgsb.so vkgs_b200/lib/libvkgsb_${name}.so
touch vkgs_b200/csrc/project.cu
python -m vkgs_b200.build > /dev/null
echo "built vkgs_b200/lib/libvkgsb_${name}.so"
