#!/bin/bash
# Build an experimental variant of libvkgsb.so next to the product library (git-ignored, travels with gpurun):
#   tools/build_variant.sh <name> "<extra nvcc flags>" [project|blend|bin|sort]     (default: project.cu)
# and run anything against it with VKGSB_LIB=vkgs_b200/lib/libvkgsb_<name>.so.  The product library is rebuilt last.
set -e
name=$1; flags=$2; what=${3:-project}
cd "$(dirname "$0")/.."
var=VKGSB_$(echo $what | tr a-z A-Z)_FLAGS
touch vkgs_b200/csrc/$what.cu
env $var="$flags" python -m vkgs_b200.build > /dev/null
cp vkgs_b200/lib/libvkgsb.so vkgs_b200/lib/libvkgsb_${name}.so
touch vkgs_b200/csrc/$what.cu
python -m vkgs_b200.build > /dev/null
echo "built vkgs_b200/lib/libvkgsb_${name}.so"
