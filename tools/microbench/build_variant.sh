#!/bin/bash
# build_variant.sh NAME "<extra nvcc -D flags>" [source tree]   ->  tools/microbench/variants/NAME.so
# One-model (ShockCooling3 unless LCF_DEV_ONLY_MODEL is given in the flags) FP32 build of the library for kernel experiments.
set -e
here=$(cd "$(dirname "$0")" && pwd)
src=${3:-$here/../..}
mkdir -p "$here/variants"
flags="$2"
case "$flags" in *LCF_DEV_ONLY_MODEL*) ;; *) flags="$flags -DLCF_DEV_ONLY_MODEL=3";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags \
     -o "$here/variants/$1.so" "$src/lightcurve_fitting_b200/csrc/lcf_api.cu"
echo "built $here/variants/$1.so"
