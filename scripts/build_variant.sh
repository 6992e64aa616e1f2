#!/bin/bash
# scripts/build_variant.sh NAME [nvcc -D flags...]: compiles the working-tree kernel into variants/libstmpm_NAME.so
# (git-ignored, travels with the gpurun snapshot) so that several variants can be timed in one GPU call.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v "$@" \
  -o variants/libstmpm_$name.so startrackermpm_b200/csrc/stmpm.cu 2>&1 | grep -A2 "trials_kernelILi0ELi64ELi0" | grep -E "spill|Used" || true
