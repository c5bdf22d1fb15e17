#!/bin/bash
# Build libmmhand_sm100.so (CUDA, sm_100a) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=${MMH_BUILD_OUT:-../libmmhand_sm100.so}
SRCS=$(ls *.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared -o $OUT $SRCS "$@"
echo "built $OUT"
