#!/bin/bash
# Build the FP64-pipe microbenchmarks for sm_100a (results: profiles/r01_microbench.md).
set -e
cd "$(dirname "$0")"
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 $ARCH -o dispatch dispatch.cu
nvcc -O3 -std=c++17 $ARCH -o operands operands.cu
nvcc -O3 -std=c++17 $ARCH -DNEWTON=2 -o seeds2 seeds.cu
nvcc -O3 -std=c++17 $ARCH -DNEWTON=3 -o seeds3 seeds.cu
nvcc -O3 -std=c++17 $ARCH -o dmma dmma.cu
