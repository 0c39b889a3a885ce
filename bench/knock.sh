#!/bin/bash
# Measurement builds of the pair engine with step-loop components left out (SWB_KNOCK bit mask, swb_engine.cuh):
#   bench/knock.sh 0 1 2 4 8 16 32 15 63     -> build/knock/libswb200_k<mask>.so
# Only the mode-1 (linear-gap) kernels carry the mask; everything else is the normal objects.  Scores are wrong by design.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/knock
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=$(ls build/csrc/*.o | grep -v swb_kernels_m1.o)
for k in "$@"; do
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSWB_KNOCK=$k \
       -c concurrentproject_b200/csrc/swb_kernels_m1.cu -o build/knock/m1_k$k.o &
done
wait
for k in "$@"; do
  nvcc $ARCH -shared -o build/knock/libswb200_k$k.so $OBJS build/knock/m1_k$k.o -lpthread
done
ls -la build/knock/*.so
