#!/bin/bash
# Like chain_variants.sh, but the flags also reach swb200.cu (variants that change the launch shape, e.g. -DSWB_CHAIN_HELPER_WARP=7).
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OBJS=$(ls build/csrc/*.o | grep -v "swb_chain.o\|swb200.o")
cp concurrentproject_b200/lib/libswb200.so /tmp/libswb200.keep
for flags in "$@"; do
  echo "=== flags: $flags"
  $NV $flags -c concurrentproject_b200/csrc/swb_chain.cu -o /tmp/swb_chain_var.o 2>/dev/null &
  $NV $flags -c concurrentproject_b200/csrc/swb200.cu -o /tmp/swb200_var.o 2>/dev/null &
  wait
  [ -f /tmp/swb_chain_var.o ] && [ -f /tmp/swb200_var.o ] || { echo build failed; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o concurrentproject_b200/lib/libswb200.so $OBJS /tmp/swb_chain_var.o /tmp/swb200_var.o -lpthread
  timeout 200 python bench/chain_stress.py ${REPS:-30} 2>&1 | tail -4
  rm -f /tmp/swb_chain_var.o /tmp/swb200_var.o
done
cp /tmp/libswb200.keep concurrentproject_b200/lib/libswb200.so
