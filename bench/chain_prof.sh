#!/bin/bash
# Builds the chain kernels with -DSWB_CHAIN_PROF plus each flag set in "$@" into a scratch copy of the library and runs cfg2.
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OBJS=$(ls build/csrc/*.o | grep -v swb_chain.o)
cp concurrentproject_b200/lib/libswb200.so /tmp/libswb200.keep
for flags in "$@"; do
  echo "=== flags: $flags"
  $NV -DSWB_CHAIN_PROF $flags -c concurrentproject_b200/csrc/swb_chain.cu -o /tmp/swb_chain_prof.o 2>/dev/null || { echo build failed; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o concurrentproject_b200/lib/libswb200.so $OBJS /tmp/swb_chain_prof.o -lpthread
  timeout 120 python bench/one.py 100000 2 2 '{"config": 7, "rows": 3}' 2>&1 | grep -E "cta 40|chainstart|engine_ms" | tail -40 | sed -E "s/.*(engine_ms.: [0-9.]+).*/\1/" | cut -c1-200
done
cp /tmp/libswb200.keep concurrentproject_b200/lib/libswb200.so
