#!/bin/bash
# Protocol checker of the CTA-chained engine (compute-sanitizer's racecheck is closed on this GPU pool, profiles/r02_compute_sanitizer_closed.txt).
# Builds swb_chain.cu with -DSWB_CHAIN_CHECK (every table / inbox slot carries the position it holds; every read of the step
# loop checks it) into a scratch copy of the library, runs cfg2 through the four kernel flavours and a set of ragged sizes,
# then repeats with -DSWB_CHAIN_CHECK_BREAK=24 (warps start each group 24 entries early) to show that the checker fires.
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OBJS=$(ls build/csrc/*.o | grep -v swb_chain.o)
cp concurrentproject_b200/lib/libswb200.so /tmp/libswb200.keep
for flags in "-DSWB_CHAIN_CHECK" "-DSWB_CHAIN_CHECK -DSWB_CHAIN_CHECK_BREAK=24"; do
  echo "=== flags: $flags"
  $NV $flags -c concurrentproject_b200/csrc/swb_chain.cu -o /tmp/swb_chain_var.o 2>/dev/null || { echo build failed; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o concurrentproject_b200/lib/libswb200.so $OBJS /tmp/swb_chain_var.o -lpthread
  timeout 300 python bench/chain_check.py > /tmp/chain_check.out 2>&1
  echo "runs: $(grep -c '^chaincheck:' /tmp/chain_check.out), runs with mismatches: $(grep '^chaincheck:' /tmp/chain_check.out | grep -vc ' 0 mismatches')"
  grep -m 6 'chaincheck [TI]' /tmp/chain_check.out
  grep '^chaincheck:' /tmp/chain_check.out | sort | uniq -c | sort -rn | head -12
  grep '^{' /tmp/chain_check.out
done
cp /tmp/libswb200.keep concurrentproject_b200/lib/libswb200.so
