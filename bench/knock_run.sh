for k in 0 1 2 4 8 15 32 63; do
  if [ $k = 0 ]; then lib=concurrentproject_b200/lib/libswb200.so; else lib=build/knock/libswb200_k$k.so; fi
  echo "== knock $k"
  SWB200_LIB=$PWD/$lib DIAG_CONFIGS=1 DIAG_ROWS=3,4,8 python bench/diag.py knock$k cfgsweep 2>&1 | grep "'lin'" | grep "'pure': 3"
done
