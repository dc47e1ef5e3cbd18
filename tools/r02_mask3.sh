#!/bin/bash
# mask kernel with the one-label look-ahead from global memory (default) against the committed kernel (build/lib_head.so):
# scan times, phase clocks, ncu --set full of the default on C3
out=gpurun_out/r02_mask3.txt
mkdir -p gpurun_out; : > $out
for cfg in C2 C3 C4; do
  echo "== $cfg look-ahead" >> $out
  timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -2 >> $out
  echo "== $cfg head" >> $out
  TA_LIB_PATH=$PWD/build/lib_head.so timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -2 >> $out
done
for cfg in C2 C3 C4; do
  echo "== $cfg look-ahead phase clocks" >> $out
  TA_PHASE_TIMING=1 TA_LIB_PATH=$PWD/build/lib_timing.so timeout 300 python tools/profile_scan.py --config $cfg --passes 2 2>&1 | tail -4 >> $out
  echo "== $cfg head phase clocks" >> $out
  TA_PHASE_TIMING=1 TA_LIB_PATH=$PWD/build/lib_head_timing.so timeout 300 python tools/profile_scan.py --config $cfg --passes 2 2>&1 | tail -4 >> $out
done
cat $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_kernel -s 1 -c 1 -o gpurun_out/r02_mask_c3v5 -f \
  python tools/profile_scan.py --config C3 --passes 2 > gpurun_out/r02_mask3_ncu.log 2>&1
tail -3 gpurun_out/r02_mask3_ncu.log
