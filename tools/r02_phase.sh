#!/bin/bash
# phase clocks (timing builds) of library variants on C3 and C2: tools/r02_phase.sh NAME lib ...
name=$1; shift
out=gpurun_out/$name.txt
mkdir -p gpurun_out; : > $out
for cfg in C3 C2; do
  for v in "$@"; do
    echo "== $cfg $v" >> $out
    TA_PHASE_TIMING=1 TA_LIB_PATH=$PWD/build/$v.so timeout 300 python tools/profile_scan.py --config $cfg --passes 2 2>&1 | grep -A3 "mask kernel, first" | tail -4 >> $out
  done
done
cat $out
