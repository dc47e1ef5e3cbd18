#!/bin/bash
# scan ms (third pass) of library variants build/lib_*.so given as arguments, on C2 C3 C4
# usage: tools/r02_variants.sh NAME lib_head lib_A ...
name=$1; shift
out=gpurun_out/$name.txt
mkdir -p gpurun_out; : > $out
for cfg in C2 C3 C4; do
  for v in "$@"; do
    t=$(TA_LIB_PATH=$PWD/build/$v.so timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep "pass 2:" | sed 's/pass 2: //')
    echo "$cfg $v: $t" >> $out
  done
done
cat $out
