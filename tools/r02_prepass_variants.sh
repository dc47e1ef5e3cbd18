#!/bin/bash
# pre-pass tunables: scan ms on C3 / C4 and the classify kernel's own time (ncu launch list) per library variant
for v in "$@"; do
  for cfg in C3 C4; do
    echo "$cfg $v: $(TA_LIB_PATH=$PWD/build/$v.so timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2:')"
  done
  TA_LIB_PATH=$PWD/build/$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:classify --log-file gpurun_out/r02_pp_$v.csv python tools/profile_scan.py --config C3 --passes 2 > /dev/null 2>&1
  echo "   classify on C3: $(grep classify gpurun_out/r02_pp_$v.csv | tail -1 | awk -F'","' '{print $NF}')"
done
