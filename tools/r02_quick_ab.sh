#!/bin/bash
# scan ms (third pass) of library variants on C3 and C4: tools/r02_quick_ab.sh lib ...
for cfg in C3 C4; do for v in "$@"; do
  echo "$cfg $v: $(TA_LIB_PATH=$PWD/build/$v.so timeout 200 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2:')"
done; done
