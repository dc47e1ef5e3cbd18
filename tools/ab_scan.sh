#!/bin/bash
# A/B timing of scan-kernel build variants on one config: tools/ab_scan.sh C3 [lib.so ...]
cfg=${1:-C3}; shift
echo "== default"; python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep "pass 2"
for lib in "$@"; do
  echo "== $lib"; TA_LIB_PATH=$PWD/$lib python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep "pass 2"
done
