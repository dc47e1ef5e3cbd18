#!/bin/bash
# box copies on their own (timing build, flag 0x800): C3 and C4, then the final parity run of the in-tree library
for cfg in C3 C4 C2; do
  echo "== $cfg staging only"
  TA_LIB_PATH=$PWD/build/lib_timing.so timeout 300 python tools/profile_scan.py --config $cfg --passes 3 --flags 2049 2>&1 | grep "pass 2"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
