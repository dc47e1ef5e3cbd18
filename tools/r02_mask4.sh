#!/bin/bash
# A/B of the in-tree library against build/lib_head.so (the committed kernel) on C1..C4, then the GPU parity suite
out=gpurun_out/${1:-r02_mask4}.txt
mkdir -p gpurun_out; : > $out
for cfg in C1 C2 C3 C4; do
  echo "== $cfg new" >> $out
  timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -2 >> $out
  echo "== $cfg head" >> $out
  TA_LIB_PATH=$PWD/build/lib_head.so timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -2 | head -1 >> $out
done
echo "== parity" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $out
cat $out
