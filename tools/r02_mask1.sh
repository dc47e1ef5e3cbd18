#!/bin/bash
# mask kernel, first GPU run: scan times next to the brick kernel, then the GPU parity suite with the mask kernel
out=gpurun_out/r02_mask1.txt
mkdir -p gpurun_out; : > $out
for cfg in C1 C2 C3; do
  for k in mask brick; do
    echo "== $cfg $k" >> $out
    TA_SCAN_KERNEL=$k timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -4 >> $out
  done
done
echo "== parity (mask kernel)" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $out
for k in mask brick; do
  echo "== C4 $k" >> $out
  TA_SCAN_KERNEL=$k timeout 300 python tools/profile_scan.py --config C4 --passes 2 2>&1 | tail -3 >> $out
done
cat $out
