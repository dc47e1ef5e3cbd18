#!/bin/bash
# mask kernel: scan times (C1..C4) and the GPU parity suite
out=gpurun_out/r02_mask2.txt
mkdir -p gpurun_out; : > $out
for cfg in C1 C2 C3 C4; do
  echo "== $cfg mask" >> $out
  timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -3 >> $out
done
echo "== parity (mask kernel)" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $out
cat $out
