#!/bin/bash
# Round 2, call 3: the two-kernel record scan (TA_PAIR_PATH=rec): parity, scan times, ncu.
out=gpurun_out/r02_call3.txt
mkdir -p gpurun_out
: > $out
LIB2=$PWD/build/libtissue_b200_m2.so
for cfg in C3 C2 C1; do
  echo "== $cfg product: $(timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
  echo "== $cfg rec (3 CTAs/SM): $(TA_PAIR_PATH=rec timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
  echo "== $cfg rec (2 CTAs/SM): $(TA_LIB_PATH=$LIB2 TA_PAIR_PATH=rec timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
done
TA_PAIR_PATH=rec timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_rec.log 2>&1
echo "parity rec: exit $? | $(tail -1 gpurun_out/r02_parity_rec.log)" >> $out
TA_PAIR_PATH=rec timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r02_fullsize_rec.log 2>&1
echo "fullsize rec: exit $? | $(tail -1 gpurun_out/r02_fullsize_rec.log)" >> $out
TA_PAIR_PATH=rec timeout 600 ncu --set full --clock-control none --import-source on -k regex:rec_ -c 2 \
  -o gpurun_out/r02_rec_c3 python tools/profile_scan.py --config C3 --passes 1 > gpurun_out/r02_ncu_rec.log 2>&1
echo "ncu rec: exit $?" >> $out
cat $out
