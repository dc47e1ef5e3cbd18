#!/bin/bash
# Round 2, call 2: the record kernel (TA_PAIR_PATH=meta): parity, scan times, ncu.
out=gpurun_out/r02_call2.txt
mkdir -p gpurun_out
: > $out
LIB2=$PWD/build/libtissue_b200_m2.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
for cfg in C1 C3 C2; do
  echo "== $cfg product: $(timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
  echo "== $cfg meta (3 CTAs/SM): $(TA_PAIR_PATH=meta timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
  echo "== $cfg meta (2 CTAs/SM): $(TA_LIB_PATH=$LIB2 TA_PAIR_PATH=meta timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
done
TA_PAIR_PATH=meta timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_meta.log 2>&1
echo "parity meta: exit $? | $(tail -1 gpurun_out/r02_parity_meta.log)" >> $out
TA_PAIR_PATH=meta timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r02_fullsize_meta.log 2>&1
echo "fullsize meta: exit $? | $(tail -1 gpurun_out/r02_fullsize_meta.log)" >> $out
TA_PAIR_PATH=meta timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_meta_kernel -c 1 \
  -o gpurun_out/r02_meta_c3 python tools/profile_scan.py --config C3 --passes 1 > gpurun_out/r02_ncu_meta.log 2>&1
echo "ncu meta: exit $?" >> $out
cat $out
