#!/bin/bash
# Round 2: quick A/B of the record scan on C3 + ncu of both kernels
out=gpurun_out/r02_call4.txt
mkdir -p gpurun_out
: > $out
for cfg in C3 C1; do
  echo "== $cfg rec: $(TA_PAIR_PATH=rec timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror\|pairs' | head -3 | tr '\n' ' ')" >> $out
done
TA_PAIR_PATH=rec timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_rec.log 2>&1
echo "parity rec: exit $? | $(tail -1 gpurun_out/r02_parity_rec.log)" >> $out
TA_PAIR_PATH=rec timeout 600 ncu --set full --clock-control none --import-source on -k regex:rec_ -c 2 \
  -o gpurun_out/r02_rec_c3 -f python tools/profile_scan.py --config C3 --passes 1 > gpurun_out/r02_ncu_rec.log 2>&1
echo "ncu rec: exit $?" >> $out
cat $out
