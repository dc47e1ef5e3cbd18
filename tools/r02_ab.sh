#!/bin/bash
# A/B of the in-tree library against build/lib_head.so on C2..C4 (scan ms of the third pass), then ncu --set full on C3
# usage: tools/r02_ab.sh NAME [parity]
name=${1:-r02_ab}
out=gpurun_out/$name.txt
mkdir -p gpurun_out; : > $out
for cfg in C2 C3 C4; do
  echo "== $cfg new" >> $out
  timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -2 | head -1 >> $out
  echo "== $cfg head" >> $out
  TA_LIB_PATH=$PWD/build/lib_head.so timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | tail -2 | head -1 >> $out
done
if [ "$2" = parity ]; then
  echo "== parity" >> $out
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $out
fi
cat $out
bash tools/r02_ncu_c3.sh ${name}_c3
