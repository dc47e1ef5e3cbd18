#!/bin/bash
# Round 2, call 1 (trimmed from r02_first_call.sh): does the level kernel run, is it exact, how fast is it.
out=gpurun_out/r02_call1.txt
mkdir -p gpurun_out
: > $out
LIB=$PWD/build/libtissue_b200_block.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
for cfg in C3 C2 C1; do
  echo "== $cfg product: $(timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror' | head -2)" >> $out
  for path in level level_simple block; do
    echo "== $cfg $path: $(TA_LIB_PATH=$LIB TA_PAIR_PATH=$path timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror' | head -2)" >> $out
  done
done
for path in level level_simple; do
  TA_LIB_PATH=$LIB TA_PAIR_PATH=$path timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_$path.log 2>&1
  echo "parity $path: exit $? | $(tail -1 gpurun_out/r02_parity_$path.log)" >> $out
done
TA_LIB_PATH=$LIB TA_PAIR_PATH=level timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_level_kernel -c 1 \
  -o gpurun_out/r02_level_c3 python tools/profile_scan.py --config C3 --passes 1 > gpurun_out/r02_ncu_level.log 2>&1
echo "ncu level: exit $?" >> $out
cat $out
