#!/bin/bash
# last single-GPU call of the round: the per-voxel path after its rewrite (C1), a regression look at C3, parity, the bench line
out=gpurun_out/r02_last.txt; : > $out
for cfg in C1 C3; do
  echo "$cfg: $(timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2')" >> $out
done
echo "C1 round-1 kernel: $(TA_SCAN_KERNEL=brick timeout 300 python tools/profile_scan.py --config C1 --passes 3 2>&1 | grep 'pass 2')" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
cat $out
timeout 1200 python bench.py > gpurun_out/r02_bench_last_n1.json 2> gpurun_out/r02_bench_last_n1.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02_bench_last_n1.json
