#!/bin/bash
# variants (default pre-pass setting each), pre-pass on / off for the in-tree library, launch list, parity suite
bash tools/r02_variants.sh r02_var6 lib_head lib_A6 lib_P lib_X
out=gpurun_out/r02_prepass4.txt; : > $out
for cfg in C2 C3 C4; do
  for pp in 0 1; do
    t=$(TA_PREPASS=$pp timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep "pass 2" | sed 's/pass 2: //')
    echo "$cfg prepass=$pp: $t" >> $out
  done
done
cat $out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_prepass_launches4.csv python tools/profile_scan.py --config C3 --passes 2 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_parity6.txt; cat gpurun_out/r02_parity6.txt
