#!/bin/bash
# pre-pass on (default from 4096 bricks) / off: scan ms of the third pass on C2 C3 C4; then its parity test
out=gpurun_out/${1:-r02_prepass}.txt
: > $out
for cfg in C2 C3 C4; do
  for pp in 0 1; do
    t=$(TA_PREPASS=$pp timeout 300 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep "pass 2:" | sed 's/pass 2: //')
    echo "$cfg prepass=$pp: $t" >> $out
  done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "prepass or variants or c1_config" 2>&1 | tail -5 >> $out
cat $out
