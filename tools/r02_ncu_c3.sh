#!/bin/bash
# ncu --set full of the in-tree mask kernel on C3 (second launch), report under gpurun_out/$1.ncu-rep
name=${1:-r02_mask_c3}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_kernel -s 1 -c 1 -o gpurun_out/$name -f \
  python tools/profile_scan.py --config ${2:-C3} --passes 2 > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log
