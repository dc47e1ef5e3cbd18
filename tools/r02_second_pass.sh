#!/bin/bash
# second-pass kernels on C3 (time and DRAM bytes per launch), then ncu --set full of the uint32 mask kernel on C4
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"map_labels|voxel_first_layer|stencil_image|wall_voxels|decode_wall" --log-file gpurun_out/r02_second_pass_c3.csv \
  python tools/second_pass_probe.py --config C3 > gpurun_out/r02_second_pass_c3.log 2>&1; echo "probe exit $?"
tail -2 gpurun_out/r02_second_pass_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mask_kernel -s 1 -c 1 -o gpurun_out/r02_mask_c4_final -f \
  python tools/profile_scan.py --config C4 --passes 2 > gpurun_out/r02_mask_c4_final.log 2>&1; echo "ncu c4 exit $?"
