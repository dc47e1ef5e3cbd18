#!/bin/bash
# second-pass kernels after the row-window rewrite: parity suite first, then time and DRAM bytes per launch on C3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "docstring or c1_config or c2_config or ragged or slab or noise" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"map_labels|voxel_first_layer|stencil_|wall_voxels|decode_wall" --log-file gpurun_out/r02_second_pass_c3_v4.csv \
  python tools/second_pass_probe.py --config C3 > gpurun_out/r02_second_pass_c3_v4.log 2>&1; echo "probe exit $?"
tail -2 gpurun_out/r02_second_pass_c3_v4.log
