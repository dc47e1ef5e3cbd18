#!/bin/bash
bash tools/r02_variants.sh r02_var5 lib_head lib_A6 lib_E
bash tools/r02_phase.sh r02_phase3 lib_timing
echo "== parity" > gpurun_out/r02_parity5.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02_parity5.txt
cat gpurun_out/r02_parity5.txt
