#!/bin/bash
out=gpurun_out/r02_call7.txt
mkdir -p gpurun_out
: > $out
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k c4 > gpurun_out/r02_c4_test.log 2>&1
echo "c4 fullsize test: exit $? | $(tail -1 gpurun_out/r02_c4_test.log)" >> $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "time_series" > gpurun_out/r02_ts_test.log 2>&1
echo "time series test: exit $? | $(tail -1 gpurun_out/r02_ts_test.log)" >> $out
timeout 900 python tools/timeseries_bench.py --frames 3 > gpurun_out/r02_c5_n1.json 2> gpurun_out/r02_c5_n1.err
echo "c5 n1 (3 frames): exit $? | $(cat gpurun_out/r02_c5_n1.json | cut -c1-600)" >> $out
tail -3 gpurun_out/r02_c5_n1.err >> $out
cat $out
