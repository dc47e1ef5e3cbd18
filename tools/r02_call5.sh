#!/bin/bash
# Round 2: after the clean-up -- GPU suite, bench N = 1 with the extras, moments-only / pairs-only scan times
out=gpurun_out/r02_call5.txt
mkdir -p gpurun_out
: > $out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1
echo "gpu tests: exit $? | $(tail -1 gpurun_out/r02_gpu_tests.log)" >> $out
for fl in 7 1 6 2 4; do
  echo "== C3 product flags=$fl: $(timeout 300 python tools/profile_scan.py --config C3 --passes 3 --flags $fl 2>&1 | grep 'pass 2\|rror' | head -2 | tr '\n' ' ')" >> $out
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench n1: exit $? | $(cut -c1-400 gpurun_out/r02_bench_n1.json)" >> $out
tail -3 gpurun_out/r02_bench_n1.err >> $out
cat $out
