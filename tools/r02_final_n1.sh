#!/bin/bash
# Final evidence of the round on one GPU: the bench line, its launch list under ncu, ncu --set full of the scan's three kernels on C3
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02_bench_final_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r02_ncu_launches_bench_n1.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mask_kernel|classify_cores|decide_kernel" -s 3 -c 3 -o gpurun_out/r02_scan_c3_final -f \
  python tools/profile_scan.py --config C3 --passes 2 > gpurun_out/r02_scan_c3_final.log 2>&1; echo "ncu full exit $?"
