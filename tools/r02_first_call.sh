#!/bin/bash
# First GPU call of the next round, as ONE gpurun command: parity and C3 / C2 / C4 scan times of the experimental block and
# level kernels next to the product kernel, then an ncu capture of the level kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/r02_first_call.sh'
# Build the experiment libraries BEFORE the call (they travel with the snapshot; build/ is git-ignored only):
#   TA_NVCC_EXTRA=-DTA_WITH_BLOCK_KERNEL TA_OUT=$PWD/build/libtissue_b200_block.so bash tissue_analysis_b200/csrc/build.sh
#   TA_NVCC_EXTRA="-DTA_WITH_BLOCK_KERNEL -DTA_LEVEL_MINB=2" TA_OUT=$PWD/build/libtissue_b200_block_2cta.so bash tissue_analysis_b200/csrc/build.sh
#   TA_NVCC_EXTRA="-DTA_WITH_BLOCK_KERNEL -DTA_LEVEL_MAXL=5" TA_OUT=$PWD/build/libtissue_b200_block_maxl5.so bash tissue_analysis_b200/csrc/build.sh
out=gpurun_out/r02_first_call.txt
mkdir -p gpurun_out
: > $out
LIB=$PWD/build/libtissue_b200_block.so
LIB2=$PWD/build/libtissue_b200_block_2cta.so
LIB3=$PWD/build/libtissue_b200_block_maxl5.so
[ -f $LIB ] || TA_NVCC_EXTRA=-DTA_WITH_BLOCK_KERNEL TA_OUT=$LIB bash tissue_analysis_b200/csrc/build.sh
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
# 1. parity of every experimental kernel on the existing suite (the C ABI is the same; TA_PAIR_PATH picks the kernel)
for path in level level_pf level_simple block block_simple; do
  TA_LIB_PATH=$LIB TA_PAIR_PATH=$path timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_$path.log 2>&1
  echo "parity $path: exit $? | $(tail -1 gpurun_out/r02_parity_$path.log)" >> $out
done
# 2. scan times (third pass of profile_scan.py)
for cfg in C3 C2 C4 C1; do
  echo "== $cfg product: $(timeout 600 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2')" >> $out
  for path in level level_pf level_simple block; do
    echo "== $cfg $path: $(TA_LIB_PATH=$LIB TA_PAIR_PATH=$path timeout 600 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror' | head -2)" >> $out
  done
  if [ -f $LIB2 ]; then
    echo "== $cfg level, 2 CTAs/SM build: $(TA_LIB_PATH=$LIB2 TA_PAIR_PATH=level timeout 600 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror' | head -2)" >> $out
  fi
  if [ -f $LIB3 ]; then
    echo "== $cfg level, 5 labels per block by masks: $(TA_LIB_PATH=$LIB3 TA_PAIR_PATH=level timeout 600 python tools/profile_scan.py --config $cfg --passes 3 2>&1 | grep 'pass 2\|rror' | head -2)" >> $out
  fi
done
# 3. full-size parity of the level kernel (conservation laws, slab split, C oracle slab)
TA_LIB_PATH=$LIB TA_PAIR_PATH=level timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r02_fullsize_level.log 2>&1
echo "fullsize level: exit $? | $(tail -1 gpurun_out/r02_fullsize_level.log)" >> $out
# 3b. the bench line with the level kernel in place of the product kernel (same step, same roofline arithmetic)
TA_LIB_PATH=$LIB TA_PAIR_PATH=level timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_level_n1.json 2> gpurun_out/r02_bench_level_n1.err
echo "bench level: exit $? | $(cut -c1-260 gpurun_out/r02_bench_level_n1.json)" >> $out
# 4. ncu: launch list, then the full set for the level kernel on C3 (only after the plain runs above)
TA_LIB_PATH=$LIB TA_PAIR_PATH=level timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_level_kernel -c 1 \
  -o gpurun_out/r02_level_c3 python tools/profile_scan.py --config C3 --passes 1 > gpurun_out/r02_ncu_level.log 2>&1
echo "ncu level: exit $?" >> $out
cat $out
