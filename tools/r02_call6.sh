#!/bin/bash
# Round 2: multi-GPU (run under gpurun --gpus 4): NCCL parity test, bench N = 2 / 4 with the merged == single check,
# deferred vs synchronous merge, one launch behind the exchange vs overlapped interior scan
out=gpurun_out/r02_call6.txt
mkdir -p gpurun_out
: > $out
nvidia-smi -L | wc -l >> $out
timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r02_gpu_dist_test.log 2>&1
echo "dist test: exit $? | $(tail -1 gpurun_out/r02_gpu_dist_test.log)" >> $out
tr() { # tr N tag args...
  n=$1; tag=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/r02_bench_n${n}_$tag.json 2> gpurun_out/r02_bench_n${n}_$tag.err
  echo "bench n$n $tag: exit $? | $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n${n}_$tag.json').read().strip().splitlines()[-1])
    print('value %.1f ms/step %.3f scan_ms/rank %s e2e %.1f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_rank'], (d.get('e2e') or {}).get('value', 0), (d.get('parity') or {}).get('merged_equals_single')))
except Exception as e:
    print('no json', e)
PY
)" >> $out
}
tr 2 deferred
tr 2 sync --sync-merge --no-parity --no-e2e
tr 2 overlap --overlap --no-parity --no-e2e
tr 4 deferred
tr 4 sync --sync-merge --no-parity --no-e2e
tr 4 overlap --overlap --no-parity --no-e2e
tail -5 gpurun_out/r02_bench_n4_deferred.err >> $out
cat $out
