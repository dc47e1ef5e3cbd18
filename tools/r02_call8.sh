#!/bin/bash
# Round 2, 8-GPU box: C3 strong scaling with the parity check, C4 (16 GiB uint32) at N = 1 / 2 / 4 / 8, C5 time series on 8 ranks
out=gpurun_out/r02_call8.txt
mkdir -p gpurun_out
: > $out
nvidia-smi -L | wc -l >> $out
run() { # run N config tag extra...
  n=$1; cfg=$2; tag=$3; shift; shift; shift
  f=gpurun_out/r02_${cfg}_n${n}_$tag
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --config $cfg --steps 10 --warmup 3 "$@" > $f.json 2> $f.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $n --config $cfg --steps 20 --warmup 3 "$@" > $f.json 2> $f.err
  fi
  echo "$cfg n$n $tag: exit $? | $(python - <<PY
import json
try:
    d=json.loads(open('$f.json').read().strip().splitlines()[-1])
    print('value %.1f ms/step %.3f scan_ms/rank %s e2e %.1f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_rank'], (d.get('e2e') or {}).get('value', 0), (d.get('parity') or {}).get('merged_equals_single')))
except Exception as e:
    print('no json', e)
PY
)" >> $out
}
run 8 C3 final
run 4 C3 final --no-e2e
run 2 C3 final --no-e2e
run 1 C3 final --no-extra --no-cpu-baseline --no-e2e
run 8 C4 final
run 4 C4 final --no-e2e
run 2 C4 final --no-e2e
run 1 C4 final --no-extra --no-cpu-baseline --no-e2e
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 tools/timeseries_bench.py --frames 10 > gpurun_out/r02_c5_n8.json 2> gpurun_out/r02_c5_n8.err
echo "c5 n8: exit $? | $(cut -c1-500 gpurun_out/r02_c5_n8.json)" >> $out
cat $out
