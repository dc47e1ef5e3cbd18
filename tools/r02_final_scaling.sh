#!/bin/bash
# Final build on one 8-GPU box: C3 strong scaling at N = 8 / 4 / 2 / 1 (parity check inside bench.py), C4 at N = 8
out=gpurun_out/r02_final_scaling.txt
mkdir -p gpurun_out; : > $out
nvidia-smi -L | wc -l >> $out
run() { n=$1; cfg=$2; shift; shift
  f=gpurun_out/r02_final_${cfg}_n${n}
  if [ "$n" = 1 ]; then timeout 900 python bench.py --gpus 1 --config $cfg --steps 20 --warmup 3 "$@" > $f.json 2> $f.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $n --config $cfg --steps 20 --warmup 3 "$@" > $f.json 2> $f.err; fi
  echo "$cfg n$n: exit $? | $(python - <<PY
import json
try:
    d=json.loads(open('$f.json').read().strip().splitlines()[-1])
    print('value %.1f ms/step %.3f scan_ms/rank %s e2e %.1f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_rank'], (d.get('e2e') or {}).get('value', 0), (d.get('parity') or {}).get('merged_equals_single')))
except Exception as e:
    print('no json', e)
PY
)" >> $out
}
run 8 C3 --no-e2e
run 4 C3 --no-e2e
run 2 C3 --no-e2e
run 1 C3 --no-extra --no-cpu-baseline --no-e2e
run 8 C4 --no-e2e
run 1 C4 --no-extra --no-cpu-baseline --no-e2e
cat $out
