#!/bin/bash
# the driver's own sequence on the committed build: smoke(), the GPU suite, the default bench line
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_check_n1.json 2> gpurun_out/r02_bench_check_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_check_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
PY
