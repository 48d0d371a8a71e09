#!/bin/bash
# round 2, GPU run 18: final state -- full GPU test suite, smoke(), default bench
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02/gpu_tests_run18.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run18.log
tail -3 gpurun_out/r02/gpu_tests_run18.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
S=$(date +%s); python bench.py --no-ingest > gpurun_out/r02/bench_default_run18.json 2> gpurun_out/r02/bench_default_run18.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_default_run18.json').read().strip().split('\n')[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], 'parity', d['parity_checked'], 'cpu', round(d['cpu_baseline']['value'],2))
print(d['ranks'])
PY
