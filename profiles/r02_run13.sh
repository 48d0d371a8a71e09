#!/bin/bash
# round 2, GPU run 13: full GPU test suite + driver-reproducible bench lines for BASELINE configs 1, 3, 4 and the default (5)
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02/gpu_tests_run13.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run13.log
tail -4 gpurun_out/r02/gpu_tests_run13.log
for c in 1 3 4; do
S=$(date +%s); python bench.py --config $c --no-ingest > gpurun_out/r02/bench_cfg${c}_run13.json 2> gpurun_out/r02/bench_cfg${c}_run13.err; echo "cfg $c rc=$? wall=$(( $(date +%s) - S ))s"; tail -2 gpurun_out/r02/bench_cfg${c}_run13.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_cfg${c}_run13.json').read().strip().split('\n')[-1])
print('cfg$c', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'step frac', round(d['roofline']['step']['frac'],3), 'decode frac', round(d['roofline']['frac'],3), 'parity', d['parity_checked'], d['parity'], 'cpu', round(d['cpu_baseline']['value'],2), d['cpu_baseline']['cores'])
PY
done
S=$(date +%s); python bench.py > gpurun_out/r02/bench_default_run13.json 2> gpurun_out/r02/bench_default_run13.err; echo "default rc=$? wall=$(( $(date +%s) - S ))s"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_default_run13.json').read().strip().split('\n')[-1])
print('cfg5', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_checked'])
PY
