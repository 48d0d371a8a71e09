#!/bin/bash
# round 2, GPU run 19: final decode build (register-staged default at 56 registers / 9 CTAs per SM; LDGSTS variant at 48 / 10)
mkdir -p gpurun_out/r02
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_largeF.py tests/test_gpu_async.py -q -x > gpurun_out/r02/gpu_tests_run19.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run19.log
tail -3 gpurun_out/r02/gpu_tests_run19.log
for ST in ldg async; do
MSG_STAGING=$ST python bench.py --records 30000000 --steps 10 --warmup 3 --no-cpu-baseline --no-ingest --no-parity 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$ST', 'ms/step', round(d['ms_per_step'],3), 'decode', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
done
