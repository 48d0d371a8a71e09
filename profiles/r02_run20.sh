#!/bin/bash
# round 2, GPU run 20: last build (accounting atomics spread over 64 slots): parity tests + a short bench
mkdir -p gpurun_out/r02
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_largeF.py tests/test_gpu_async.py tests/test_cli_gpu.py -q -x > gpurun_out/r02/gpu_tests_run20.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run20.log
tail -3 gpurun_out/r02/gpu_tests_run20.log
python bench.py --records 30000000 --steps 10 --warmup 3 --no-cpu-baseline --no-ingest 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step', round(d['ms_per_step'],3), 'decode', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'alg/launch', d['roofline']['alg_bytes_per_launch'], 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_checked'])"
