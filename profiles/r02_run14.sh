#!/bin/bash
# round 2, GPU run 14: tests after the walk / writer changes; PropSharing CTAs/SM at the bench size; ncu launch list + full capture of the bench command
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02/gpu_tests_run14.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run14.log
tail -4 gpurun_out/r02/gpu_tests_run14.log
for c in 2 3 4 6; do MSG_EM_CTAS=$c python bench.py --no-cpu-baseline --no-ingest --no-parity --e2e-steps 0 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('EM_CTAS=$c', 'ms/step', round(d['ms_per_step'],3), 'finish host ms', round(d['run']['host_ms_per_step']['finish'],3), 'push', round(d['run']['host_ms_per_step']['push'],3))"; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/launches_bench_default.csv \
    python bench.py --records 20000000 --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-parity --e2e-steps 0 > gpurun_out/r02/ncu_bench_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'decode_kernel|fused_warp_kernel|fused_walk_kernel|em_loop_kernel' -s 8 -c 4 \
    -o gpurun_out/r02/prof_bench_r02 python bench.py --records 20000000 --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-parity --e2e-steps 0 > gpurun_out/r02/ncu_bench_full.log 2>&1
tail -2 gpurun_out/r02/ncu_bench_full.log
python bench.py --config 1 --no-ingest --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg1', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'M aln/s', 'parity', d['parity_checked'])"
