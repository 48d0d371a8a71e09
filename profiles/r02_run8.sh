#!/bin/bash
# round 2, GPU run 8: A/B of the decode staging (LDGSTS vs register-staged LDG+STS.128) on resident and on pinned-host (zero-copy) chunks;
# config 4 with the coverage bitmap pinned in L2
mkdir -p gpurun_out/r02
for ST in async ldg; do
  MSG_STAGING=$ST python bench.py --records 30000000 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02/bench_staging_$ST.json 2> gpurun_out/r02/bench_staging_$ST.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_staging_$ST.json').read().strip().split('\n')[-1])
print('$ST', 'resident ms/step', round(d['ms_per_step'],3), 'decode launch ms', round(d['roofline']['launch_ms'],4), 'e2e M aln/s', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'h2d GB/s', round(d['e2e']['h2d_gbs'],1))
PY
done
python bench.py --records 30000000 --steps 5 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/r02/bench_staging_default.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_staging_default.json').read().strip().split('\n')[-1])
print('default', 'resident ms/step', round(d['ms_per_step'],3), 'decode launch ms', round(d['roofline']['launch_ms'],4), 'e2e M aln/s', round(d['e2e']['value'],1), 'h2d GB/s', round(d['e2e']['h2d_gbs'],1))
PY
python profiles/configs_probe.py 20000000 4 10 20 2>&1 | cut -c1-300
MSG_NO_L2_PERSIST=1 python profiles/configs_probe.py 20000000 4 10 20 2>&1 | cut -c1-300
