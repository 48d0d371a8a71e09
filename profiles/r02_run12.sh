#!/bin/bash
# round 2, GPU run 12: default bench with the ingest leg; PropSharing CTAs/SM sweep on the config-5 shape
mkdir -p gpurun_out/r02
S=$(date +%s); python bench.py > gpurun_out/r02/bench_default_run12.json 2> gpurun_out/r02/bench_default_run12.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
tail -3 gpurun_out/r02/bench_default_run12.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/bench_default_run12.json').read().strip().split('\n')[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))
print('ingest', json.dumps(d.get('ingest'))[:1500])
print('ranks', d['ranks'])
PY
for c in 2 3 4 6 8; do MSG_EM_CTAS=$c python profiles/configs_probe.py 20000000 5 5 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('EM_CTAS=$c', round(d['ms_per_step'],3), 'decode', round(d['decode_ms'],3))"; done
