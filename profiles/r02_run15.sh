#!/bin/bash
# round 2, GPU run 15 (8-GPU node): the default bench at N=8 and N=2 after the timing fix, with per-rank diagnostics
mkdir -p gpurun_out/r02
run() { # N tag extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $1 $3 --no-cpu-baseline --no-ingest > gpurun_out/r02/scale2_$2.json 2> gpurun_out/r02/scale2_$2.err; echo "$2 rc=$?"
  grep '^{' gpurun_out/r02/scale2_$2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('$2', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'parity', d['parity_checked'])
for r in d['ranks']: print('   rank', r['rank'], 'decode', round(r['decode_launch_ms'],4), 'step', round(r['step_ms'],3), 'host', {k: round(v,2) for k,v in r['host_ms_per_step'].items()}, 'sm', r['sm_mhz'], 'mem', r['mem_mhz_min'], 'T', r['temp_c_max'], r['reasons'])"
}
run 8 n8 "--steps 10 --warmup 3"
run 2 n2 "--steps 10 --warmup 3 --e2e-steps 0"
run 1 n1 "--steps 10 --warmup 3 --e2e-steps 0"
