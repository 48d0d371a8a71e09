#!/bin/bash
# round 2, GPU run 17 (8-GPU node): does the decode bimodality at N=8 come from the CUDA-IPC peer mappings?  control: MSG_NO_P2P=1 (NCCL-only exchange)
mkdir -p gpurun_out/r02
run() { # tag env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 2 --e2e-steps 0 --no-parity --no-cpu-baseline --no-ingest > gpurun_out/r02/n8probe_$1.json 2> gpurun_out/r02/n8probe_$1.err; echo "$1 rc=$?"
  grep '^{' gpurun_out/r02/n8probe_$1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('$1', 'ms', round(d['ms_per_step'],3), 'decode per rank', [round(r['decode_launch_ms'],3) for r in d['ranks']], 'push', [round(r['host_ms_per_step']['push'],1) for r in d['ranks']])"
}
run nop2p MSG_NO_P2P=1
