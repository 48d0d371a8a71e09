#!/bin/bash
# round 2, GPU run 11 (8-GPU node): where does an N=2 step lose ~2 ms on devices (0,1)?  host-side phase timers, device pairs, NCCL-only control
mkdir -p gpurun_out/r02
run() { # tag N env...
  tag=$1; N=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config 12 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02/n2probe_$tag.json 2> gpurun_out/r02/n2probe_$tag.err
  grep '^{' gpurun_out/r02/n2probe_$tag.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', 'ms/step', round(d['ms_per_step'],3), 'host', {k: round(v,3) for k,v in d['run']['host_ms_per_step'].items()})"
}
run default01 2 X=1
run vis01 2 CUDA_VISIBLE_DEVICES=0,1
run vis23 2 CUDA_VISIBLE_DEVICES=2,3
run vis04 2 CUDA_VISIBLE_DEVICES=0,4
run nop2p01 2 MSG_NO_P2P=1
run n1 1 X=1
